#!/usr/bin/env python
"""Times DensePolynomial::commit (src/dense_mlpoly.rs:214-239) on the device at the witness
geometries of the BASELINE configs: a section of 2^ell scalars is 2^floor(ell/2) row commitments
over 2^ceil(ell/2) generators. Prints one JSON line per size: seconds, scalars/s, point additions/s.

  python scripts/commit_bench.py --ell 22 24 26 [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ell", type=int, nargs="+", default=[22, 24, 26])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import spartan_parallel_b200 as sp

    def device_gens(ctx, label, n):
        """MultiCommitGens::new (src/commitments.rs:15-33): SHAKE256(label || basepoint) in 64-byte blocks"""
        import hashlib

        base = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
        return sp.MultiCommitGens.from_uniform(ctx, hashlib.shake_256(label + base).digest(64 * (n + 1)))

    ctx = sp.Context(0)
    rng = np.random.default_rng(7)
    for ell in args.ell:
        n = 1 << ell
        L, R = 1 << (ell // 2), 1 << (ell - ell // 2)
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        poly = sp.DensePolynomial.new(ctx, a)
        del a
        t0 = time.perf_counter()
        gens = device_gens(ctx, b"gens_r1cs_sat", R)
        ctx.sync()
        t_gens = time.perf_counter() - t0
        t0 = time.perf_counter()
        gens.prepare(R, L)
        t_table = time.perf_counter() - t0
        times = []
        first = None
        for _ in range(args.reps + 1):
            ctx.sync()
            t0 = time.perf_counter()
            rows = gens.commit_poly(poly, L)
            times.append(time.perf_counter() - t0)
            if first is None:
                first = rows
            assert rows == first
        best = min(times[1:])
        info = gens.info()
        rows_tab = info.pop("rows_table")
        if rows_tab["table_bytes"]:  # the many-row path (single-window table + Horner) served the commitment
            info = {"path": "k_msm_hrows", "per_window_table_bytes": info["table_bytes"], **rows_tab}
        else:
            info["path"] = "k_msm_rows"
        adds = info["adds_per_scalar"]
        print(json.dumps({"ell": ell, "rows": L, "cols": R, "gens_s": t_gens, "table_s": t_table, "first_call_s": times[0], "commit_s": best,
                          "scalars_per_s": n / best, "point_adds_per_s": n * adds / best, "adds_per_scalar": adds, **info}), flush=True)
        poly.free()
        gens.free()


if __name__ == "__main__":
    main()
