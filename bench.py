#!/usr/bin/env python
"""Benchmark of the accelerated path: the two sumchecks of R1CSProof::prove (with the
table builders that feed them) on a synthetic data-parallel R1CS batch, plus the witness
commitment and the whole proof around them.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference loops

metric: sumcheck constraints/sec = sum_p Q_p * X_p / time(one full pass of
z_mat -> SpMV -> phase-1 rounds -> ABC/Z tables -> phase-2 rounds), per-round host
round trips included. One JSON line on stdout (rank 0). The same line carries BASELINE's other
metric, prove time: witness commitment (polycommit) + R1CSProof::prove (+ the sparse-polynomial
evaluation proof) with the reference's Timer labels, and a roofline entry per kernel class.

Workload (BASELINE.json configs[4], "2^20 x 64"): P = 1 instance, X = 2^20 constraints,
Q = 64 proofs, sections (u, v), constraint x: u_x * u_{x+1} = v_x. Under torchrun the N ranks
prove that ONE batch sharded over the proof axis (strong scaling: 64 / N proofs per GPU; the
default for N > 1) -- per round 3 scalars per rank cross a host mailbox, and the rq-bound Z table
is summed once over NVLink peer memory (parallel.py); the rows of the witness commitment are
sharded the same way. --scaling weak keeps 64 proofs PER GPU instead (a batch of 64 N proofs);
in the default mode that figure is measured too and reported under "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_MOD = (1 << 252) + 27742317777372353535851937790883648493


def log2(n):
    return n.bit_length() - 1


# ----------------------------------------------------------------------------- inputs
def random_canonical(rng, n):
    """n scalars whose Montgomery limbs are uniform below 2^252 (< q): valid `Scalar`s."""
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def challenges(rng, n):
    return random_canonical(rng, max(n, 1))[:n]


def synthetic_matrices(X, one):
    rows = np.arange(X, dtype=np.uint32)
    ones = np.tile(one, (X, 1))
    A = (rows, rows.copy(), ones)
    B = (rows, ((rows + 1) % X).astype(np.uint32), ones)
    Cm = (rows, (rows + X).astype(np.uint32), ones)
    return A, B, Cm


ZERO = np.zeros(4, dtype=np.uint64)
ONE = np.array([0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF], dtype=np.uint64)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.samples = []
        self.stop_flag = False
        self.thread = None

    def _run(self):
        """NVML in-process when the bindings load (one cheap query per sample); the nvidia-smi
        subprocess otherwise. A fresh nvidia-smi per sample re-initialises NVML over every GPU of
        the box and contends with the ranks' kernel launches, so it is the fallback only."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.device)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.05)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
# kernel -> (algorithmic bytes per item as the launcher counts them in `units`, Montgomery products
# per item); DESIGN.md section 4 states both per kernel. MSM kernels count point additions instead.
KERNEL_MODEL = {
    "k_rows_rolled": (576.0, 10.0),      # fused bind_j + eval_{j+1}: 4 read + 2 written scalars x 3 tables, 6 binds + 4
    "k_rows_spmv": (288.0, 2.0),         # SpMV + round 0 for a satisfying witness (spg_sc1_set_satisfied): z read + 3 tables written per pair, the point t = 2 only: 1 product + 1 weighted sum (4 with spg_sc1_set_claim)
    "k_rows": (192.0, 4.0),
    "k_quad_bind_eval": (576.0, 13.0),
    "k_quad_split": (576.0, 13.0),       # the same round with one item over eight lanes (late rounds: latency, not throughput)
    "k_pair_eval": (192.0, 7.0),
    "k_pair_bind": (288.0, 3.0),
    "k_z_bind_rq": (32.0, 74.0 / 104.0),  # one product per witness scalar read, as four-term dot products: 296 wide multiplies per 4 instead of 4 x 104
    "k_cubic_eval_rlc": (192.0, 6.0),
    "k_cubic_eval_split": (192.0, 6.0),
    "k_cubic_bind_eval_rlc": (288.0, 12.0),    # fused bind_j + eval_{j+1}: 12 scalars read, 6 written, 6 binds + 6 products per item and triple
    "k_cubic_bind_eval_split": (288.0, 12.0),
    "k_multi_bind_top": (96.0, 1.0),
    "k_prod_layer": (96.0, 1.0),
    "k_hash_layer_fq": (128.0, 2.0),
    "k_eq_expand": (96.0, 1.0),
}
MSM_WIDE_IMAD_PER_ADD = 7 * 72.0  # mixed addition: 7 field products of 64 + 8 IMAD.WIDE.U32 (csrc/fe8.cuh)


def kernel_rooflines(prof, hbm_peak, int_peak, imad_peak):
    """One entry per kernel of a profile: achieved GB/s of algorithmic bytes against the measured HBM
    peak and, where the model above knows the kernel, modmul/s against the measured integer peak; for
    the MSM kernels point additions/s against the IMAD.WIDE issue bound. `bound` names the larger
    of the two fractions."""
    out = []
    for r in sorted(prof, key=lambda r: -r["total_ms"]):
        name, t = r["kernel"], r["total_ms"] * 1e-3
        if t <= 0 or not r["units"]:
            continue
        e = {"kernel": name, "launches": r["launches"], "total_ms": r["total_ms"]}
        if name.startswith("k_msm") or name.startswith("(k_msm"):
            adds = r["units"] / t
            e.update({"point_adds_per_s": adds, "bound": "int_pipe",
                      "frac": (adds * MSM_WIDE_IMAD_PER_ADD / imad_peak) if imad_peak else None,
                      "peak": "IMAD.WIDE.U32 issue rate (tools/imad_peak.cu) / 504 per mixed addition"})
        else:
            gbs = r["units"] / t / 1e9
            e.update({"achieved_GBps": gbs, "hbm_frac": gbs / hbm_peak})
            key = next((k for k in sorted(KERNEL_MODEL, key=len, reverse=True) if k in name), None)
            if key and int_peak:
                bpi, mpi = KERNEL_MODEL[key]
                mm = r["units"] / bpi * mpi / t
                e.update({"modmul_per_s": mm, "int_frac": mm / int_peak})
            e["bound"] = "int_pipe" if e.get("int_frac", 0) > e["hbm_frac"] else "hbm"
            e["frac"] = max(e.get("int_frac", 0), e["hbm_frac"])
        out.append(e)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import spartan_parallel_b200 as sp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from spartan_parallel_b200 import parallel

    numa = parallel.bind_host_to_gpu(local)  # before any pinned allocation (first-touch pages)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=dev)
    ctx = sp.Context(local)
    scaling = args.scaling or "strong"
    X = 1 << args.log_x
    nx = args.log_x
    ng = log2(world)
    assert world & (world - 1) == 0, "world size must be a power of two"
    comm = parallel.ShmComm(device=dev) if world > 1 else parallel.LocalComm()
    peer = parallel.PeerTable(ctx, comm, 2 * X) if world > 1 else None  # the rq-bound Z table: W * Y scalars
    shard_y = world > 1 and not args.replicated_phase2  # W = 2 sections, world a power of two >= 2: chunks stay inside a section
    A, B, Cm = synthetic_matrices(X, ONE)
    inst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    trace, tracing = [], [False]
    nosync_trace = bool(os.environ.get("SPG_BENCH_TRACE_NOSYNC"))  # development: host timestamps in the timed loop

    def mark(name):
        if tracing[0]:
            ctx.sync()
            trace.append((name, time.perf_counter()))
        elif nosync_trace:
            trace.append((name, time.perf_counter()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.sync()
        e1.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms, wall * 1e3)

    class Shard:
        """This rank's slice of one batch of Q_total proofs (proofs rank * Q .. rank * Q + Q - 1)."""

        def __init__(self, Q_total, seed):
            self.Q_total, self.Q = Q_total, Q_total // world
            assert self.Q >= 1 and self.Q * world == Q_total, f"{Q_total} proofs do not split over {world} ranks"
            Q, N = self.Q, X * self.Q
            self.N, self.nq = N, log2(Q)
            rng = np.random.default_rng(seed + rank)
            # witness: u random, v = u * roll(u) computed with the library (no CPU field code here)
            u = random_canonical(rng, N)
            du = sp.DensePolynomial.new(ctx, u)
            u_next = np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -1, axis=1).reshape(N, 4))
            dun = sp.DensePolynomial.new(ctx, u_next)
            dv = sp.vec_op(ctx, "mul", du, dun)
            v = dv.to_host()
            del du, dun, dv, u_next
            # pinned host copies for the end-to-end leg
            self.hu = torch.from_numpy(u.view(np.int64)).pin_memory()
            self.hv = torch.from_numpy(v.view(np.int64)).pin_memory()
            self.u_pin, self.v_pin = self.hu.numpy().view(np.uint64), self.hv.numpy().view(np.uint64)
            # challenges are common to all ranks (in production they come from rank 0's transcript)
            crng = np.random.default_rng(0xC4A11E46E)
            nq = self.nq
            self.tau_q, self.tau_x = challenges(crng, nq + ng), challenges(crng, nx)
            self.ch1, self.ch2 = challenges(crng, nx + nq + ng), challenges(crng, 1 + nx)
            self.r_abc = challenges(crng, 3)

        def upload(self, asynchronous=False):
            Q = self.Q
            return [sp.ProverWitnessSecInfo(ctx, [Q], [X], self.u_pin, asynchronous),
                    sp.ProverWitnessSecInfo(ctx, [Q], [X], self.v_pin, asynchronous)]

        def one_pass(self, secs):
            """The hot path for one batch: everything R1CSProof::prove does on tables."""
            Q, nq, ch1, ch2, r_abc = self.Q, self.nq, self.ch1, self.ch2, self.r_abc
            mark("start")
            z = sp.ZMat(ctx, [Q], [X], secs)
            rx = ch1[:nx][::-1].copy()
            if world == 1:
                sc1 = sp.sumcheck_phase1(ctx, inst, z, [Q], Q, [X], X, X, self.tau_q[:0], self.tau_q, self.tau_x)
                sc1.set_satisfied()  # claim_phase1 = 0 (src/r1csproof.rs:330) and the witness satisfies the instance row by row
                sc1.run_rounds(ch1[:sc1.num_rounds])  # C loop: eval -> host -> bind per round, no Python in between
                c1 = sc1.final()
                sc1.free()
                sc2 = sp.SumcheckPhase2(ctx, inst, z, [Q], Q, [X], X, 2, rx, ch1[nx:nx + nq], ch1[:0], r_abc[0], r_abc[1], r_abc[2])
            else:
                # one proof over Q * world proofs: shards exchange 3 scalars per round, then one
                # modular reduction of the rq-bound Z table over peer memory; phase 2 (independent of Q) is cut over y
                sc1 = parallel.gpu_phase1(ctx, comm, inst, z, Q, X, X, self.tau_q, self.tau_x, satisfied=True)
                mark("phase1 create")
                sc1.run_rounds(ch1[:sc1.num_rounds])
                mark("phase1 rounds (local C loop + tail)")
                c1 = sc1.final()
                sc1.free()
                mark("phase1 final")
                zrq = parallel.gpu_bind_rq_sharded(ctx, comm, z, ch1[nx:nx + nq + ng], Q, peer, scatter=shard_y)
                mark("Z bind + peer reduce")
                if shard_y:  # phase 2 cut over y: each rank builds and folds 1/world of the ABC and Z tables
                    sc2 = parallel.gpu_phase2_sharded(ctx, comm, inst, zrq, X, 2, rx, r_abc[0], r_abc[1], r_abc[2])
                else:
                    sc2 = sp.SumcheckPhase2.from_zrq(ctx, inst, zrq, [X], X, 2, rx, ch1[:0], r_abc[0], r_abc[1], r_abc[2])
            mark("phase2 create")
            sc2.run_rounds(ch2[:sc2.num_rounds])
            c2 = sc2.final()
            mark("phase2 rounds")
            sc2.free()
            z.free()
            mark("free")
            return c1, c2

        def measure_resident(self, steps, warmup, profile=True):
            secs = self.upload()
            for _ in range(warmup):
                first = self.one_pass(secs)
            launches0 = ctx.launches
            if profile:
                ctx.profile_begin()
            del trace[:]
            ms, wall = timed(lambda: self.one_pass(secs), steps)
            prof = ctx.profile_end() if profile else []
            launches = (ctx.launches - launches0) // max(steps, 1)
            for s in secs:
                s.free()
            return {"ms": ms / steps, "wall_ms": wall / steps, "prof": prof, "launches": int(launches), "first": first}

        def measure_e2e(self, steps, warmup):
            """host buffers in, claims out, every batch's copies inside the timed region. Batches are
            double-buffered: the H2D copy of batch i+1 (copy stream) overlaps the proving of batch i
            (compute stream); the first batch's copy is not overlapped."""
            outs = []

            def run(k):
                nxt = self.upload(True)
                for i in range(k):
                    cur, nxt = nxt, (self.upload(True) if i + 1 < k else None)
                    outs.append(self.one_pass(cur))
                    for s in cur:
                        s.free()

            run(min(warmup, 2))
            ms, _ = timed(lambda: run(steps), 1)
            return ms / steps, outs[-1]

    t_setup = time.time()
    Q_main = args.proofs if scaling == "strong" else args.proofs * world
    main = Shard(Q_main, 0x5EED0000)
    setup_s = time.time() - t_setup
    total_units = X * Q_main

    # ---- N > 1: the sharded path must reproduce the unsharded one bit for bit (C3, 2^16 x 256)
    parity = None
    if world > 1 and not args.no_parity:
        try:
            parity = sharded_equals_unsharded(sp, parallel, ctx, comm, rank, world, 16, 256)
        except Exception as e:
            parity = {"ok": False, "error": str(e)[:200]}
        ok = max_over_ranks(0.0 if parity.get("ok") else 1.0)[0] == 0.0
        parity["ok"] = ok

    # ---- device-resident leg
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("SPG_NO_CLOCKS"):
        sampler.start()
    res = main.measure_resident(args.steps, args.warmup, profile=not os.environ.get("SPG_BENCH_NO_PROFILE"))
    if nosync_trace and rank in (0, world - 1):
        print(f"[rank {rank}] timed loop, host clock, no syncs: " + ", ".join(f"{n}: {(t - trace[i][1]) * 1e3:.2f}" for i, (n, t) in enumerate(trace[1:])),
              file=sys.stderr, flush=True)
    # ---- end-to-end leg
    if args.no_e2e:  # development runs only: the reported line then carries no end-to-end number
        e2e_ms, last = float("nan"), res["first"]
    else:
        e2e_ms, last = main.measure_e2e(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(res["first"][0], last[0]) and np.array_equal(res["first"][1], last[1])
    phase_trace = None
    if args.trace_phases or world > 1:
        # wall clock per phase of one more pass (a device sync between the phases, so the sum is a little above
        # ms_per_step): at N > 1 it goes into the line as `phase_trace_ms` -- what limits the strong scaling
        secs = main.upload()
        barrier()
        del trace[:]
        tracing[0] = True
        main.one_pass(secs)
        tracing[0] = False
        phase_trace = {n: round((t - trace[i][1]) * 1e3, 3) for i, (n, t) in enumerate(trace[1:])}
        if args.trace_phases:
            print(f"[rank {rank}] " + ", ".join(f"{n}: {v:.2f} ms" for n, v in phase_trace.items()), file=sys.stderr, flush=True)
        for s_ in secs:
            s_.free()

    # ---- witness commitment (polycommit): DensePolynomial::commit of both sections, rows sharded
    # over the ranks exactly like the proofs (a rank's proofs are a contiguous row range), 32 bytes
    # per row gathered over NCCL. Generators and their window tables are setup (SNARKGens).
    from spartan_parallel_b200 import host

    commit = None
    gens = None
    ell = nx + log2(Q_main)
    rows_total, R = 1 << (ell // 2), 1 << (ell - ell // 2)
    rows_local = rows_total // world
    if not args.no_commit and rows_local >= 1 and (main.N % R) == 0:
        t0 = time.perf_counter()
        gens = host.R1CSGens(ctx, b"gens_r1cs_sat", X * Q_main)
        pc = gens.gens_pc()
        pc.prepare(R, rows_local)
        ctx.sync()
        gens_s = time.perf_counter() - t0
        secs = main.upload()
        polys = [s.poly_w(0) for s in secs]

        def commit_all():
            out = []
            for poly in polys:
                mine = pc.commit_poly_rows(poly, rows_local, 0, rows_local)
                if world > 1:
                    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
                    full = torch.empty(world * t.numel(), dtype=torch.uint8, device=dev)
                    dist.all_gather_into_tensor(full, t)
                    mine = full.cpu().numpy().tobytes()
                out.append(mine)
            return out

        first_c = commit_all()
        ctx.profile_begin()
        reps = 3
        ms_c, wall_c = timed(commit_all, reps)
        prof_c = ctx.profile_end()
        assert commit_all() == first_c
        info = pc.info()
        rows_tab = info.pop("rows_table")
        if rows_tab["table_bytes"]:  # the many-row path served the commitment: report ITS window and additions
            info = {"window_bits": rows_tab["window_bits"], "adds_per_scalar": rows_tab["adds_per_scalar"],
                    "table_bytes": rows_tab["table_bytes"] + info["table_bytes"], "table_bases": rows_tab["table_bases"],
                    "tables": "single-window table per base + Horner over the windows (k_msm_hrows) for the row commitments; "
                              f"per-window table (c = {info['window_bits']}, {info['table_bytes']} bytes) for the few-row MSMs of the openings"}
        commit = {"seconds": wall_c * 1e-3 / reps, "device_ms": ms_c / reps, "sections": 2, "scalars": 2 * X * Q_main,
                  "rows_per_section": rows_total, "cols": R, "rows_per_rank": rows_local,
                  "scalars_per_s": 2 * X * Q_main / (wall_c * 1e-3 / reps), **info,
                  "setup_s": gens_s, "prof": prof_c,
                  "what": "DensePolynomial::commit (src/dense_mlpoly.rs:214-239) of the two witness sections, compressed row commitments on the host"
                          + ("" if world == 1 else f"; rows sharded over {world} ranks, all-gather of 32 B per row inside the timed region")}
        for s_ in secs:
            s_.free()

    # ---- whole R1CSProof::prove (transcript, sigma protocols, openings) through the C++ host mirror
    full_proof = None
    if world == 1 and not args.no_full_proof:
        try:
            if gens is None:
                gens = host.R1CSGens(ctx, b"gens_r1cs_sat", main.N)
            secs = main.upload()
            seed = np.array([1, 2, 3, 4], dtype=np.uint64)
            Q = main.Q
            times, stages = [], None
            for _ in range(3):
                host.timings_reset()
                t0 = time.perf_counter()
                blob, _ = host.r1cs_prove(ctx, inst, secs, [Q], Q, [X], X, b"bench", b"gens_r1cs_sat", seed, main.N, gens)
                times.append(time.perf_counter() - t0)
                stages = host.timings()
            for s in secs:
                s.free()
            st = {}
            for k, v in stages:
                st[k] = st.get(k, 0.0) + v * 1e-3
            g = lambda *ks: sum(st.get(k, 0.0) for k in ks)
            full_proof = {"seconds": min(times[1:]), "proof_bytes": len(blob),
                          "timers": {"prove_sc_phase_one": g("z_mat + SpMV + sc1 setup", "phase-1 rounds (zk)"),
                                     "sigma_protocols_between_phases": g("sigma protocols"),
                                     "prove_sc_phase_two": g("sc2 setup (ABC, Z bind)", "phase-2 rounds (zk)"),
                                     "polyeval": g("witness evaluations", "opening proofs"),
                                     "serialization": g("tail + serialization")},
                          "what": "R1CSProof::prove end to end for the same batch (witness resident): both ZK sumchecks with their "
                                  "per-round sigma protocols on the host mirror, witness evaluations, Hyrax openings with the bullet "
                                  "reduction MSMs on the device; timers are the reference's labels (src/r1csproof.rs:274,421,518) "
                                  "from the last of 3 runs, `seconds` the best of the last 2"}
        except Exception as e:  # never lose the bench line over the extra measurement
            full_proof = {"error": str(e)[:200]}
    if gens is not None:
        gens.free()
        gens = None

    # ---- sparse-polynomial evaluation proof (memory-check product trees), C5's second half
    sparse = None
    if world == 1 and args.with_sparse:
        try:
            sparse = sparse_leg(sp, host, ctx, args.sparse_log_nnz)
        except Exception as e:
            sparse = {"error": str(e)[:200]}

    witness_gen = None
    if world == 1 and not args.no_witness_gen:
        try:
            witness_gen = witness_gen_leg(sp, ctx)
        except Exception as e:
            witness_gen = {"error": str(e)[:200]}

    # ---- the general row path of the fused SpMV + first round: 3 non-unit entries per row and matrix
    general_rows = None
    if world == 1 and not args.no_general_rows:
        try:
            general_rows = general_rows_leg(sp, ctx, args.log_x, min(args.proofs, 16), args.steps, args.warmup)
        except Exception as e:
            general_rows = {"error": str(e)[:200]}

    # ---- N > 1 in strong mode: the weak figure (proofs per GPU fixed) as an extra
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        del main.hu, main.hv, main.u_pin, main.v_pin
        w = Shard(args.proofs * world, 0x5EED1000)
        k = max(3, args.steps // 4)
        r = w.measure_resident(k, 3, profile=False)
        weak = {"value": X * w.Q_total / (r["ms"] * 1e-3), "unit": "constraints/s", "ms_per_step": r["ms"], "steps": k,
                "workload": f"X=2^{args.log_x} x Q={args.proofs} proofs PER GPU ({w.Q_total} in the batch)"}
        del w

    if world > 1:  # orderly teardown of the peer mappings, the mailbox and the process group
        try:
            peer.close()
            comm.close()
            dist.destroy_process_group()
        except Exception:
            pass
    if rank != 0:
        return
    step_ms, prof = res["ms"], res["prof"]
    value = total_units / (step_ms * 1e-3)
    e2e_value = total_units / (e2e_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    int_peak = imad_peak = None
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "r1_imad_peak.json")))
        int_peak, imad_peak = ip["modmul_lazy_ilp2_t256_per_s"], ip["imad_wide_carry_per_s"]
    except Exception:
        pass
    dom = max(prof, key=lambda r: r["total_ms"]) if prof else None
    roofline = None
    if dom:
        per_launch_bytes = dom["units"] / max(dom["launches"], 1)
        avg_ms = dom["total_ms"] / max(dom["launches"], 1)
        achieved = dom["units"] / (dom["total_ms"] * 1e-3) / 1e9 if dom["total_ms"] else 0.0
        roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_ms,
                    "launches_per_step": dom["launches"] / args.steps,
                    "share_of_step": dom["total_ms"] / (step_ms * args.steps)}
        key = next((k for k in sorted(KERNEL_MODEL, key=len, reverse=True) if k in dom["kernel"]), None)
        if key and int_peak:
            bpi, mpi = KERNEL_MODEL[key]
            mm = dom["units"] / bpi * mpi / (dom["total_ms"] * 1e-3)
            roofline["int_pipe"] = {"achieved_modmul_per_s": mm, "peak_modmul_per_s": int_peak, "frac": mm / int_peak,
                                    "peak_source": "tools/imad_peak.cu in-register fq_mul_lazy rate (profiles/r1_imad_peak.json)"}
        tr = ncu_traffic(dom["kernel"])
        if tr:
            roofline["traffic"], roofline["traffic_note"] = tr
    by_kernel = kernel_rooflines(prof, hbm_peak, int_peak, imad_peak)
    for e in by_kernel:
        e["ms_per_step"] = e.pop("total_ms") / args.steps
        e["launches_per_step"] = e.pop("launches") / args.steps
    if commit:
        cp = commit.pop("prof")
        for e in kernel_rooflines(cp, hbm_peak, int_peak, imad_peak):
            e["ms_per_commit"] = e.pop("total_ms") / 3
            e["launches_per_commit"] = e.pop("launches") / 3
            e["leg"] = "polycommit"
            by_kernel.append(e)
    if sparse and "prof" in sparse:
        for e in kernel_rooflines(sparse.pop("prof"), hbm_peak, int_peak, imad_peak)[:6]:
            e["leg"] = "sparse proof"
            by_kernel.append(e)
    cpu = cpu_baseline_sample(args, threads=1)
    # BASELINE metric (1): prove time
    prove_time = {"polycommit": commit["seconds"] if commit else None}
    if world == 1:
        prove_time["R1CSProof::prove"] = full_proof.get("seconds") if full_proof else None
        if full_proof and "timers" in full_proof:
            prove_time.update(full_proof["timers"])
        if sparse and "seconds" in sparse:
            prove_time["SparseMatPolyEvalProof::prove"] = sparse["seconds"]
            prove_time.update(sparse.get("timers", {}))
        parts = [prove_time["polycommit"], prove_time["R1CSProof::prove"], prove_time.get("SparseMatPolyEvalProof::prove", 0.0)]
        prove_time["what"] = "witness commitment + R1CSProof::prove" + (" + SparseMatPolyEvalProof::prove (3 matrices x 2^%d non-zeros)" % args.sparse_log_nnz if sparse and "seconds" in sparse else "") + ", wall clock, witness resident in HBM"
    else:
        prove_time["sumcheck_table_pass"] = step_ms * 1e-3
        parts = [prove_time["polycommit"], prove_time["sumcheck_table_pass"]]
        prove_time["what"] = ("sharded legs only: witness commitment (rows over ranks) + the table pass of both sumchecks; the transcript-side "
                              "host mirror (sigma protocols, openings) drives one GPU and is timed at N = 1")
    prove_time_s = sum(parts) if all(p is not None for p in parts) else None
    shard_note = ("single GPU" if world == 1 else
                  f"one batch of {Q_main} proofs sharded by proof index over {world} ranks ({main.Q} per rank): per-round exchange of 3 scalars per rank "
                  "through host shared memory (the values already live in pinned host memory) + one modular all-reduce of the rq-bound Z table as a "
                  "kernel over NVLink peer memory (CUDA IPC; rank r sums chunk r with P2P loads" +
                  ("; phase 2 is sharded over y the same way: rank r builds and folds chunk r of the ABC and Z tables, the last log2(N) rounds run on the gathered scalars)"
                   if shard_y else " and writes it to every peer with P2P stores; phase 2 replicated)"))
    line = {
        "metric": "sumcheck_constraints_per_sec", "value": value, "unit": "constraints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "u256 (F_q, 8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"data-parallel R1CS batch, X=2^{args.log_x} constraints x Q={Q_main} proofs in total ({main.Q} per GPU), P=1 instance, W=2 sections (BASELINE configs[4] shape)",
                   "constraints_per_step": total_units, "sharding": shard_note,
                   "e2e_pipeline": "double-buffered: the H2D copy of batch i+1 overlaps the proving of batch i; all copies are inside the timed region",
                   "l2": f"inputs ({2 * main.N * 32 / 2**30:.2f} GiB per GPU per step) exceed the 126 MB L2; no flush needed",
                   "first_round": "spg_sc1_set_satisfied: claim 0 and Az Bz - Cz = 0 row by row (a satisfying witness, what R1CSProof::prove is called with), so the fused SpMV + first round evaluates t = 2 only", "challenges": "precomputed per-round challenges replayed by a C loop (spg_sc1_run_rounds / spg_sc2_run_rounds); every round still returns its 3 evaluations to the host (96 B) before the bind with that round's challenge (32 B) is issued; `value` is therefore the rate of the table work, `prove_time` the whole protocol",
                   "phases": "z_mat + SpMV + phase-1 rounds + ABC/Z tables + phase-2 rounds"},
        "e2e": {"value": e2e_value, "unit": "constraints/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(2 * main.N * 32 * world), "d2h_bytes_per_step": int(96 * (2 * nx + main.nq + ng + 1) + 7 * 32)},
        "gpu_launches": res["launches"], "full_proof": full_proof, "wall_ms_per_step": res["wall_ms"],
        "prove_time_s": prove_time_s, "prove_time": prove_time, "polycommit": commit, "sparse_proof": sparse, "witness_gen": witness_gen, "general_rows": general_rows, "phase_trace_ms": phase_trace,
        "clocks": clocks, "roofline": roofline, "roofline_by_kernel": by_kernel, "cpu_baseline": cpu,
        "kernels": sorted(prof, key=lambda r: -r["total_ms"])[:8], "setup_s": setup_s, "host_affinity": numa,
    }
    if parity is not None:
        line["parity"] = bool(parity.get("ok"))
        line["parity_check"] = parity
    if weak is not None:
        line["weak"] = weak
    if world > 1 and phase_trace:
        fixed = sum(v for n, v in phase_trace.items() if "rounds" not in n)
        line["scaling_limiter"] = (
            f"per-round latency: a pass is {2 * nx + main.nq + ng + 1} sequential rounds, each a kernel + 96 bytes to the host + a "
            f"{world}-rank exchange before the next challenge (~20 us on the late rounds whose tables are a few thousand scalars), plus "
            f"{fixed:.2f} ms per pass outside the round loops (prover set-up, Z bind + peer reduce-scatter with two host barriers); "
            "see phase_trace_ms and DESIGN.md section 5")
    emit(line)


def ncu_traffic(kernel_name):
    """dram bytes of the largest launch of this kernel from the newest committed `ncu --set full`
    summary under profiles/ (a cross-reference measured under the profiler, not part of this run)."""
    import glob

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full*.txt")), reverse=True):
        try:
            rd = wr = None
            seen = False
            for ln in open(path):
                if ln.startswith("== launch"):
                    if seen and rd and wr:
                        break
                    seen = kernel_name.split("<")[0] in ln
                    rd = wr = None
                elif seen and (ln.startswith("dram__bytes_read.sum =") or ln.startswith("dram__bytes_write.sum =")):
                    val, unit = ln.split("=")[1].split()[:2]
                    b = float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
                    if "read" in ln.split("=")[0]:
                        rd = b
                    else:
                        wr = b
            if seen and rd and wr:
                best = (rd + wr, f"largest launch of {kernel_name.split('<')[0]} in {os.path.basename(path)} (ncu --set full, captured separately)")
                break
        except Exception:
            continue
    return best


def sharded_equals_unsharded(sp, parallel, ctx, comm, rank, world, log_x, Q):
    """BASELINE config C3 (X = 2^16, Q = 256) proven sharded over all ranks and unsharded on this
    rank's own GPU from the same seeded batch: every round polynomial of both sumchecks and both
    claim vectors must be bit-identical (src/sumcheck.rs:1166-1275, src/r1csproof.rs:469-501)."""
    X, Ql = 1 << log_x, Q // world
    nq = log2(Q)
    rng = np.random.default_rng(77)  # the same stream on every rank
    u = random_canonical(rng, X * Q)
    du = sp.DensePolynomial.new(ctx, u)
    dun = sp.DensePolynomial.new(ctx, np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -1, axis=1).reshape(X * Q, 4)))
    v = sp.vec_op(ctx, "mul", du, dun).to_host()
    del du, dun
    A, B, Cm = synthetic_matrices(X, ONE)
    inst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    tau_q, tau_x = challenges(rng, nq), challenges(rng, log_x)
    ch1, ch2, r_abc = challenges(rng, log_x + nq), challenges(rng, 1 + log_x), challenges(rng, 3)
    none = ch1[:0]
    rx = ch1[:log_x][::-1].copy()
    secs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], u), sp.ProverWitnessSecInfo(ctx, [Q], [X], v)]
    z = sp.ZMat(ctx, [Q], [X], secs)
    sc1 = sp.sumcheck_phase1(ctx, inst, z, [Q], Q, [X], X, X, none, tau_q, tau_x)
    want1, wantc1 = sc1.run_rounds(ch1), None
    wantc1 = sc1.final()
    sc1.free()
    sc2 = sp.SumcheckPhase2(ctx, inst, z, [Q], Q, [X], X, 2, rx, ch1[log_x:], none, *r_abc)
    want2 = sc2.run_rounds(ch2)
    wantc2 = sc2.final()
    sc2.free()
    z.free()
    for s_ in secs:
        s_.free()
    lo, hi = rank * Ql * X, (rank + 1) * Ql * X
    secs = [sp.ProverWitnessSecInfo(ctx, [Ql], [X], u[lo:hi]), sp.ProverWitnessSecInfo(ctx, [Ql], [X], v[lo:hi])]
    z = sp.ZMat(ctx, [Ql], [X], secs)
    peer = parallel.PeerTable(ctx, comm, 2 * X)
    sh = parallel.gpu_phase1(ctx, comm, inst, z, Ql, X, X, tau_q, tau_x, satisfied=True)
    got1 = sh.run_rounds(ch1)
    gotc1 = sh.final()
    sh.free()
    zrq = parallel.gpu_bind_rq_sharded(ctx, comm, z, ch1[log_x:], Ql, peer)
    sc2 = sp.SumcheckPhase2.from_zrq(ctx, inst, zrq, [X], X, 2, rx, none, *r_abc)
    got2 = sc2.run_rounds(ch2)
    gotc2 = sc2.final()
    sc2.free()
    ok_y = True
    if world <= 2 * X and (2 * X) % world == 0 and world % 2 == 0:  # the y-sharded phase 2 must agree as well
        zrq = parallel.gpu_bind_rq_sharded(ctx, comm, z, ch1[log_x:], Ql, peer, scatter=True)
        sy = parallel.gpu_phase2_sharded(ctx, comm, inst, zrq, X, 2, rx, *r_abc)
        ok_y = np.array_equal(sy.run_rounds(ch2), want2) and np.array_equal(sy.final(), wantc2)
        sy.free()
    peer.close()
    z.free()
    for s_ in secs:
        s_.free()
    inst.free()
    ok = (np.array_equal(got1, want1) and np.array_equal(gotc1, wantc1) and np.array_equal(got2, want2)
          and np.array_equal(gotc2, wantc2) and ok_y)
    return {"ok": bool(ok), "config": f"C3: X=2^{log_x} x Q={Q} sharded over {world} ranks vs unsharded on each rank's GPU",
            "compared": f"{len(want1)} + {len(want2)} round polynomials, 4 + 3 final claims, bit for bit"}


def general_rows_leg(sp, ctx, log_x, Q, steps, warmup):
    """The table pass on an instance whose rows have THREE entries per matrix with non-unit coefficients
    (the synthetic north-star instance has one unit entry per row, which the fused SpMV + first round
    serves from its fast path; real block circuits do not). Constraint x:
      (a1 u_x + a2 u_{x+1} + a3 u_{x+2}) * (b1 u_{x+3} + b2 u_{x+4} + b3 u_{x+5}) = v_x + c2 u_x + c3 u_{x+7}
    with v solved for on the device. Timed like the main leg and against the unit instance at the same size."""
    X, nx, nq = 1 << log_x, log_x, log2(Q)
    N = X * Q
    rng = np.random.default_rng(0x6E7A1)
    coef = random_canonical(rng, 8)  # a1 a2 a3 b1 b2 b3 c2 c3
    rows = np.arange(X, dtype=np.uint32)

    def mat(entries):
        r = np.concatenate([rows for _ in entries])
        c = np.concatenate([((rows + sh) % X + base).astype(np.uint32) for sh, base, _ in entries])
        v = np.concatenate([np.tile(val, (X, 1)) for _, _, val in entries])
        return r, c, v

    A = mat([(0, 0, coef[0]), (1, 0, coef[1]), (2, 0, coef[2])])
    B = mat([(3, 0, coef[3]), (4, 0, coef[4]), (5, 0, coef[5])])
    Cm = mat([(0, X, ONE), (0, 0, coef[6]), (7, 0, coef[7])])
    inst3 = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    A1, B1, C1 = synthetic_matrices(X, ONE)
    inst1 = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A1], [B1], [C1])
    u = random_canonical(rng, N)
    du = sp.DensePolynomial.new(ctx, u)

    def shifted(sh):
        return sp.DensePolynomial.new(ctx, np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -sh, axis=1).reshape(N, 4)))

    def lin(terms):  # sum of coefficient * shifted u, on the device
        acc = None
        for sh, c in terms:
            t = sp.vec_op(ctx, "mul", shifted(sh), sp.DensePolynomial.new(ctx, np.tile(c, (N, 1))))
            acc = t if acc is None else sp.vec_op(ctx, "add", acc, t)
        return acc

    az = lin([(0, coef[0]), (1, coef[1]), (2, coef[2])])
    bz = lin([(3, coef[3]), (4, coef[4]), (5, coef[5])])
    v3 = sp.vec_op(ctx, "sub", sp.vec_op(ctx, "mul", az, bz), lin([(0, coef[6]), (7, coef[7])])).to_host()
    v1 = sp.vec_op(ctx, "mul", du, shifted(1)).to_host()
    del az, bz, du
    crng = np.random.default_rng(0xC4A11E46E)
    tau_q, tau_x = challenges(crng, nq), challenges(crng, nx)
    ch1, ch2, r_abc = challenges(crng, nx + nq), challenges(crng, 1 + nx), challenges(crng, 3)

    def run(inst, v, checked):
        secs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], u), sp.ProverWitnessSecInfo(ctx, [Q], [X], v)]

        def one_pass(check=False):
            z = sp.ZMat(ctx, [Q], [X], secs)
            sc1 = sp.sumcheck_phase1(ctx, inst, z, [Q], Q, [X], X, X, tau_q[:0], tau_q, tau_x)
            if check:
                sc1.set_claim_checked(ZERO)  # round 0 recomputes the sum: an unsatisfied witness is an error
            else:
                sc1.set_claim(ZERO)
            sc1.run_rounds(ch1)
            c1 = sc1.final()
            sc1.free()
            sc2 = sp.SumcheckPhase2(ctx, inst, z, [Q], Q, [X], X, 2, ch1[:nx][::-1].copy(), ch1[nx:], ch1[:0], *r_abc)
            sc2.run_rounds(ch2)
            c2 = sc2.final()
            sc2.free()
            z.free()
            return c1, c2

        first = one_pass(checked)
        for _ in range(max(warmup - 1, 1)):
            again = one_pass()
        assert all(np.array_equal(a, b) for a, b in zip(first, again)), "checked and unchecked passes differ"
        ctx.profile_begin()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass()
        ctx.sync()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        prof = ctx.profile_end()
        for s_ in secs:
            s_.free()
        spmv = [r for r in prof if "spmv" in r["kernel"]]
        return ms, (sum(r["total_ms"] for r in spmv) / steps if spmv else None), [r["kernel"] for r in spmv]

    ms3, spmv3, names3 = run(inst3, v3, True)
    ms1, spmv1, _ = run(inst1, v1, True)
    # products of the fused SpMV + first round: one per non-unit entry and proof, plus 4 per pair of rows
    modmul = Q * X * 8 + 4 * (N // 2)
    int_peak = None
    try:
        int_peak = json.load(open(os.path.join(ROOT, "profiles", "r1_imad_peak.json")))["modmul_lazy_ilp2_t256_per_s"]
    except Exception:
        pass
    return {"ms_per_pass": ms3, "constraints_per_s": N / (ms3 * 1e-3), "spmv_first_round_ms": spmv3, "kernels": names3,
            "spmv_first_round_modmul_per_s": modmul / (spmv3 * 1e-3) if spmv3 else None,
            "spmv_first_round_int_frac": (modmul / (spmv3 * 1e-3) / int_peak) if spmv3 and int_peak else None,
            "unit_instance_ms_per_pass": ms1, "unit_instance_spmv_first_round_ms": spmv1, "ratio": ms3 / ms1,
            "workload": f"X=2^{log_x} x Q={Q}, 3 entries per row and matrix, random coefficients (8 of 9 non-unit), witness satisfying; "
                        "the first pass verifies the claim against the tables (spg_sc1_set_claim_checked)",
            "what": "table pass (z_mat, SpMV + round 0, phase-1 rounds, ABC / Z tables, phase-2 rounds), wall clock, witness resident"}


def witness_gen_leg(sp, ctx, log_rows=18, n=16, phy=4, vir=2):
    """SNARK::prove's derived block sections on the device (src/lib.rs:1481-1613, 1667-1676): block_w2,
    block_w3 and block_w3_shifted computed from block_vars, so that only the primary section crosses
    PCIe. Reports the device time next to the bytes that no longer need uploading."""
    rows = 1 << log_rows
    io_width = 2 * n
    vars_width = 1 << (io_width + 2 * phy + 4 * vir - 1).bit_length()
    w2_width = 1 << (2 * n + 2 * phy + 4 * vir - 1).bit_length()
    rng = np.random.default_rng(11)
    vars_ = random_canonical(rng, rows * vars_width)
    vars_[::vars_width] = ONE  # validity column
    tau, r = challenges(rng, 2)
    d_vars = sp.DensePolynomial.new(ctx, vars_)
    w0 = sp.wit_perm_w0(ctx, tau, r, 2 * n, 2 * n)
    best = None
    for _ in range(4):
        ctx.sync()
        t0 = time.perf_counter()
        w2, w3 = sp.wit_block(ctx, d_vars, rows, vars_width, w0, tau, r, n, io_width, phy, vir, w2_width)
        sh = sp.wit_shift(ctx, w3, rows)
        ctx.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        w2.free(); w3.free(); sh.free()
    primary, derived = rows * vars_width * 32, rows * (w2_width + 16) * 32
    return {"seconds": best, "rows": rows, "primary_bytes": primary, "derived_bytes_not_uploaded": derived,
            "derived_share_of_upload": derived / (primary + derived), "derived_GBps": derived / best / 1e9,
            "what": f"block_w2 ({w2_width} wide), block_w3 and block_w3_shifted of one block instance with {rows} executions, "
                    f"{n} inputs, {phy} physical and {vir} virtual memory operations, derived on the device from block_vars "
                    f"({vars_width} wide); PCIe moves ~54 GB/s, so uploading them instead would take {derived / 54e9:.3f} s"}


def sparse_leg(sp, host, ctx, lg):
    """SparseMatPolynomial::multi_commit + SparseMatPolyEvalProof::prove (src/sparse_mlpoly.rs:566-586,
    1497-1564) for three matrices of 2^lg non-zeros over 2^lg rows x 2^(lg+1) columns."""
    batch, nvx, nvy = 3, lg, lg + 1
    rng = np.random.default_rng(3)
    nnz = 1 << lg
    polys = []
    for _ in range(batch):
        rows = rng.integers(0, 1 << nvx, size=nnz).astype(np.uint32)
        cols = rng.integers(0, 1 << nvy, size=nnz).astype(np.uint32)
        polys.append((rows, cols, random_canonical(rng, nnz)))
    rx, ry = challenges(rng, nvx), challenges(rng, nvy)
    hrx, hry = sp.EqPolynomial(ctx, rx).evals().to_host(), sp.EqPolynomial(ctx, ry).evals().to_host()
    evals = []
    for rows, cols, vals in polys:  # M_i(rx, ry) on the device
        a, b, v = sp.DensePolynomial.new(ctx, hrx[rows]), sp.DensePolynomial.new(ctx, hry[cols]), sp.DensePolynomial.new(ctx, vals)
        evals.append(sp.dot(ctx, sp.vec_op(ctx, "mul", a, b), v))
    seed = np.array([1, 2, 3, 4], dtype=np.uint64)
    best, stages, prof = None, None, []
    t0 = time.perf_counter()
    gens = host.SparseGens(ctx, b"gens_sparse_poly", nvx, nvy, nnz, batch)  # setup: SNARKGens::new
    gens_s = time.perf_counter() - t0
    for it in range(3):
        host.timings_reset()
        if it == 2:
            ctx.profile_begin()
        t0 = time.perf_counter()
        comm, proof = host.sparse_prove(ctx, polys, nvx, nvy, rx, ry, np.stack(evals), b"bench", b"gens_sparse_poly", seed, gens)
        dt = time.perf_counter() - t0
        if it == 2:
            prof = ctx.profile_end()
        if it and (best is None or dt < best):
            best, stages = dt, host.timings()
    st = {}
    for k, v in stages:
        st[k] = st.get(k, 0.0) + v * 1e-3
    g = lambda *ks: sum(st.get("sparse: " + k, 0.0) for k in ks)
    pre = g("dense representation", "generators", "multi_commit")
    prove = g("eq tables + derefs", "derefs commitment", "hash layers + trees", "tree evals + dotp", "product-circuit sumchecks",
              "hash-layer evals + 3 openings")
    gens.free()
    # `seconds` is SparseMatPolyEvalProof::prove itself (src/sparse_mlpoly.rs:1497-1564: it receives the dense
    # representation built at encode time); what the python call adds on top -- marshalling the three COO
    # matrices (120 MB) into the library -- is instance setup and reported apart
    return {"gens_setup_s": gens_s, "seconds": prove, "with_preprocessing_s": prove + pre, "call_s": best,
            "instance_marshalling_s": best - prove - pre, "proof_bytes": len(proof), "commitment_bytes": len(comm),
            "timers": {"commit_nondet_witness": g("eq tables + derefs", "derefs commitment"),
                       "build_layered_network": g("hash layers + trees"),
                       "evalproof_layered_network": g("tree evals + dotp", "product-circuit sumchecks"),
                       "hash_layer_evals_and_openings": g("hash-layer evals + 3 openings")},
            "preprocessing": {"dense_representation": g("dense representation"), "generators": g("generators"), "multi_commit": g("multi_commit")},
            "stages_ms": {k: round(v * 1e3, 3) for k, v in st.items()},
            "prof": prof,
            "what": f"3 matrices x 2^{lg} non-zeros; `seconds` = SparseMatPolyEvalProof::prove (timers: src/sparse_mlpoly.rs:1522-1545), "
                    "preprocessing (SNARK::encode: dense representation, generators, multi_commit) reported apart"}


# ----------------------------------------------------------------------------- CPU arm
def oracle_pass(X, Q, seed=1):
    """One pass of the same path in the oracle (C restatement of the reference loops)."""
    from oracle import cbind as O
    from oracle import r1cs as R

    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=seed)
    rng = np.random.default_rng(seed)
    nx, nq = log2(X), log2(Q)
    tau_q, tau_x = challenges(rng, nq), challenges(rng, nx)
    ch1, ch2 = challenges(rng, nx + nq), challenges(rng, 1 + nx)
    r_abc = challenges(rng, 3)
    t0 = time.perf_counter()
    R.prove_tables(inst, 1, Q, [Q], X, [X], secs, tau_q[:0], tau_q, tau_x, ch1, r_abc, ch2)
    return time.perf_counter() - t0


def cpu_baseline_sample(args, threads):
    from oracle import cbind as O

    O.lib()
    set_oracle_threads(threads)
    X, Q = 1 << min(args.log_x, args.cpu_log_x), min(args.proofs, args.cpu_proofs)
    dt = oracle_pass(X, Q)
    return {"value": X * Q / dt, "unit": "constraints/s", "cores": threads, "kind": "port",
            "sample": f"X=2^{log2(X)} x Q={Q} ({X * Q} constraints) of the same synthetic workload, one pass, {dt:.2f} s"}


def set_oracle_threads(n):
    from oracle import cbind as O

    try:
        O.lib().omp_set_num_threads(int(n))
    except Exception:
        os.environ["OMP_NUM_THREADS"] = str(n)


def run_reference(args):
    """The reference's CPU implementation of the same path on the host cores: the OpenMP C
    restatement of its loops (oracle/; the Rust crate cannot be built here or on the GPU box:
    no cargo). Same config as the GPU arm (X = 2^20 x Q = 64 by default); a pass takes ~20 s on
    16 cores, so the number of timed passes is capped by --ref-budget-s and reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from oracle import cbind as O
    from oracle import r1cs as R

    O.lib()
    log_x = args.ref_log_x if args.ref_log_x is not None else args.log_x
    Q = args.ref_proofs if args.ref_proofs is not None else args.proofs
    X = 1 << log_x
    host_cores = os.cpu_count() or 1
    # the reference's default build is single-threaded (rayon is optional and only used in
    # commit_inner); the restatement's (q, x) loops are OpenMP-parallel. Use whichever thread
    # count is faster on this box (containers are often CPU-throttled below their core count).
    probe = {}
    for n in sorted({1, host_cores}):
        set_oracle_threads(n)
        probe[n] = oracle_pass(1 << min(16, log_x), min(8, Q))
    cores = min(probe, key=probe.get)
    set_oracle_threads(cores)
    # one synthetic batch, reused by every pass (generation is not part of the path)
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=1)
    rng = np.random.default_rng(1)
    nx, nq = log2(X), log2(Q)
    tau_q, tau_x = challenges(rng, nq), challenges(rng, nx)
    ch1, ch2 = challenges(rng, nx + nq), challenges(rng, 1 + nx)
    r_abc = challenges(rng, 3)

    def one():
        t0 = time.perf_counter()
        R.prove_tables(inst, 1, Q, [Q], X, [X], secs, tau_q[:0], tau_q, tau_x, ch1, r_abc, ch2)
        return time.perf_counter() - t0

    t_first = one()  # warm-up pass (page faults, thread pool); also sizes the run
    steps = int(max(2, min(args.steps, args.ref_budget_s // max(t_first, 1e-3))))
    warm = 1
    t = [one() for _ in range(steps)]
    dt = float(np.mean(t))
    val = X * Q / dt
    same = (log_x == args.log_x and Q == args.proofs)
    sample = (f"the whole X=2^{log_x} x Q={Q} batch ({X * Q} constraints) per step" if same else
              f"X=2^{log_x} x Q={Q} ({X * Q} constraints) per step of the X=2^{args.log_x} x Q={args.proofs} workload")
    line = {
        "impl": "reference", "metric": "sumcheck_constraints_per_sec", "value": val, "unit": "constraints/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling or "strong",
        "vs_baseline": None, "dtype": "u256 (F_q, 4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"data-parallel R1CS batch, X=2^{args.log_x} constraints x Q={args.proofs} proofs in total, P=1 instance, W=2 sections (BASELINE configs[4] shape)",
                   "constraints_per_step": X * Q,
                   "note": "the Rust crate cannot be built here (no cargo on either box); this arm times the C restatement of its loops (oracle/) on the host cores; "
                           f"timed passes capped at {steps} by --ref-budget-s {args.ref_budget_s} (a pass takes {t_first:.1f} s)",
                   "phases": "z_mat + SpMV + phase-1 rounds + ABC/Z tables + phase-2 rounds",
                   "thread_probe_s": {str(k): v for k, v in probe.items()}},
        "cpu_baseline": {"value": val, "unit": "constraints/s", "cores": cores, "kind": "port", "sample": sample, "host_cores": host_cores},
        "e2e": {"value": val, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on
    stdout. Point fd 1 at stderr for the duration of the run and keep the real stdout for the line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-x", type=int, default=20)
    ap.add_argument("--proofs", type=int, default=64, help="proofs in the batch (strong scaling) / per GPU (--scaling weak)")
    ap.add_argument("--scaling", choices=["strong", "weak"], default=None,
                    help="N > 1: strong (default) splits the batch of --proofs over the ranks; weak keeps --proofs per GPU")
    ap.add_argument("--cpu-log-x", type=int, default=18, help="bounded CPU sample inside the GPU arm's line")
    ap.add_argument("--cpu-proofs", type=int, default=16)
    ap.add_argument("--ref-log-x", type=int, default=None, help="--impl reference: default = the GPU arm's config")
    ap.add_argument("--ref-proofs", type=int, default=None)
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget of the timed passes")
    ap.add_argument("--no-full-proof", action="store_true", help="skip the whole-proof timing")
    ap.add_argument("--no-commit", action="store_true", help="skip the witness-commitment leg")
    ap.add_argument("--no-sparse", dest="with_sparse", action="store_false", help="skip the sparse-polynomial evaluation proof leg")
    ap.add_argument("--sparse-log-nnz", type=int, default=20)
    ap.add_argument("--no-witness-gen", action="store_true", help="skip the derived-witness-section leg")
    ap.add_argument("--no-general-rows", action="store_true", help="skip the leg with 3 non-unit entries per row and matrix")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded == unsharded check (C3)")
    ap.add_argument("--replicated-phase2", action="store_true", help="N > 1: run phase 2 replicated on every rank instead of sharded over y")
    ap.add_argument("--no-weak", action="store_true", help="N > 1, strong mode: skip the extra weak-scaling figure")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the end-to-end leg")
    ap.add_argument("--trace-phases", action="store_true", help="development: per-phase wall clock of one extra pass on stderr")
    args = ap.parse_args()
    protect_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
