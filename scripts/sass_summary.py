"""Instruction mix of the hot kernels from the built library (no GPU needed):
  python scripts/sass_summary.py > profiles/r2_sass_summary.txt
For each kernel: SASS size, the mnemonic histogram (top 14), and the counts that matter here --
carry-chained IMAD.WIDE.U32[.X] (the integer-multiplier bound), 256-bit global accesses,
shared-memory traffic, and (absent by design) tensor / TMA instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spartan_parallel_b200", "libspgpu.so")
HOT = ["k_rows_rolled", "k_rows_spmv<2, true>", "k_z_bind_rq", "k_msm_hrows<4>", "k_msm_rows", "k_msm_wide", "k_quad_bind_eval<1>", "k_cubic_eval_rlc"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a); instruction mix of the hot kernels")
    for blk in blocks[1:]:
        mangled = blk.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        short = dem.split("(")[0].replace("void ", "").replace("spg::", "")
        if not any(short == h or short.startswith(h.split("<")[0]) and ("<" not in h or h in short.replace("(bool)1", "true").replace("(bool)0", "false")) for h in HOT):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
        hist = collections.Counter(o for o in ops)
        base = collections.Counter(o.split(".")[0] for o in ops)
        wide = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE.U32"))
        l256 = sum(v for k, v in hist.items() if k.startswith("LDG") and "256" in k)
        s256 = sum(v for k, v in hist.items() if k.startswith("STG") and "256" in k)
        tens = sum(v for k, v in hist.items() if k.startswith(("UTC", "HMMA", "LDTM", "STTM", "UTMA", "UBLKCP")))
        print(f"\n== {short}   ({len(ops)} instructions, {16 * len(ops) / 1024:.1f} KiB)")
        print("   " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)))
        print(f"   IMAD.WIDE.U32[.X] {wide} | LDG.E.256 {l256} | STG.E.256 {s256} | LDS {base.get('LDS', 0)} STS {base.get('STS', 0)} | "
              f"SHFL {base.get('SHFL', 0)} | tensor/TMA {tens}")


if __name__ == "__main__":
    main()
