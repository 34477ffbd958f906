"""CPU checks of bench.py's contract: the reference arm prints exactly ONE JSON line on stdout
with the keys the driver reads, and the host-affinity helper degrades gracefully without NVML."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-log-x", "10", "--ref-proofs", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sumcheck_constraints_per_sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "constraints/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] >= 1 and d["ms_per_step"] > 0 and d["config"]["constraints_per_step"] == 2048


def test_kernel_roofline_table():
    """the per-kernel roofline arithmetic of bench.py on a synthetic profile"""
    sys.path.insert(0, ROOT)
    import bench

    prof = [{"kernel": "k_rows_rolled", "launches": 2, "total_ms": 2.0, "units": 576.0 * 1e7, "max_ms": 1, "max_units": 1},
            {"kernel": "k_msm_rows", "launches": 1, "total_ms": 100.0, "units": 1.4e9, "max_ms": 1, "max_units": 1},
            {"kernel": "k_unknown", "launches": 1, "total_ms": 1.0, "units": 1e9, "max_ms": 1, "max_units": 1}]
    out = {e["kernel"]: e for e in bench.kernel_rooflines(prof, 6550.4, 8.459e10, 9.113e12)}
    r = out["k_rows_rolled"]
    assert abs(r["achieved_GBps"] - 576.0 * 1e7 / 2e-3 / 1e9) < 1e-6 and abs(r["modmul_per_s"] - 1e8 / 2e-3) < 1
    assert r["bound"] == "int_pipe" and abs(r["frac"] - 5e10 / 8.459e10) < 1e-9
    m = out["k_msm_rows"]
    assert abs(m["point_adds_per_s"] - 1.4e10) < 1 and abs(m["frac"] - 1.4e10 * 504 / 9.113e12) < 1e-9
    assert out["k_unknown"]["bound"] == "hbm" and "int_frac" not in out["k_unknown"]


def test_ncu_traffic_reads_the_committed_summary_with_units():
    """roofline.traffic comes from the committed `ncu --set full` summary: DRAM reads + writes of the kernel's
    launch, whatever unit ncu printed them in (the commitment kernel reads gigabytes and writes megabytes)"""
    sys.path.insert(0, ROOT)
    import bench

    t, note = bench.ncu_traffic("k_rows_rolled<true>")
    assert 9.0e9 < t < 1.1e10 and "r2_ncu_full.txt" in note            # 9.66 GB algorithmic, 1.00x
    t, _ = bench.ncu_traffic("k_msm_hrows<4>")
    assert 1.5e11 < t < 1.7e11                                          # 160.1 Gbyte + 509 Mbyte
    t, _ = bench.ncu_traffic("k_rows_spmv<2, true>")
    assert 1.05e10 < t < 1.15e10                                        # the shipped one-point first round comes first in the file
    assert bench.ncu_traffic("k_no_such_kernel") is None


def test_committed_bench_lines_keep_the_contract():
    """the bench lines under profiles/ (written by bench.py on the pool's B200s) carry every key the driver
    and the roofline bookkeeping read"""
    def last_line(name):
        return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])

    d = last_line("r2_bench_plain.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "prove_time_s"):
        assert k in d, k
    assert d["metric"] == "sumcheck_constraints_per_sec" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["config"]["constraints_per_step"] == 1 << 26 and "workload" in d["config"]
    assert abs(d["value"] - d["config"]["constraints_per_step"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 4 << 30 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and 0.9 < r["traffic"] / (r["algorithmic_bytes_per_launch"] * 5.5) < 1.2   # the largest launch is 5.5 x the average
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["prove_time"]["polycommit"] + d["prove_time"]["R1CSProof::prove"] + d["prove_time"]["SparseMatPolyEvalProof::prove"] == pytest.approx(d["prove_time_s"])
    for n in (2, 4, 8):
        s_ = last_line(f"r2_scale{n}_strong.json")
        assert s_["n_gpus"] == n and s_["scaling"] == "strong" and s_["parity"] is True
        assert s_["config"]["constraints_per_step"] == 1 << 26 and s_["weak"]["value"] > s_["value"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bind_host_to_gpu_never_raises():
    sys.path.insert(0, ROOT)
    from spartan_parallel_b200 import parallel

    before = os.sched_getaffinity(0)
    info = parallel.bind_host_to_gpu(0)
    assert info["device"] == 0
    # no NVML / no GPU in the CPU container: nothing may have changed
    if info.get("cpus") is None:
        assert os.sched_getaffinity(0) == before
