"""The evidence under profiles/ is reproducible from the committed raw files with tools/ (CPU only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_list_summary_matches_committed_table():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), "list",
                          os.path.join(ROOT, "profiles", "r01v_launches.csv"), "254"],
                         capture_output=True, text=True, check=True).stdout
    assert "over 254 launches" in out and "conv3x3_tc_kernel<16, 128>" in out
    committed = open(os.path.join(ROOT, "profiles", "r01v_launch_share.md")).read()
    first_row = [l for l in out.splitlines() if l.startswith("| `")][0]
    assert first_row in committed


def test_bench_lines_are_complete_json():
    """Every committed bench line of the final code carries the keys the measurement contract names."""
    for name in ("r01v_bench_default.json", "r01v_bench_2gpu.json"):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
        assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1
        assert not d["clocks"]["reasons"]
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01v_bench_default.json")).read().strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r01v_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["unit"] == d["unit"]


def test_round2_evidence_is_consistent():
    """Round-2 evidence: the launch-list table is reproducible from its raw CSV, the final bench lines carry the
    contract keys, and the committed ncu traffic capture belongs to the dominant kernel's sources as they are now."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), "list",
                          os.path.join(ROOT, "profiles", "r02z_launches.csv"), "155"],
                         capture_output=True, text=True, check=True).stdout
    assert "over 155 launches" in out
    committed = open(os.path.join(ROOT, "profiles", "r02z_launch_share.md")).read()
    assert [l for l in out.splitlines() if l.startswith("| `")][0] in committed
    for name in ("r02z_bench.json", "r02z_bench_8gpu.json", "r02ac_bench_profile_after_timed_legs.json"):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["vs_baseline"] is None
        r = d["roofline"]
        assert r["bound"] == "tensor" and 0 < r["frac"] < 1 and r["traffic"] > 0
        assert r["traffic"] < r["algorithmic_bytes_per_launch"] * 1.5      # no wasted re-reads from DRAM
        if d["n_gpus"] > 1:
            assert d["replicas_identical"] is True and "fs_dp_allreduce_adam" in d["config"]["dp_exchange"]
    sys.path.insert(0, ROOT)
    import bench
    t = bench.load_ncu_traffic()
    assert t is not None and t["launches"] == 59 and t["kernel_source_matches"], \
        "profiles/ncu_traffic.json was captured from different conv3x3_tc sources: re-run tools/ncu_capture.py"
