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
