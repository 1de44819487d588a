#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: value, e2e, per-kernel-class device time.  python tools/show_bench.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1f ms %.3f e2e %.1f clocks %s launches %d" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["gpu_launches"]))
    r = d["roofline"]
    print("  roofline: achieved %.1f %s frac %.3f (burst %.3f sustained %.3f) share %.3f traffic %s alg bytes %s" % (
        r["achieved"], r["unit"], r["frac"], r.get("frac_of_burst") or 0, r.get("frac_of_sustained") or 0,
        r["share_of_step"], r.get("traffic"), r.get("algorithmic_bytes_per_launch")))
    tot = 0.0
    for k, v in r["by_kernel_class"].items():
        tot += v["ms_per_step"]
        print("   %-20s %.3f ms  %3d launches  %s" % (k, v["ms_per_step"], v["launches_per_step"],
                                                     "%.1f TF" % v["tflops"] if v["tflops"] else ""))
    print("   sum %.3f ms" % tot)
    for k in ("transform_fwd_b32", "train_step_b4", "slow_style_1024", "strong_scaling", "cpu_baseline", "replicas_identical"):
        if k in d:
            v = d[k]
            if isinstance(v, dict) and "roofline" in v:
                v = {kk: vv for kk, vv in v.items() if kk != "roofline"}
            print("  ", k, json.dumps(v)[:300])
