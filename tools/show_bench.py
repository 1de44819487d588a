#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: value, e2e, per-kernel-class device time.  python tools/show_bench.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1f ms %.3f e2e %.1f clocks %s launches %d" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["gpu_launches"]))
    r = d["roofline"]
    print("  roofline: achieved %.1f %s frac %.3f issue %.3f share %.3f" % (
        r["achieved"], r["unit"], r["frac"], r.get("mma_issue_frac") or 0, r["share_of_step"]))
    tot = 0.0
    for k, v in r["by_kernel_class"].items():
        tot += v["ms_per_step"]
        print("   %-20s %.3f ms  %3d launches  %s" % (k, v["ms_per_step"], v["launches_per_step"],
                                                     "%.1f TF" % v["tflops"] if v["tflops"] else ""))
    print("   sum %.3f ms;  fwd b32 %s img/s" % (tot, d.get("transform_fwd_b32_images_per_s", 0)))
