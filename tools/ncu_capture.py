#!/usr/bin/env python
"""ncu evidence for profiles/ - run on the GPU box (1 GPU):   python tools/ncu_capture.py TAG [--gram] [--extra REGEX]

1. launch list of train steps (gpu__time_duration per launch; cold-cache + serialised: compare SHARES)
   -> gpurun_out/TAG_launches.csv, TAG_launch_share.md (one step = the last `period` launches)
2. `--set full` of EVERY tcgen05 conv launch of one whole step (the dominant kernel of bench.py's roofline)
   -> gpurun_out/TAG_conv.ncu-rep, TAG_conv_raw.csv, TAG_ncu_tc_kernels.md and ncu_traffic.json (what
   bench.py reports as roofline.traffic, tagged with the hash of the CUDA sources)
3. optionally the Gram kernels / any other regex.
Copy the .md / .csv / .json files you want judged into profiles/ (gpurun_out/ is scratch)."""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PY = sys.executable
STEP = os.path.join(ROOT, "tools", "step_once.py")
SUMM = os.path.join(ROOT, "tools", "ncu_summarize.py")


def run(cmd, log=None):
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    if log:
        open(log, "w").write(r.stdout + r.stderr)
    return r


def drop_report(path):
    """gpurun copies back at most 64 MiB of gpurun_out/: the raw CSV pages are the evidence, the binary report goes
    (keep it with FS_KEEP_NCU_REP=1 when the source page is wanted and the capture is small)."""
    if os.environ.get("FS_KEEP_NCU_REP") != "1" and os.path.exists(path):
        os.remove(path)


def kernel_names(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ki = rows[hdr].index("Kernel Name")
    return [r[ki] for r in rows[hdr + 1:] if len(r) > ki]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02x"
    os.makedirs(OUT, exist_ok=True)
    steps = 3
    if "--only-extra" in sys.argv:          # just the --extra / --gram captures (no launch list, no full conv capture)
        return extras(tag)
    lst = os.path.join(OUT, tag + "_launches.csv")
    r = run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "4000", "--csv",
             "--log-file", lst, PY, STEP, str(steps)], os.path.join(OUT, tag + "_launches.log"))
    log = open(os.path.join(OUT, tag + "_launches.log")).read()
    m = re.search(r"launches_per_step (\d+)", log)
    if not m:
        print(log[-2000:])
        raise SystemExit("step_once.py did not report launches_per_step")
    period = int(m.group(1))
    names = kernel_names(lst)
    total = len(names)
    setup = total - steps * period
    print("launch list: %d launches = %d setup + %d steps x %d" % (total, setup, steps, period))
    # keep only the LAST step in the committed list
    rows = list(csv.reader(open(lst)))
    hdr = next(i for i, rr in enumerate(rows) if "Kernel Name" in rr)
    with open(lst, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        for rr in rows[:hdr + 1] + rows[len(rows) - period:]:
            w.writerow(rr)
    md = run([PY, SUMM, "list", lst, str(period)]).stdout
    open(os.path.join(OUT, tag + "_launch_share.md"), "w").write(
        "# %s: ncu launch list of one training step (batch 8, 256x256)\n\n`ncu --metrics gpu__time_duration.sum "
        "--clock-control none python tools/step_once.py %d`, last step (%d launches). ncu times are serialised + "
        "cold-cache: compare SHARES with the live CUDA-event profile of the bench line.\n\n" % (tag, steps, period) + md)
    tc_setup = sum(1 for n in names[:setup] if "conv3x3_tc" in n)
    tc_step = sum(1 for n in names[total - period:] if "conv3x3_tc" in n)
    print("tcgen05 conv launches: %d in setup, %d per step" % (tc_setup, tc_step))
    rep = os.path.join(OUT, tag + "_conv")
    run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:conv3x3_tc",
         "-s", str(tc_setup + tc_step), "-c", str(tc_step), "-o", rep, "-f", PY, STEP, "2"],
        os.path.join(OUT, tag + "_conv.log"))
    raw = rep + "_raw.csv"
    open(raw, "w").write(run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"]).stdout)
    open(os.path.join(OUT, tag + "_ncu_tc_kernels.md"), "w").write(
        "# %s: `ncu --set full` of every tcgen05 conv launch of one train step (%d launches)\n\n" % (tag, tc_step)
        + run([PY, SUMM, "raw", raw]).stdout)
    print(run([PY, SUMM, "traffic", raw, "conv3x3_tc", os.path.join(OUT, "ncu_traffic.json")]).stdout)
    drop_report(rep + ".ncu-rep")
    extras(tag)


def extras(tag):
    extra = []
    if "--gram" in sys.argv:
        extra.append(("gram", "regex:gram_tc", 4, 4))
    for i, a in enumerate(sys.argv):
        if a == "--extra" and i + 1 < len(sys.argv):
            extra.append((re.sub(r"\W+", "_", sys.argv[i + 1])[:40], "regex:" + sys.argv[i + 1], 0, 40))
    for name, kreg, skip, count in extra:
        rep = os.path.join(OUT, "%s_%s" % (tag, name))
        run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", kreg, "-s", str(skip),
             "-c", str(count), "-o", rep, "-f", PY, STEP, "2"], rep + ".log")
        raw = rep + "_raw.csv"
        open(raw, "w").write(run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"]).stdout)
        open(os.path.join(OUT, "%s_ncu_%s.md" % (tag, name)), "w").write(run([PY, SUMM, "raw", raw]).stdout)
        drop_report(rep + ".ncu-rep")


if __name__ == "__main__":
    main()
