#!/usr/bin/env python
"""Summarise ncu CSVs into markdown for profiles/.

    python tools/ncu_summarize.py raw  FILE_raw.csv      # `ncu -i x.ncu-rep --page raw --csv` -> per-launch table
    python tools/ncu_summarize.py list FILE_launches.csv [period]   # launch list -> per-kernel shares
    python tools/ncu_summarize.py traffic FILE_raw.csv REGEX OUT.json   # DRAM bytes / launch of the matched kernel
"""
import csv
import re
import sys
from collections import OrderedDict

RAW_COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %act"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.avg.per_second", "GHz"),
]


def short(name):
    name = re.sub(r"fs::\(anonymous namespace\)::|fs::|<unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


def to_unit(val, unit, want):
    v = float(val.replace(",", ""))
    u = unit.strip().lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "hz": 1e-9, "khz": 1e-6, "mhz": 1e-3, "ghz": 1.0, "cycle/nsecond": 1.0, "cycle/second": 1e-9}
    if want == "MB":
        return v * scale.get(u, 1.0) / 1e6
    if want in ("us", "GHz"):
        return v * scale.get(u, 1.0)
    return v


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    cols = [(m, lab) for m, lab in RAW_COLS if m in idx]
    print("| # | " + " | ".join(lab for _, lab in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for k, r in enumerate(rows[hdr + 2:]):
        if len(r) < len(names):
            continue
        out = []
        for m, lab in cols:
            v, u = r[idx[m]], units[idx[m]]
            if lab == "kernel":
                out.append("`%s`" % short(v))
            elif lab.endswith("MB"):
                out.append("%.1f" % to_unit(v, u, "MB"))
            elif lab == "us":
                out.append("%.1f" % to_unit(v, u, "us"))
            elif lab == "GHz":
                out.append("%.2f" % to_unit(v, u, "GHz"))
            elif lab in ("grid", "regs"):
                out.append(v.replace(",", ""))
            else:
                out.append("%.1f" % float(v.replace(",", "")))
        print("| %d | " % k + " | ".join(out) + " |")


def launch_list(path, period=None):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    ki, vi, ui = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        seq.append((short(r[ki]), to_unit(r[vi], r[ui], "us")))
    if period:
        seq = seq[:period]
    agg = OrderedDict()
    for k, us in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("Total: %.0f us over %d launches.\n" % (tot, len(seq)))
    print("| kernel | launches | us | share |")
    print("|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / tot))


def traffic(path, pattern, out_path):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels matching ``pattern`` in a
    `--set full` capture -> JSON that bench.py reads for `roofline.traffic` (tagged with the hash of the CUDA
    sources the capture was taken from)."""
    import hashlib
    import json
    import os
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    rd, wr, ki = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"], idx["Kernel Name"]
    tot, n, per = 0.0, 0, []
    for r in rows[hdr + 2:]:
        if len(r) < len(names) or not re.search(pattern, r[ki]):
            continue
        b = (to_unit(r[rd], units[rd], "MB") + to_unit(r[wr], units[wr], "MB")) * 1e6
        tot += b; n += 1; per.append(round(b))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    d = os.path.join(root, "faststyle_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    hk = hashlib.sha256()
    for f in ("conv3x3_tc.cu", "tc.cuh", "tc_ptx.cuh", "common.cuh"):        # bench.KERNEL_SOURCES
        hk.update(f.encode()); hk.update(open(os.path.join(d, f), "rb").read())
    out = {"kernel_regex": pattern, "launches": n, "dram_bytes_per_launch": tot / max(n, 1), "unit": "bytes/launch "
           "(dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, mean over the captured launches)",
           "file": os.path.basename(path), "csrc_sha16": h.hexdigest()[:16], "kernel_sha16": hk.hexdigest()[:16],
           "per_launch_bytes": per}
    json.dump(out, open(out_path, "w"), indent=1)
    print("%d launches of /%s/: %.1f MB per launch -> %s" % (n, pattern, out["dram_bytes_per_launch"] / 1e6, out_path))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        launch_list(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
