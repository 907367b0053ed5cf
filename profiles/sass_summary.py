#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library: counts of the instructions that show HOW each kernel moves data.

    python profiles/sass_summary.py [parakeet_slam_b200/libparakeet_b200.so] > profiles/r2_sass_summary.md

UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier, LDGSTS = cp.async (per-thread async copy, .LTC64B = L2 fetch
limited to 64 B), LDG..LTC64B = scattered record loads with the 64-byte L2 fetch, STG.E.ENL2.256 = 256-bit full-sector
stores, VABSDIFF4 / IDP.4A = the byte-SIMD colour screen, MATCH = same-landmark ordering, REDUX/VOTE = warp collectives.
No UTMALDG / UTC*MMA / LDTM appears: the path has no tiled tensor copy and no dense contraction.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "parakeet_slam_b200", "libparakeet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"),
            ("LDGSTS .LTC64B", r"\bLDGSTS\S*LTC64B"), ("LDG .LTC64B", r"\bLDG\S*LTC64B"), ("STG .256", r"\bSTG\S*\.256"),
            ("LDG .128", r"\bLDG\.E\S*\.128"), ("VABSDIFF4", r"\bVABSDIFF4"), ("IDP.4A", r"\bIDP"), ("MATCH", r"\bMATCH"),
            ("REDUX / CREDUX", r"\bC?REDUX"), ("SHFL", r"\bSHFL"), ("DFMA+DMUL+DADD", r"\bD(FMA|MUL|ADD)\b"),
            ("MUFU", r"\bMUFU"), ("UTMALDG/UTC*MMA/LDTM", r"\b(UTMALDG|UTMASTG|UTC\w*MMA|LDTM|STTM)")]
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    ins = line.split("*/", 1)[1] if "*/" in line else line
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        kernels[cur]["total"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, ins):
                kernels[cur][name] += 1
demangled = subprocess.run(["c++filt"] + list(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# SASS summary of `%s` (sm_100a, nvcc 12.9)\n" % os.path.relpath(lib, ROOT))
print(__doc__.split("\n\n", 2)[2].strip() + "\n")
cols = [n for n, _ in PATTERNS]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for (k, c), d in zip(kernels.items(), demangled):
    name = re.sub(r"\(.*", "", d).replace("void ", "").replace("pk::", "")
    print("| `%s` | %d | " % (name, c["total"]) + " | ".join(str(c[n]) if c[n] else "" for n in cols) + " |")
