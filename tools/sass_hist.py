#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of the built library (runs here, no GPU): proves which kernels carry tcgen05
(UTC*MMA / UTCBAR / LDTM / STTM), bulk-async copies (UBLKCP / UTMA*) and mbarrier waits (SYNCS).

  python tools/sass_hist.py [lib.so] > profiles/rNN_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "grid-gcn_b200", "libgridgcn_b200.so")
KEY = re.compile(r"^(UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|REDUX|ATOM|RED|HMMA|IMMA|FFMA|LDG|STG|LDS|STS|BAR|MEMBAR|NANOSLEEP|SHFL|VOTE|MATCH|POPC|ISETP|FSETP|FMNMX)")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        hist[cur]["_total"] += 1
        op = m.group(1)
        if KEY.match(op):
            full = op + m.group(2)
            hist[cur][full if op.startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "SYNCS", "REDUX")) else op] += 1
for fn, h in hist.items():
    name = re.sub(r"\(.*$", "", demangle(fn)).replace("void ", "").replace("gg::", "")
    tot = h.pop("_total", 0)
    tc = sum(v for k, v in h.items() if k.startswith(("UTC", "LDTM", "STTM")))
    print("%s  [%d instr%s]" % (name, tot, ", tcgen05" if tc else ""))
    print("    " + "  ".join("%s=%d" % kv for kv in sorted(h.items(), key=lambda kv: (-kv[1], kv[0]))))
