#!/usr/bin/env python
"""Per-kernel totals of ONE training step from an ncu launch list of tools/train_graph_check.py (eager part):
   python tools/train_launches.py gpurun_out/X_train_launches.csv [step#]"""
import collections
import csv
import io
import re
import sys

text = open(sys.argv[1]).read()
rows = list(csv.reader(io.StringIO(text[text.find('"ID"'):])))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(L) if "edge_rows" in k]      # three per step (one per layer)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
step = L[idx[3 * n]:idx[3 * n + 3]]
agg = collections.OrderedDict()
for k, t in step:
    k = re.sub(r"\(.*$", "", k.replace("void ", "")).replace("gg::", "")[:78]
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += t
tot = sum(t for _, t in step)
print("kernel,launches,total_us,share_pct")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.1f' % (k, c, t, 100 * t / tot))
print('"TOTAL",%d,%.1f,100.0' % (len(step), tot))
