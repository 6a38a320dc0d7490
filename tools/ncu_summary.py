#!/usr/bin/env python
"""Condenses ncu CSV exports into the small tables committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/X_launches.csv [first_step_launch last_step_launch]
      -> per-kernel totals and shares of a `--metrics gpu__time_duration.sum` launch list
  python tools/ncu_summary.py full gpurun_out/X_full_raw.csv > profiles/X_ncu_full_summary.csv
      -> selected columns of `ncu -i rep --page raw --csv`
"""
import csv
import io
import re
import sys
from collections import OrderedDict

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]


def rows_of(path):
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    return list(csv.reader(io.StringIO(text[start:])))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("gg::", "")


def launches(path, lo=None, hi=None):
    rows = rows_of(path)
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    recs = [(int(r[0]), short(r[ik]), r[ig], r[ib], float(r[iv].replace(",", "")) / 1e3) for r in rows[1:] if len(r) > iv]
    if lo is not None:
        recs = [r for r in recs if lo <= r[0] <= hi]
    tot = sum(r[4] for r in recs)
    agg = OrderedDict()
    for _, k, g, b, us in recs:
        key = (k, g, b)
        n, t = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, t + us)
    print("kernel,grid,block,launches,total_us,avg_us,share_pct")
    for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('"%s","%s","%s",%d,%.1f,%.1f,%.1f' % (k, g, b, n, t, t / n, 100 * t / tot))
    print('"TOTAL","","",%d,%.1f,,100.0' % (len(recs), tot))


def full(path):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(c) for c in KEEP if c in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in rows[2:]:
        if len(r) >= len(hdr):
            w.writerow([short(r[c]) if hdr[c] == "Kernel Name" else r[c] for c in cols])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        a = sys.argv[3:]
        launches(sys.argv[2], *(int(x) for x in a))
    else:
        full(sys.argv[2])
