"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of a bench.py run:
per-kernel time of ONE device-resident step (the launches between two consecutive fm_kernel launches of the
timed region) and their shares.   python tools/launch_summary.py launches.csv [step_index]"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 4  # 0-based fm_kernel launch that starts the step shown (3 warm-up steps first)
rows = []
with open(path) as f:
    for r in csv.reader(f):
        if len(r) > 14 and r[12] == "gpu__time_duration.sum":
            rows.append((r[4], float(r[14]) / 1e6))  # ms


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name[:86]


# a step begins at pack_queries_kernel (the launch before fm_kernel) and ends before the next pack_queries_kernel
starts = [i for i, (n, _) in enumerate(rows) if "pack_queries_kernel" in n]
fm = [i for i, (n, _) in enumerate(rows) if "fm_kernel" in n]
print(f"launches: {len(rows)}  fm_kernel launches: {len(fm)}")
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
agg = OrderedDict()
for n, ms in rows[lo:hi]:
    k = short(n)
    a = agg.setdefault(k, [0.0, 0])
    a[0] += ms
    a[1] += 1
total = sum(v[0] for v in agg.values())
print(f"step that starts at launch {lo} ({hi - lo} launches; serialised, cold-cache timings under ncu):")
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {ms:9.3f} ms  {100 * ms / total:5.1f}%  n={n:3d}  {k}")
print(f"  total {total:.3f} ms")
