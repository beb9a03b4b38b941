"""Per-source-line totals of a kernel from an .ncu-rep captured with --import-source on (kernel compiled with -lineinfo):
warp instructions, lanes per instruction and stall samples aggregated over the SASS of every CUDA source line.

  python tools/ncu_lines.py file.ncu-rep <kernel regex> [n_lines]"""
import collections
import csv
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern, "--print-source-line-info"],
                     capture_output=True, text=True).stdout
if "Line No" not in out:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one table per source file ("File Path" row, "Function Name" row, header row, then one row per source line)
lines = []
fname = ""
iI = iS = iT = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        iI, iS, iT = r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
    elif r[0].isdigit() and iI is not None:
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[iI]), int(r[iS]), int(r[iT])))
        except (ValueError, IndexError):
            pass
tot = sum(x[3] for x in lines) or 1
smp = sum(x[4] for x in lines) or 1
thr = sum(x[5] for x in lines)
print(f"kernel {kern}: warp instructions {tot}, thread instructions {thr} ({thr / tot:.1f} lanes per instruction), stall samples {smp}")
print("  file:line            inst%  smp%  lanes  source")
for fn, ln, src, ins, sm, t in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{fn[:14]:>14s}:{ln:<5d} {100 * ins / tot:6.2f} {100 * sm / smp:5.2f} {t / max(ins, 1):6.1f}  {src[:120]}")
