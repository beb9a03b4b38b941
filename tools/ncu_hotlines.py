"""Hot source lines of a kernel from an .ncu-rep captured with --import-source on (compiled with -lineinfo):
python tools/ncu_hotlines.py file.ncu-rep [n_lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
lines = []
for r in rows[hi + 1:]:
    if r and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[iI]), int(r[iS]), int(r[iT])))
        except ValueError:
            pass
tot = sum(x[2] for x in lines)
smp = sum(x[3] for x in lines)
thr = sum(x[4] for x in lines)
print(f"{rows[1][1] if len(rows[1]) > 1 else ''}")
print(f"warp instructions {tot}, thread instructions {thr} ({thr / tot:.1f} lanes per instruction), stall samples {smp}")
print("  line  inst%  smp%  lanes  source")
for ln, src, ins, s, t in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{ln:6d} {100 * ins / tot:6.2f} {100 * s / smp:5.2f} {t / max(ins, 1):6.1f}  {src[:120]}")
