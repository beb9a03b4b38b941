"""Wall clock of the drop-in CLI (`sahara index`, `sahara search`) on a synthetic FASTA.  GPU only.

  python tools/cli_probe.py [--genome BP] [--reads R] [--len M] [--errors K] [--gpus N ...] [--dir D] [--keep]

Writes the genome and the reads as FASTA (numpy, vectorised), runs `sahara index` once and `sahara search` once per --gpus
value, and prints one JSON line per command with its wall clock and the phase times the CLI itself reports."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "sahara_b200", "sahara")
ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=100_000_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--len", type=int, default=150)
ap.add_argument("--errors", type=int, default=2)
ap.add_argument("--gpus", type=int, nargs="+", default=[1])
ap.add_argument("--dir", default="/tmp/cli_probe")
ap.add_argument("--keep", action="store_true")
a = ap.parse_args()
n, R, m, k, d = a.genome, a.reads, a.len, a.errors, a.dir
os.makedirs(d, exist_ok=True)
t0 = time.time()
rng = np.random.default_rng(42)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
txt = lut[rng.integers(0, 4, size=n, dtype=np.uint8)]
with open(f"{d}/ref.fa", "wb") as f:
    f.write(b">chr1\n")
    full = n // 80 * 80
    lines = np.empty((full // 80, 81), np.uint8)
    lines[:, :80] = txt[:full].reshape(-1, 80)
    lines[:, 80] = 10
    f.write(lines.tobytes())
    del lines
    if full < n:
        f.write(txt[full:].tobytes() + b"\n")
# reads: sampled from the genome, up to k substitutions each (half of the draws applied), fixed-width headers
pos = rng.integers(0, n - m, size=R)
rec = np.empty((R, 11 + m + 1), np.uint8)  # ">r%08d\n" + read + "\n"
rec[:, 0] = ord(">")
rec[:, 1] = ord("r")
ids = np.arange(R)
for c in range(8):
    rec[:, 2 + c] = ord("0") + (ids // 10 ** (7 - c)) % 10
rec[:, 10] = 10
chunk = 1 << 20
for s in range(0, R, chunk):
    e = min(R, s + chunk)
    rec[s:e, 11:11 + m] = txt[pos[s:e, None] + np.arange(m)[None, :]]
rec[:, 11 + m] = 10
for _ in range(k):
    p = rng.integers(0, m, size=R)
    on = rng.random(R) < 0.5
    rec[np.arange(R)[on], 11 + p[on]] = lut[rng.integers(0, 4, size=int(on.sum()))]
with open(f"{d}/reads.fa", "wb") as f:
    f.write(rec.tobytes())
del rec, txt
print(json.dumps({"made": f"{n} bp genome, {R} x {m} bp reads", "seconds": round(time.time() - t0, 1)}), flush=True)


def run(cmd, what):
    t = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t
    phases = {mm.group(1).strip(): float(mm.group(2)) for mm in re.finditer(r"^\s+(.+?) time:\s+([0-9.]+)s", res.stdout, re.M)}
    line = {"cmd": what, "rc": res.returncode, "wall_s": round(wall, 2), "phases_s": phases}
    if res.returncode != 0:
        line["stderr"] = res.stderr[-400:]
    return line


line = run([exe, "index", f"{d}/ref.fa"], "sahara index ref.fa")
line["index_bytes"] = os.path.getsize(f"{d}/ref.fa.idx") if os.path.exists(f"{d}/ref.fa.idx") else 0
print(json.dumps(line), flush=True)
for g in a.gpus:
    out = f"{d}/out_{g}.txt"
    line = run([exe, "search", "-q", f"{d}/reads.fa", "-i", f"{d}/ref.fa.idx", "-e", str(k), "-o", out, "--gpus", str(g)],
               f"sahara search -e {k} --gpus {g} ({n} bp, {R} x {m} bp)")
    if os.path.exists(out):
        line["output_bytes"] = os.path.getsize(out)
        line["reads_per_s_wall"] = round(R / line["wall_s"])
        if g != a.gpus[0]:
            line["same_output_as_first"] = subprocess.run(["cmp", "-s", out, f"{d}/out_{a.gpus[0]}.txt"]).returncode == 0
    print(json.dumps(line), flush=True)
if not a.keep:
    shutil.rmtree(d, ignore_errors=True)
