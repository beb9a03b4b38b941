"""Times the CLI phases (`sahara index`, `sahara search`) on a synthetic FASTA.  GPU only.
python tools/cli_probe.py [genome_bp] [reads]"""
import os, subprocess, sys, time
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "sahara_b200", "sahara")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
m, k = 150, 2
d = "/tmp/cli_probe"; os.makedirs(d, exist_ok=True)
rng = np.random.default_rng(42)
g = rng.integers(0, 4, size=n, dtype=np.uint8)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
txt = lut[g]
with open(f"{d}/ref.fa", "wb") as f:
    f.write(b">chr1\n")
    pad = (-n) % 80
    body = np.concatenate([txt, np.full(pad, ord("A"), np.uint8)]).reshape(-1, 80) if pad else txt.reshape(-1, 80)
    lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1)
    f.write(lines.tobytes())
pos = rng.integers(0, n - m, size=R)
idx = pos[:, None] + np.arange(m)[None, :]
reads = txt[idx]
# up to k substitutions per read
for _ in range(k):
    p = rng.integers(0, m, size=R); on = rng.random(R) < 0.5
    reads[np.arange(R)[on], p[on]] = lut[rng.integers(0, 4, size=int(on.sum()))]
hdr = np.array([f">r{i}\n".encode().ljust(12, b" ") for i in range(1)])  # (headers written in the loop below)
with open(f"{d}/reads.fa", "wb") as f:
    out = bytearray()
    for i in range(R):
        out += b">r%d\n" % i
        out += reads[i].tobytes() + b"\n"
        if len(out) > 1 << 24:
            f.write(out); out = bytearray()
    f.write(out)
for cmd in ([exe, "index", f"{d}/ref.fa"], [exe, "search", "-q", f"{d}/reads.fa", "-i", f"{d}/ref.fa.idx", "-e", str(k), "-o", f"{d}/out.txt"]):
    t = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True)
    print(" ".join(cmd[1:3]), "rc", res.returncode, "wall %.2f s" % (time.time() - t), flush=True)
    print(res.stdout[-1500:], res.stderr[-500:], flush=True)
print("output bytes", os.path.getsize(f"{d}/out.txt"))
