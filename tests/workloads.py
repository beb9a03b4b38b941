"""Small seeded workloads for the parity tests (numpy; independent of the CUDA generators)."""
import numpy as np


def random_genome(rng, n, with_n=False):
    g = rng.integers(1, 5, size=n, dtype=np.uint8)
    if with_n and n > 200:
        for _ in range(max(1, n // 5000)):
            s = int(rng.integers(0, n - 50))
            g[s:s + int(rng.integers(1, 40))] = 5
    return g


def repetitive_genome(rng, n):
    """random DNA with planted repeats, tandem repeats and a homopolymer run"""
    g = random_genome(rng, n)
    unit = rng.integers(1, 5, size=300, dtype=np.uint8)
    for _ in range(6):
        s = int(rng.integers(0, n - 400))
        g[s:s + 300] = unit
    s = int(rng.integers(0, n - 700))
    g[s:s + 600] = np.tile(rng.integers(1, 5, size=6, dtype=np.uint8), 100)
    s = int(rng.integers(0, n - 300))
    g[s:s + 200] = 1
    return g


def revcomp(r):
    r = np.asarray(r, dtype=np.uint8)[::-1].copy()
    m = (r >= 1) & (r <= 4)
    r[m] = 5 - r[m]
    return r


def mutate(rng, window, m, n_err, edit):
    """apply n_err random errors (substitutions only unless edit) and cut / pad to length m"""
    r = list(int(x) for x in window)
    for _ in range(n_err):
        t = int(rng.integers(0, 3)) if edit else 0
        p = int(rng.integers(0, max(1, len(r))))
        if t == 0 and r:
            r[p] = 1 + (r[p] - 1 + int(rng.integers(1, 4))) % 4 if 1 <= r[p] <= 4 else int(rng.integers(1, 5))
        elif t == 1:
            r.insert(p, int(rng.integers(1, 5)))
        elif r:
            del r[p]
    while len(r) < m:
        r.append(int(rng.integers(1, 5)))
    return np.array(r[:m], dtype=np.uint8)


def sample_reads(rng, seqs, n_reads, m, k, edit, frac_random=0.1):
    """-> dense queries [2*n_reads, m]: read, reverse complement (the reference's order)"""
    out = np.zeros((2 * n_reads, m), dtype=np.uint8)
    for i in range(n_reads):
        if rng.random() < frac_random:
            r = rng.integers(1, 5, size=m, dtype=np.uint8)
        else:
            s = seqs[int(rng.integers(0, len(seqs)))]
            while s.size < m + k + 1:
                s = seqs[int(rng.integers(0, len(seqs)))]
            p = int(rng.integers(0, s.size - m - k))
            r = mutate(rng, s[p:p + m + k], m, int(rng.integers(0, k + 1)), edit)
            if rng.random() < 0.5:
                r = revcomp(r)
        out[2 * i] = r
        out[2 * i + 1] = revcomp(r)
    return out
