"""numpy mirror of the counter-based synthetic generators in sahara_b200/csrc/synth.cuh (bit-identical).

genome(n, seed)                      -> uint8 ranks in 1..4
reads(genome, n_reads, len, k, edit, seed, first_read=0) -> uint8 [2*n_reads, len]  (read, reverse complement)
Only meant for small sizes (tests): the read generator is a Python loop."""
import numpy as np

_M = (1 << 64) - 1


def splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & _M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M
    return x ^ (x >> 31)


def rnd(seed, counter):
    return splitmix((seed * 0xD1342543DE82EF95 + counter) & _M)


def genome(n, seed):
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64)
        x = np.uint64((seed * 0xD1342543DE82EF95) & _M) + i
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (1 + (x >> np.uint64(62))).astype(np.uint8)


def reads(g, n_reads, length, k, edit, seed, first_read=0):
    n_bases = len(g)
    out = np.zeros((2 * n_reads, length), dtype=np.uint8)
    for t in range(n_reads):
        r = first_read + t
        base = r * 1024
        read = [0] * length
        if rnd(seed, base + 0) % 10 == 0:
            for i in range(length):
                read[i] = 1 + (rnd(seed, base + 256 + i) >> 62)
        else:
            ne = rnd(seed, base + 1) % (k + 1)
            tr = ["M"] * length
            draw = 16
            for x in range(ne):
                typ = rnd(seed, base + 2 + x) % 3 if edit else 0
                if typ < 2:
                    while True:
                        pos = rnd(seed, base + draw) % len(tr)
                        draw += 1
                        if tr[pos] == "M":
                            break
                    tr[pos] = "S" if typ == 0 else "I"
                else:
                    pos = rnd(seed, base + draw) % (len(tr) + 1)
                    draw += 1
                    tr.insert(pos, "D")
            ref_len = sum(1 for c in tr if c != "I")
            p = rnd(seed, base + 201) % (n_bases - ref_len + 1)
            o = 0
            for c in tr:
                if c == "M":
                    read[o] = int(g[p]); p += 1; o += 1
                elif c == "S":
                    gg = int(g[p]) - 1; p += 1
                    read[o] = 1 + (gg + 1 + rnd(seed, base + 256 + o) % 3) % 4
                    o += 1
                elif c == "I":
                    read[o] = 1 + (rnd(seed, base + 256 + o) >> 62)
                    o += 1
                else:
                    p += 1
            if rnd(seed, base + 200) & 1:
                read = [5 - c for c in reversed(read)]
        out[2 * t] = read
        out[2 * t + 1] = [5 - c for c in reversed(read)]
    return out
