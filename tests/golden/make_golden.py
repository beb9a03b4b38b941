"""Generates tests/golden/*.npz — fixtures produced by the CPU oracle in this container.

PARITY UNPINNED: the reference ships no golden vectors and cannot be built here, so these
fixtures do NOT pin the oracle to real sahara output.  They pin (a) the oracle against regressions and
(b) the brute-force ground truth for Hamming distance, which is implementation independent.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
import workloads as W  # noqa: E402
import sahara_b200 as sb  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    seqs = [W.repetitive_genome(rng, 6000), W.random_genome(rng, 3000, with_n=True), W.random_genome(rng, 40)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m = 32
    cases = {}
    for name, edit, k, gen in [("ham_k2_h2", False, 2, "h2-k2"), ("ham_k3_pigeon", False, 3, "pigeon"), ("lev_k1_h2", True, 1, "h2-k2"),
                               ("lev_k2_h2", True, 2, "h2-k2"), ("lev_k2_01s0", True, 2, "01*0"), ("lev_k3_h2", True, 3, "h2-k2")]:
        q = W.sample_reads(rng, seqs, 40, m, k, edit)
        sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
        cur = O.sort_rows(ix.search(q, sch, edit))
        hits = O.sort_rows(ix.locate(cur))
        cases[name] = dict(q=q, pi=sch.pi, l=sch.l, u=sch.u, edit=edit, k=k, cursors=cur, hits=hits)
        if not edit:  # brute-force ground truth
            bf = []
            for qi in range(q.shape[0]):
                for sid, s in enumerate(seqs):
                    for p, e in O.bf_hamming(s, q[qi], k):
                        bf.append((qi, sid, int(p), int(e)))
            cases[name]["bruteforce"] = O.sort_rows(np.array(bf, dtype=np.uint64).reshape(-1, 4))
    np.savez_compressed(os.path.join(HERE, "small_index.npz"), seq0=seqs[0], seq1=seqs[1], seq2=seqs[2], bwt=ix.bwt(0), bwt_rev=ix.bwt(1),
                        C=np.array(ix.info()["C"], dtype=np.uint64))
    for name, c in cases.items():
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **c)
    # the index file itself (byte-exact layout pin)
    ix.save(os.path.join(HERE, "small_index.idx"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
