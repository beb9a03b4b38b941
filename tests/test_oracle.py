"""CPU tests of the oracle: golden fixtures, brute-force ground truth, index self-consistency.
The reference ships no golden vectors: these checks are implementation independent
(brute force / DP) or regression pins of the oracle itself."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest

import oracle as O
import workloads as W
import sahara_b200 as sb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_index():
    z = np.load(os.path.join(GOLD, "small_index.npz"))
    seqs = [z["seq0"], z["seq1"], z["seq2"]]
    return seqs, O.OracleIndex.build(seqs, 6, 16), z


def test_golden_index_bwt_and_file(gold_index, tmp_path):
    seqs, ix, z = gold_index
    assert np.array_equal(ix.bwt(0), z["bwt"]) and np.array_equal(ix.bwt(1), z["bwt_rev"])
    assert ix.info()["C"] == [int(x) for x in z["C"]]
    out = os.path.join(tmp_path, "x.idx")
    ix.save(out)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "small_index.idx"), "rb").read()
    again = O.OracleIndex.load(out)
    assert again.info() == ix.info() and np.array_equal(again.bwt(0), ix.bwt(0))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "case_*.npz"))))
def test_golden_cases(gold_index, path):
    seqs, ix, _ = gold_index
    c = np.load(path)
    sch = SimpleNamespace(pi=np.ascontiguousarray(c["pi"]), l=np.ascontiguousarray(c["l"]), u=np.ascontiguousarray(c["u"]))
    cur = O.sort_rows(ix.search(c["q"], sch, bool(c["edit"])))
    assert np.array_equal(cur, c["cursors"])
    hits = O.sort_rows(ix.locate(cur))
    assert np.array_equal(hits, c["hits"])
    if "bruteforce" in c.files:  # Hamming: the hit set is mathematically defined
        assert np.array_equal(np.unique(hits, axis=0), c["bruteforce"])


def naive_sa(text):
    n = len(text)
    return sorted(range(n), key=lambda i: bytes(text[i:]))


def test_index_matches_naive_suffix_array():
    rng = np.random.default_rng(5)
    seqs = [W.random_genome(rng, 300), W.repetitive_genome(rng, 900)[:700], np.array([1, 1, 1, 1], np.uint8)]
    ix = O.OracleIndex.build(seqs, 6, 4)
    text = np.concatenate([np.concatenate([s, [0]]) for s in seqs]).astype(np.uint8)
    sa = naive_sa(text)
    n = len(text)
    assert np.array_equal(ix.bwt(0), np.array([text[(p - 1) % n] for p in sa], dtype=np.uint8))
    # locate(row) == SA[row] expressed as (seqId, pos)
    starts = np.cumsum([0] + [len(s) + 1 for s in seqs])
    loc = ix.locate_rows(np.arange(n))
    for r in range(n):
        sid = int(np.searchsorted(starts, sa[r], side="right") - 1)
        assert (int(loc[r, 0]), int(loc[r, 1])) == (sid, sa[r] - int(starts[sid]))
        assert loc[r, 2] < 4
    # ranks: LF is a permutation
    C = ix.info()["C"]
    ranks = ix.all_ranks(0, np.arange(n))
    bwt = ix.bwt(0)
    lf = np.array([C[bwt[r]] + int(ranks[r, bwt[r]]) for r in range(n)])
    assert sorted(lf.tolist()) == list(range(n))


@pytest.mark.parametrize("k", [0, 1, 2, 3])
@pytest.mark.parametrize("gen", ["h2-k2", "pigeon", "01*0", "suffix", "optimum", "kianfar", "backtracking", "lam", "hato", "pex-td", "pex-td-l", "pex-bu", "pex-bu-l"])
def test_hamming_equals_brute_force_for_every_scheme(k, gen):
    if gen == "backtracking" and k > 2:
        pytest.skip("too slow")
    rng = np.random.default_rng(100 + k)
    seqs = [W.repetitive_genome(rng, 4000), W.random_genome(rng, 1500)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m = 24
    q = W.sample_reads(rng, seqs, 30, m, k, False)
    sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=True)
    hits = ix.locate(ix.search(q, sch, False))
    got = set((int(a), int(b), int(c), int(d)) for a, b, c, d in hits)
    want = set()
    for qi in range(q.shape[0]):
        for sid, s in enumerate(seqs):
            for p, e in O.bf_hamming(s, q[qi], k):
                want.add((qi, sid, int(p), int(e)))
    assert got == want


@pytest.mark.parametrize("k", [1, 2, 3])
def test_edit_hits_are_sound_and_complete(k):
    rng = np.random.default_rng(200 + k)
    seqs = [W.repetitive_genome(rng, 3000)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m = 28
    q = W.sample_reads(rng, seqs, 24, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    hits = ix.locate(ix.search(q, sch, True))
    by_q = {}
    for a, b, c, d in hits:
        by_q.setdefault(int(a), []).append((int(c), int(d)))
    for qi in range(q.shape[0]):
        best = O.bf_edit_starts(seqs[0], q[qi], k)  # min edit distance of an alignment starting at p (capped k+1)
        found = by_q.get(qi, [])
        for p, e in found:  # soundness: an alignment with <= e errors starts there
            assert best[p] <= e <= k
        # completeness: every locus within distance k is reported at a start within +-k
        starts = sorted(set(p for p, _ in found))
        for p in np.nonzero(best <= k)[0]:
            assert any(abs(int(p) - s) <= k for s in starts), (qi, int(p))


def test_search_rejects_bad_input():
    ix = O.OracleIndex.build([np.array([1, 2, 3, 4] * 10, np.uint8)], 6, 16)
    sch = sb.SearchScheme.generate("h2-k2", 0, 1, 8)
    with pytest.raises(O.OracleError):
        ix.search(np.full((2, 8), 7, np.uint8), sch, True)  # invalid rank
    with pytest.raises(O.OracleError):
        ix.search(np.zeros((0, 8), np.uint8), sch, True)  # empty query set


def test_search_n_is_a_prefix_of_the_recursion_order():
    """search_n (src/sahara/search.cpp:228,231): per query the rows of the plain search in recursion order, cut
    after max_hits rows"""
    rng = np.random.default_rng(5)
    seqs = [W.repetitive_genome(rng, 12000)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m, k = 30, 2
    q = W.sample_reads(rng, seqs, 50, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    full = ix.search(q, sch, True)  # reference order: query, search, recursion
    for n in (1, 3, 8):
        got = ix.search(q, sch, True, max_hits=n)
        for qid in range(q.shape[0]):
            f = full[full[:, 0] == qid]
            g = got[got[:, 0] == qid]
            rows = np.concatenate([np.arange(lb, lb + ln) for _, lb, ln, _ in f] or [np.zeros(0, np.uint64)])
            rows_n = np.concatenate([np.arange(lb, lb + ln) for _, lb, ln, _ in g] or [np.zeros(0, np.uint64)])
            assert len(rows_n) == min(n, len(rows))
            assert np.array_equal(rows_n, rows[: len(rows_n)])
