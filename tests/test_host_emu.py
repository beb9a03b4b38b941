"""The CUDA search kernel source (sahara_b200/csrc/search.cuh) compiled for the host with g++ + UBSan and run
as a single thread against the oracle: checks the traversal logic (pair frames, insertion chains, chunked
output, query staging) without a GPU."""
import os
import sys

import numpy as np
import pytest

import oracle as O
import workloads as W
import sahara_b200 as sb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))
import emu  # noqa: E402


@pytest.fixture(scope="module")
def indexes():
    rng = np.random.default_rng(77)
    out = {}
    out["repeats"] = (rng, [W.repetitive_genome(rng, 20000), W.repetitive_genome(rng, 3000)], 6)
    out["multi"] = (rng, [W.random_genome(rng, int(n), with_n=True) for n in (9000, 1, 17, 64, 4096, 33)], 6)
    out["dna4"] = (rng, [W.random_genome(rng, 15000)], 5)
    res = {}
    for k, (r, s, sig) in out.items():
        ix = O.OracleIndex.build(s, sig, 16)
        res[k] = (r, s, ix, emu.text_tables(ix, s))
    return res


@pytest.mark.parametrize("key", ["repeats", "multi", "dna4"])
@pytest.mark.parametrize("edit,k", [(False, 0), (False, 2), (False, 3), (True, 1), (True, 2), (True, 3), (True, 4)])
def test_kernel_source_matches_oracle(indexes, key, edit, k):
    rng, seqs, ix, tt = indexes[key]
    m = 40
    q = W.sample_reads(rng, seqs, 60 if k < 4 else 12, m, k, edit)
    q[3, 9] = 0  # one query with the delimiter (never verified in the text)
    for gen in ("h2-k2", "pigeon_opt", "01*0"):
        sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
        before = int(ix.counters[0])
        want = O.sort_rows(ix.search(q, sch, edit))
        nodes_oracle = int(ix.counters[0]) - before
        # the walk alone (fm_roots_kernel + fm_items_kernel), the walk + text_pool_kernel (full and tiny pool)
        for text, flags, small in ((None, 0, False), (tt, 0, False), (tt, 0, True)):
            got, nodes = emu.search(ix, q, sch, edit, flags, text, small)
            assert got.shape == want.shape and np.array_equal(got, want)
            assert nodes == nodes_oracle  # every state the kernels expand is one extension of the reference recursion
        # with a q-gram jump table the leading error-free steps are skipped: same cursors, fewer nodes
        for text, flags in ((None, emu.QGRAM(3)), (None, emu.QGRAM(4)), (tt, emu.QGRAM(5))):
            got, nodes = emu.search(ix, q, sch, edit, flags, text)
            assert got.shape == want.shape and np.array_equal(got, want)
            assert nodes <= nodes_oracle


def test_debug_variants_agree(indexes):
    rng, seqs, ix, tt = indexes["repeats"]
    q = W.sample_reads(rng, seqs, 40, 36, 2, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, 2, 36)
    want = O.sort_rows(ix.search(q, sch, True))
    for flags in (1, 2, 3):  # no pair frames / no insertion chains / neither
        got, _ = emu.search(ix, q, sch, True, flags)
        assert np.array_equal(got, want)


def test_qgram_covers_whole_query(indexes):
    """q-gram as long as the query: the walk ignores the table"""
    rng, seqs, ix, tt = indexes["dna4"]
    q = W.sample_reads(rng, seqs, 50, 5, 0, False)
    sch = sb.SearchScheme.generate("h2-k2", 0, 0, 5, limit_to_hamming=True)
    want = O.sort_rows(ix.search(q, sch, False))
    for flags in (emu.QGRAM(5), emu.QGRAM(6)):
        got, _ = emu.search(ix, q, sch, False, flags)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("key", ["repeats", "multi", "dna4"])
@pytest.mark.parametrize("edit,k", [(False, 0), (False, 2), (True, 1), (True, 2), (True, 3)])
def test_ordered_walk_matches_search_n(indexes, key, edit, k):
    """fm_ordered_kernel's source against the oracle's search_n: the same cursors (cut at the limit) and the same
    number of cursor extensions, for limits below, around and above the number of hits of a query"""
    rng, seqs, ix, tt = indexes[key]
    m = 36
    q = W.sample_reads(rng, seqs, 40, m, k, edit)
    q[3, 9] = 0
    for gen in ("h2-k2", "pigeon_opt", "backtracking"):
        sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
        full = O.sort_rows(ix.search(q, sch, edit))
        for n in (1, 2, 5, 17, 10**9):
            before = int(ix.counters[0])
            want = O.sort_rows(ix.search(q, sch, edit, max_hits=n))
            nodes_oracle = int(ix.counters[0]) - before
            for text in (None, tt):  # probes only / unique cursors verified in the text
                got, nodes = emu.search(ix, q, sch, edit, text=text, max_hits=n)
                assert got.shape == want.shape and np.array_equal(got, want)
                assert nodes == nodes_oracle
                # started from a q-gram table: the same cursors, the leading extensions skipped
                got, nodes = emu.search(ix, q, sch, edit, emu.QGRAM(4), text=text, max_hits=n)
                assert got.shape == want.shape and np.array_equal(got, want)
                assert nodes <= nodes_oracle
            if n == 10**9:  # a limit nobody reaches: the plain search
                assert np.array_equal(got, full)


GENERATORS = ["h2-k2", "h2-k1", "h2-k3", "pigeon", "pigeon_opt", "suffix", "01*0", "01*0_opt", "optimum", "kianfar", "kucherov-k1",
              "kucherov-k2", "backtracking"]


@pytest.mark.parametrize("seed", range(64))
def test_ordered_walk_random_configurations(seed):
    """the CPU twin of the GPU fuzz sweep for search_n: random genomes, read lengths, error counts, generators, limits;
    the ordered walk with and without the text tables and the q-gram start against the oracle's search_n"""
    rng = np.random.default_rng(7000 + seed)
    sigma = 6 if rng.random() < 0.7 else 5
    kind = rng.integers(0, 3)
    if kind == 0:
        seqs = [W.random_genome(rng, int(rng.integers(2000, 20000)), with_n=(sigma == 6))]
    elif kind == 1:
        seqs = [W.repetitive_genome(rng, int(rng.integers(3000, 10000))) for _ in range(int(rng.integers(1, 4)))]
    else:
        seqs = [W.random_genome(rng, int(n), with_n=(sigma == 6)) for n in rng.integers(1, 2000, size=int(rng.integers(2, 20)))]
        seqs.append(W.random_genome(rng, 5000))
    ix = O.OracleIndex.build(seqs, sigma, 16)
    tt = emu.text_tables(ix, seqs)
    for it in range(3):
        edit = bool(rng.random() < 0.65)
        k = int(rng.integers(0, 5 if edit else 4))
        m = int(rng.integers(max(8, k + 3), 40 if k == 4 else 120))
        gen = str(rng.choice(GENERATORS))
        if gen == "backtracking" and (k > 2 or m > 40):
            gen = "h2-k2"
        q = W.sample_reads(rng, seqs, 12 if k >= 3 else 40, m, k, edit)
        if rng.random() < 0.3:
            q[int(rng.integers(0, q.shape[0])), int(rng.integers(0, m))] = 0  # a delimiter inside a query
        sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
        n = (1, 2, 3, 7, 50)[(seed + it) % 5]
        before = int(ix.counters[0])
        want = O.sort_rows(ix.search(q, sch, edit, max_hits=n))
        nodes_oracle = int(ix.counters[0]) - before
        for text, flags in ((None, 0), (tt, 0), (tt, emu.QGRAM(int(rng.integers(1, 7))))):
            got, nodes = emu.search(ix, q, sch, edit, flags, text=text, max_hits=n)
            assert got.shape == want.shape and np.array_equal(got, want), (seed, gen, k, m, edit, n, flags)
            assert nodes == nodes_oracle if flags == 0 else nodes <= nodes_oracle
