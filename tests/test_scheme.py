"""CPU tests of the host-side search-scheme module (sahara_b200/host/scheme.hpp through libsahara_host.so)."""
import numpy as np
import pytest

import sahara_b200 as sb

ALL = ["backtracking", "optimum", "01*0", "01*0_opt", "pigeon", "pigeon_opt", "suffix", "h2-k1", "h2-k2", "h2-k3", "kianfar",
       "kucherov-k1", "kucherov-k2", "lam", "hato", "pex-td", "pex-td-l", "pex-bu", "pex-bu-l"]


def test_names_match_reference_list():
    # every name of /root/reference/src/sahara/search_scheme.cpp:192
    assert sorted(sb.SearchScheme.names()) == sorted(ALL)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_generators_are_valid_and_complete(name, k):
    s = sb.SearchScheme.generate(name, 0, k)
    valid, complete, _ = s.check(0, k)
    assert valid and complete


def test_known_tables():
    s = sb.SearchScheme.generate("optimum", 0, 2)  # SeqAn optimum_search_scheme<0,2>, 0-based parts
    assert s.pi.tolist() == [[0, 1, 2, 3], [2, 1, 0, 3], [3, 2, 1, 0]]
    assert s.l.tolist() == [[0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 2]]
    assert s.u.tolist() == [[0, 0, 2, 2], [0, 1, 1, 2], [0, 1, 2, 2]]
    assert s.check(0, 2) == (True, True, True)
    assert sb.SearchScheme.generate("h2-k2", 0, 2).pi.tolist() == s.pi.tolist()  # default generator of sahara search


def test_incomplete_scheme_is_detected():
    s = sb.SearchScheme(np.array([[0, 1]], np.uint16), np.array([[0, 0]], np.uint8), np.array([[0, 1]], np.uint8))
    assert s.check(0, 1) == (True, False, False)
    bad = sb.SearchScheme(np.array([[0, 2, 1]], np.uint16), np.zeros((1, 3), np.uint8), np.ones((1, 3), np.uint8))
    assert bad.check(0, 1)[0] is False  # not connected


@pytest.mark.parametrize("length", [4, 5, 31, 100, 150, 151])
def test_expand(length):
    s = sb.SearchScheme.generate("h2-k2", 0, 2)
    e = sb.SearchScheme.generate("h2-k2", 0, 2, length)
    P = s.n_entries
    counts = [length // P + (1 if i < length % P else 0) for i in range(P)]
    starts = np.cumsum([0] + counts)
    for j in range(s.n_searches):
        assert sorted(e.pi[j].tolist()) == list(range(length))
        o = 0
        for i in range(P):
            part = int(s.pi[j][i])
            seg = e.pi[j][o:o + counts[part]].tolist()
            right = (s.pi[j][0] < s.pi[j][1]) if i == 0 else (s.pi[j][i - 1] < s.pi[j][i])
            rng_ = list(range(int(starts[part]), int(starts[part + 1])))
            assert seg == (rng_ if right else rng_[::-1])
            assert set(e.u[j][o:o + counts[part]].tolist()) == {int(s.u[j][i])}
            assert e.l[j][o + counts[part] - 1] == s.l[j][i]
            prev = int(s.l[j][i - 1]) if i else 0
            assert all(x == prev for x in e.l[j][o:o + counts[part] - 1].tolist())
            o += counts[part]


def test_limit_to_hamming_and_columba_roundtrip():
    e = sb.SearchScheme.generate("pigeon", 0, 3, 40)
    h = sb.SearchScheme.generate("pigeon", 0, 3, 40, limit_to_hamming=True)
    assert np.array_equal(h.u, np.minimum(e.u, np.arange(1, 41, dtype=np.uint8)[None, :]))
    s = sb.SearchScheme.generate("kianfar", 0, 2)
    back = sb.SearchScheme.from_columba(s.to_columba())
    assert np.array_equal(back.pi, s.pi) and np.array_equal(back.l, s.l) and np.array_equal(back.u, s.u)


def test_unknown_generator_message():
    with pytest.raises(sb.SaharaError, match="unknown search scheme generetaror"):
        sb.SearchScheme.generate("nope", 0, 1)


def test_node_counts_are_monotone_in_k():
    prev = 0
    for k in range(4):
        s = sb.SearchScheme.generate("h2-k2", 0, k, 100)
        nc, wnc = s.node_count(True, 6, 3_100_000_000)
        assert nc >= prev and wnc <= nc
        prev = nc


@pytest.mark.parametrize("name,k,edit", [("h2-k2", 2, True), ("h2-k2", 3, True), ("pigeon_opt", 2, False), ("01*0_opt", 2, True), ("optimum", 2, True)])
def test_dynamic_expansion(name, k, edit):
    """`--dynamic_generator` (search.cpp:193-195): part sizes by weighted node count — a partition of the query
    length with no empty part, never worse than the uniform expansion, still a valid expansion of the scheme."""
    m, sigma, n = 100, 6, 50_000_001
    dyn, part = sb.SearchScheme.generate_dynamic(name, 0, k, m, edit, sigma, n)
    uni = sb.SearchScheme.generate(name, 0, k, m, limit_to_hamming=not edit)
    base = sb.SearchScheme.generate(name, 0, k)
    assert sum(part) == m and min(part) >= 1 and len(part) == base.n_entries
    assert dyn.n_searches == uni.n_searches and dyn.n_entries == m
    assert dyn.node_count(edit, sigma, n)[1] <= uni.node_count(edit, sigma, n)[1] + 1e-9
    for j in range(dyn.n_searches):  # every search still visits every query position once, bounds monotone
        assert sorted(int(x) for x in dyn.pi[j]) == list(range(m))
        assert all(int(a) <= int(b) for a, b in zip(dyn.u[j][:-1], dyn.u[j][1:]))


def test_dynamic_expansion_finds_the_same_hamming_hits():
    import oracle as O
    import workloads as W
    import numpy as np
    rng = np.random.default_rng(3)
    seqs = [W.random_genome(rng, 4000)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    q = W.sample_reads(rng, seqs, 40, 30, 2, False)
    dyn, _ = sb.SearchScheme.generate_dynamic("h2-k2", 0, 2, 30, False, 6, ix.info()["n_rows"])
    uni = sb.SearchScheme.generate("h2-k2", 0, 2, 30, limit_to_hamming=True)
    a = O.sort_rows(ix.locate(ix.search(q, dyn, False)))
    b = O.sort_rows(ix.locate(ix.search(q, uni, False)))
    assert np.array_equal(np.unique(a, axis=0), np.unique(b, axis=0))  # the hit SET does not depend on the partition
