"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

import oracle as O
import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import sahara_b200
    return sahara_b200


@pytest.fixture(scope="module")
def ctx(sb):
    c = sb.Context(0)
    yield c
    c.close()


def make_case(seed, kind, sigma=6):
    rng = np.random.default_rng(seed)
    if kind == "random":
        seqs = [W.random_genome(rng, 120000)]
    elif kind == "multi":
        seqs = [W.random_genome(rng, int(n), with_n=(sigma == 6)) for n in (30000, 1, 17, 64, 4096, 50000, 33)]
    elif kind == "repeats":
        seqs = [W.repetitive_genome(rng, 60000), W.repetitive_genome(rng, 20000)]
    else:
        raise ValueError(kind)
    return rng, seqs


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    d = tmp_path_factory.mktemp("idx")
    out = {}
    for i, (kind, sigma) in enumerate([("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)]):
        rng, seqs = make_case(100 + i, kind, sigma)
        ix = O.OracleIndex.build(seqs, sigma, 16)
        path = os.path.join(d, f"{kind}{sigma}.idx")
        ix.save(path)
        out[(kind, sigma)] = (rng, seqs, ix, path)
    return out


@pytest.mark.parametrize("key", [("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)])
def test_rank_probe_matches_oracle(ctx, cases, key):
    rng, seqs, ix, path = cases[key]
    ctx.load_index(path)
    n = ix.info()["n_rows"]
    pos = np.concatenate([np.arange(0, min(n, 300)), np.arange(max(0, n - 300), n + 1),
                          rng.integers(0, n + 1, size=20000), np.arange(4032, 4200), np.arange(65500, min(n, 65700))])
    pos = pos[pos <= n].astype(np.uint64)
    for which in (0, 1):
        assert np.array_equal(ctx.rank_probe(which, pos), ix.all_ranks(which, pos))


@pytest.mark.parametrize("key", [("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)])
@pytest.mark.parametrize("edit,k", [(False, 0), (False, 1), (False, 2), (False, 3), (True, 1), (True, 2), (True, 3)])
def test_search_and_locate_match_oracle(sb, ctx, cases, key, edit, k):
    rng, seqs, ix, path = cases[key]
    ctx.load_index(path)
    m = 48
    q = W.sample_reads(rng, seqs, 300, m, k, edit)
    for gen in ("h2-k2", "pigeon"):
        sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
        ctx.set_scheme(sch, edit)
        want_cur = O.sort_rows(ix.search(q, sch, edit))
        got_cur = ctx.search_cursors(q)
        assert got_cur.shape == want_cur.shape
        assert np.array_equal(got_cur, want_cur)
        want_hits = O.sort_rows(ix.locate(want_cur))
        got_hits = ctx.search(q)
        assert np.array_equal(got_hits, want_hits)
        assert np.array_equal(ctx.locate(got_cur), want_hits)


@pytest.mark.parametrize("key", [("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)])
def test_gpu_built_index_is_byte_identical(ctx, cases, key, tmp_path):
    rng, seqs, ix, path = cases[key]
    ctx.build_index(seqs, sigma=key[1], sampling_rate=16)
    out = os.path.join(tmp_path, "gpu.idx")
    ctx.save_index(out)
    assert open(out, "rb").read() == open(path, "rb").read()


def test_upload_download_roundtrip(ctx, cases, tmp_path):
    rng, seqs, ix, path = cases[("multi", 6)]
    ctx.load_index(path)
    out = os.path.join(tmp_path, "rt.idx")
    ctx.save_index(out)
    assert open(out, "rb").read() == open(path, "rb").read()


@pytest.mark.parametrize("rate", [8, 4, 1])
def test_densified_locate_is_unchanged(sb, ctx, cases, rate):
    rng, seqs, ix, path = cases[("repeats", 6)]
    ctx.load_index(path)
    ctx.densify(rate)
    m, k = 40, 2
    q = W.sample_reads(rng, seqs, 200, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
    assert np.array_equal(ctx.search(q), want)


@pytest.mark.parametrize("qlen", [4, 7, 9])
def test_qgram_jump_table_keeps_results(sb, ctx, cases, qlen):
    rng, seqs, ix, path = cases[("random", 6)]
    ctx.load_index(path)
    ctx.build_qgram(qlen)
    m = 50
    for edit, k in ((False, 2), (True, 2), (True, 0)):
        q = W.sample_reads(rng, seqs, 200, m, k, edit)
        sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=not edit)
        ctx.set_scheme(sch, edit)
        assert np.array_equal(ctx.search_cursors(q), O.sort_rows(ix.search(q, sch, edit)))
    ctx.build_qgram(0)
