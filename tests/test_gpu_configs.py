"""Parity on the SHAPES of the BASELINE.json configs (-m gpu): read lengths, error counts, distance metrics and device
options of cfg1..cfg5 at sizes the CPU oracle finishes in seconds, every index built by the ORACLE's own builder
(O.OracleIndex.build -> X.idx -> sb200_index_upload), so that neither the BWT nor the sampled suffix array the
checker searches comes from the GPU.  Full-size runs are covered by bench.py's parity samples and by
tests/test_gpu_scale.py (size-independent properties).  Reference call sites: /root/reference/src/sahara/search.cpp:221-250."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import workloads as W

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sb():
    import sahara_b200
    return sahara_b200


@pytest.fixture(scope="module")
def ctx(sb):
    c = sb.Context(0)
    yield c
    c.close()


def fast_reads(rng, genome, n_reads, m, k, edit):
    """vectorised version of workloads.sample_reads for one long sequence: 90 % sampled with e ~ U{0..k} errors
    (substitutions, and insertions / deletions when edit), 10 % random; both strands in the reference's order"""
    out = np.zeros((2 * n_reads, m), dtype=np.uint8)
    for i in range(n_reads):
        if rng.random() < 0.1:
            r = rng.integers(1, 5, size=m, dtype=np.uint8)
        else:
            p = int(rng.integers(0, genome.size - m - k))
            r = W.mutate(rng, genome[p:p + m + k], m, int(rng.integers(0, k + 1)), edit)
            if rng.random() < 0.5:
                r = W.revcomp(r)
        out[2 * i] = r
        out[2 * i + 1] = W.revcomp(r)
    return out


@pytest.fixture(scope="module")
def genome_1m(tmp_path_factory):
    rng = np.random.default_rng(4201)
    g = W.random_genome(rng, 1_000_000)
    ix = O.OracleIndex.build([g], 6, 16)
    path = os.path.join(tmp_path_factory.mktemp("cfg1"), "g1m.idx")
    ix.save(path)
    return rng, g, ix, path


@pytest.fixture(scope="module")
def genome_20m(tmp_path_factory):
    rng = np.random.default_rng(4202)
    g = W.random_genome(rng, 20_000_000)
    ix = O.OracleIndex.build([g], 6, 16)
    path = os.path.join(tmp_path_factory.mktemp("cfg2"), "g20m.idx")
    ix.save(path)
    return rng, g, ix, path


def check_all(sb, ctx, ix, q, sch, edit, threads=8, nodes=False):
    """cursors, hits (sync call, reads call, async pair with packed reads) against the oracle"""
    before = int(ix.counters[0])
    want_cur = O.sort_rows(ix.search(q, sch, edit, threads))
    nodes_oracle = int(ix.counters[0]) - before
    want = O.sort_rows(ix.locate(want_cur, threads))
    ctx.reset_counters()
    got_cur = ctx.search_cursors(q)
    assert got_cur.shape == want_cur.shape and np.array_equal(got_cur, want_cur)
    if nodes:
        assert ctx.counters()["nodes"] == nodes_oracle
    got = ctx.search(q)
    assert got.shape == want.shape and np.array_equal(got, want)
    reads = np.ascontiguousarray(q[0::2])
    assert np.array_equal(ctx.search_reads(reads).astype(np.uint64), want)
    got_async = ctx.search_reads_async(reads, packed4=True, batch=max(1, reads.shape[0] // 3 + 1))
    assert got_async.shape == want.shape and np.array_equal(got_async, want)
    return want


def test_cfg1_in_full(sb, ctx, genome_1m):
    """configs[0]: 1 Mbp random-DNA genome, 10 000 x 100 bp reads, --errors 2 edit distance — ALL hits of ALL reads,
    with the plain tables (LF-walking locate) and with in-text verification + q-gram table"""
    rng, g, ix, path = genome_1m
    m, k = 100, 2
    q = fast_reads(rng, g, 10_000, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.load_index(path)
    ctx.set_scheme(sch, True)
    want = check_all(sb, ctx, ix, q, sch, True, nodes=True)
    assert ctx.counters()["lf_steps"] > 0  # the hits above were located by walking LF to sampled rows
    found = np.zeros(10_000, bool)
    found[(want[:, 0] // 2).astype(np.int64)] = True
    assert found.mean() > 0.88
    ctx.enable_text(True)
    ctx.build_qgram(9)
    check_all(sb, ctx, ix, q, sch, True)
    ctx.build_qgram(0)
    ctx.enable_text(False)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_cfg2_shape_hamming(sb, ctx, genome_20m, k):
    """configs[1]: 150 bp reads, k = 0 / 1 / 2 Hamming distance (limitToHamming), 20 Mbp, 2 000 reads"""
    rng, g, ix, path = genome_20m
    m = 150
    q = fast_reads(rng, g, 2000, m, k, False)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=True)
    ctx.load_index(path)
    ctx.set_scheme(sch, False)
    want = check_all(sb, ctx, ix, q, sch, False, nodes=True)
    # the Hamming hit set is mathematically defined: every sampled read is found where it was taken from
    assert (np.bincount((want[:, 0] // 2).astype(np.int64), minlength=2000) > 0).mean() > 0.88
    ctx.enable_text(True)
    ctx.build_qgram(11)
    check_all(sb, ctx, ix, q, sch, False)
    ctx.build_qgram(0)
    ctx.enable_text(False)


@pytest.mark.parametrize("text,qgram", [(False, 0), (True, 0), (True, 11), (False, 11)])
def test_cfg3_cfg4_shape_edit(sb, ctx, genome_20m, text, qgram):
    """configs[2] / configs[3]: 150 bp reads, k = 2 edit distance, generator h2-k2; in-text verification and q-gram table
    on and off"""
    rng, g, ix, path = genome_20m
    m, k = 150, 2
    q = fast_reads(rng, g, 2000, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.load_index(path)
    ctx.set_scheme(sch, True)
    ctx.enable_text(text)
    ctx.build_qgram(qgram)
    try:
        check_all(sb, ctx, ix, q, sch, True, nodes=(qgram == 0))
    finally:
        ctx.build_qgram(0)
        ctx.enable_text(False)


def test_cfg5_shape_k3_lf_walk_locate(sb, ctx, genome_20m):
    """configs[4]: 250 bp reads, k = 3 edit distance, EVERY hit located — with in-text verification off, so that the hits
    are carried by the LF-walking locate kernel over the sampled suffix array (sampling rate 16 as the reference builds
    it), then with the tables on"""
    rng, g, ix, path = genome_20m
    m, k = 250, 3
    q = fast_reads(rng, g, 600, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.load_index(path)
    ctx.set_scheme(sch, True)
    ctx.reset_counters()
    want = check_all(sb, ctx, ix, q, sch, True, nodes=True)
    c = ctx.counters()
    assert c["lf_steps"] > 3 * want.shape[0]  # mean 7.5 LF steps per located row at sampling rate 16
    ctx.enable_text(True)
    ctx.build_qgram(11)
    check_all(sb, ctx, ix, q, sch, True)
    ctx.build_qgram(0)
    ctx.enable_text(False)


def test_cli_scheme_file(sb, genome_1m, tmp_path):
    """`sahara search --scheme-file F`: the scheme comes from a Columba-format table file — the format
    `sahara search_scheme --columba` dumps (/root/reference/src/sahara/search_scheme.cpp:252-276) and the way to pin the
    real upstream tables.  A hand-written non-default scheme must give the oracle's hits for THAT scheme."""
    exe = os.path.join(ROOT, "sahara_b200", "sahara")
    rng, g, ix, path = genome_1m
    m, k = 100, 2
    q = fast_reads(rng, g, 300, m, k, True)
    qa = os.path.join(tmp_path, "reads.fa")
    with open(qa, "w") as f:
        for i in range(0, q.shape[0], 2):
            f.write(f">r{i // 2}\n" + "".join("$ACGTN"[c] for c in q[i]) + "\n")
    # a 3-part pigeonhole-style scheme for k = 2, written in the file format (0-based parts)
    text = "{0,1,2} {0,0,0} {0,2,2}\n{1,2,0} {0,0,0} {0,2,2}\n{2,1,0} {0,0,0} {0,2,2}\n"
    sf = os.path.join(tmp_path, "scheme.txt")
    open(sf, "w").write(text)
    out = os.path.join(tmp_path, "out.txt")
    for extra, edit in (([], True), (["-d", "ham"], False)):
        res = subprocess.run([exe, "search", "-q", qa, "-i", path, "-e", str(k), "-o", out, "--scheme-file", sf] + extra,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
        sch = sb.SearchScheme.from_columba(text, m, limit_to_hamming=not edit)
        want = sorted((int(a), int(b), int(c)) for a, b, c, d in ix.locate(ix.search(q, sch, edit, 8), 8))
        assert got == want
    bad = os.path.join(tmp_path, "bad.txt")
    open(bad, "w").write("{0,1} {0,0} {0,1}\n")  # covers at most 1 error: not complete for -e 2
    res = subprocess.run([exe, "search", "-q", qa, "-i", path, "-e", "2", "-o", out, "--scheme-file", bad], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr.strip() != ""
