"""Full-size checks (-m gpu) on the headline configuration of BASELINE.json — a 3.1 Gbp genome, 150 bp reads, k = 2 edit
distance, generator h2-k2 — through properties that do not need the CPU oracle at that size:

  * soundness      a sample of the reported hits is verified against the genome itself: the read aligns at (seqId, pos)
                   with at most the reported number of edits (banded dynamic programming in numpy on windows copied back
                   from the device);
  * agreement      the same reads through three independent routes give the same hit list: in-text verification + q-gram
                   table (text_pool_kernel), the walk over the occurrence tables alone (fm_items_kernel, LF-walking locate),
                   and the asynchronous batches with packed reads and delta-coded records;
  * order          hits come back sorted by (queryId, seqId, pos, errors), query ids within the batch;
  * determinism    a second run of the same batch returns the identical array.

bench.py adds the bit-exact comparison with the oracle on 40 k reads of the same index (parity_samples); the shapes of all
five configs are compared with the oracle in tests/test_gpu_configs.py.  Reference call sites: /root/reference/src/sahara/search.cpp:221-250."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GENOME = 3_100_000_000
READS = 100_000
M, K = 150, 2


@pytest.fixture(scope="module")
def sb():
    import sahara_b200
    return sahara_b200


@pytest.fixture(scope="module")
def setup(sb):
    ctx = sb.Context(0)
    dg = ctx.synth_genome(GENOME, 42)
    ctx.build_index_device(dg, [GENOME], 6, 16)
    ctx.set_scheme(sb.SearchScheme.generate("h2-k2", 0, K, M), True)
    dq = ctx.synth_reads(dg, GENOME, READS, M, K, True, 43)
    q = ctx.to_host(dq, 2 * READS * M).reshape(-1, M)
    yield ctx, dg, q
    ctx.close()


def edit_distance_to_prefixes(reads, windows):
    """reads [n, m], windows [n, w] (symbols; 0 = behind the sequence end) -> [n, w + 1]: edit distance of read i to
    windows[i, :L] for every L (row m of the dynamic programme, all alignments at once)"""
    n, m = reads.shape
    w = windows.shape[1]
    cols = np.arange(w + 1, dtype=np.int32)
    prev = np.broadcast_to(cols, (n, w + 1)).copy()
    for i in range(1, m + 1):
        cur = np.empty_like(prev)
        cur[:, 0] = i
        sub = prev[:, :-1] + (windows != reads[:, i - 1:i]).astype(np.int32)
        cur[:, 1:] = np.minimum(sub, prev[:, 1:] + 1)
        # horizontal moves (a text symbol the read does not have): running minimum of cur[j] - j, plus j
        cur = np.minimum.accumulate(cur - cols, axis=1) + cols
        prev = cur
    return prev


def test_hits_are_sound_sorted_and_reproducible(setup):
    ctx, dg, q = setup
    ctx.enable_text(True)
    ctx.build_qgram(15)
    hits = ctx.search(q)
    assert hits.shape[0] > READS  # (90 % of the reads come from the genome, several alignments each)
    # order: (queryId, seqId, pos, errors) ascending, ids inside the batch
    key = [hits[:, 3], hits[:, 2], hits[:, 1], hits[:, 0]]
    assert np.array_equal(np.lexsort(key), np.arange(hits.shape[0]))
    assert hits[:, 0].max() < 2 * READS and hits[:, 1].max() == 0 and hits[:, 3].max() <= K
    # determinism
    assert np.array_equal(ctx.search(q), hits)
    # soundness of a sample: the read aligns at the reported position with at most the reported errors
    rng = np.random.default_rng(5)
    pick = rng.choice(hits.shape[0], size=1500, replace=False)
    w = M + K
    windows = np.zeros((pick.size, w), dtype=np.uint8)
    for j, h in enumerate(pick):
        pos = int(hits[h, 2])
        n = min(w, GENOME - pos)
        windows[j, :n] = ctx.to_host(dg + pos, n)
    d = edit_distance_to_prefixes(q[hits[pick, 0].astype(np.int64)].astype(np.uint8), windows)
    best = d[:, M - K:M + K + 1].min(axis=1)
    assert np.all(best <= hits[pick, 3].astype(np.int64)), "a reported hit does not align within its error count"


def test_three_routes_agree(setup, sb):
    ctx, dg, q = setup
    n = 20_000  # reads of the batch that also go through the (much slower) walk over the occurrence tables alone
    sub = np.ascontiguousarray(q[:2 * n])
    ctx.enable_text(True)
    ctx.build_qgram(15)
    with_text = ctx.search(sub)
    ctx.set_option("delta_records", 1)
    asynchronous = ctx.search_reads_async(np.ascontiguousarray(sub[0::2]), packed4=True, batch=7_000)
    assert np.array_equal(asynchronous, with_text)
    with_tables = ctx.info()["device_bytes"]
    ctx.build_qgram(0)
    ctx.enable_text(False)
    try:
        assert ctx.info()["device_bytes"] < with_tables - 40e9  # (the verification tables and the q-gram table are gone)
        walk_only = ctx.search(sub)
    finally:
        ctx.enable_text(True)
        ctx.build_qgram(15)
    assert np.array_equal(walk_only, with_text)
