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


@pytest.mark.parametrize("variant", ["plain", "text+qgram", "densified"])
def test_cloned_index_answers_like_its_source(sb, cases, variant):
    """sb200_index_clone: a second context (the next GPU when the box has one, else the same device) gets the index and
    its derived tables GPU to GPU and must give the oracle's hits"""
    rng, seqs, ix, path = cases[("multi", 6)]
    src = sb.Context(0)
    dst = sb.Context(1 if sb.device_count() > 1 else 0)
    try:
        src.load_index(path)
        if variant == "text+qgram":
            src.enable_text(True)
            src.build_qgram(6)
        elif variant == "densified":
            src.densify(4)
        dst.clone_index_from(src)
        assert dst.info()["n_rows"] == src.info()["n_rows"] and dst.info()["device_bytes"] > 0
        src.close()  # the copy stands alone
        src = None
        m, k = 48, 2
        q = W.sample_reads(rng, seqs, 300, m, k, True)
        sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
        dst.set_scheme(sch, True)
        assert np.array_equal(dst.search(q), O.sort_rows(ix.locate(ix.search(q, sch, True))))
    finally:
        if src is not None:
            src.close()
        dst.close()


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


def test_edge_cases(sb, ctx, cases):
    rng, seqs, ix, path = cases[("multi", 6)]
    ctx.load_index(path)
    m = 20
    sch = sb.SearchScheme.generate("h2-k2", 0, 2, m)
    ctx.set_scheme(sch, True)
    with pytest.raises(sb.SaharaError, match="empty"):
        ctx.search(np.zeros((0, m), np.uint8))
    with pytest.raises(sb.SaharaError, match="does not match the expanded search scheme"):
        ctx.search(np.ones((4, m + 1), np.uint8))  # ragged / wrong length
    with pytest.raises(sb.SaharaError, match="invalid character"):
        ctx.search(np.full((2, m), 9, np.uint8))
    # queries made of N, of '$', and reads spanning a sequence boundary never crash and match the oracle
    q = np.stack([np.full(m, 5, np.uint8), np.full(m, 1, np.uint8), np.concatenate([seqs[0][-10:], seqs[1], seqs[2][:9]])[:m],
                  np.concatenate([seqs[4][-8:], [0], seqs[5][:11]]).astype(np.uint8)])
    want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
    assert np.array_equal(ctx.search(q), want)
    # a single query, an odd number of queries
    for n in (1, 3):
        qq = W.sample_reads(rng, seqs, 2, m, 2, True)[:n]
        assert np.array_equal(ctx.search(qq), O.sort_rows(ix.locate(ix.search(qq, sch, True))))


def test_k4_and_long_reads(sb, ctx, cases):
    rng, seqs, ix, path = cases[("repeats", 6)]
    ctx.load_index(path)
    for m, k, n in ((250, 3, 60), (64, 4, 40), (301, 2, 20)):
        q = W.sample_reads(rng, seqs, n, m, k, True)
        sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
        ctx.set_scheme(sch, True)
        assert np.array_equal(ctx.search(q), O.sort_rows(ix.locate(ix.search(q, sch, True))))


def test_synthetic_generators_match_numpy_mirror(sb, ctx):
    from sahara_b200 import synth
    n = 20000
    d = ctx.synth_genome(n, 42)
    g = ctx.to_host(d, n)
    assert np.array_equal(g, synth.genome(n, 42))
    for edit, k in ((True, 2), (False, 3)):
        dq = ctx.synth_reads(d, n, 64, 50, k, edit, 43, first_read=7)
        got = ctx.to_host(dq, 2 * 64 * 50).reshape(-1, 50)
        assert np.array_equal(got, synth.reads(g, 64, 50, k, edit, 43, first_read=7))
        ctx.device_free(dq)
    ctx.device_free(d)


def test_device_pipeline_and_properties_at_scale(sb, ctx):
    """20 Mbp synthetic genome built on the GPU; properties that need no oracle run:
    sampled reads are found, the hit list is sorted, and exact (0-error) Hamming hits are edit hits too
    (hits with a mismatch at a read end are not: the edit search reports those as insertion variants)."""
    n, R, m, k = 20_000_000, 20000, 100, 2
    d = ctx.synth_genome(n, 42)
    ctx.build_index_device(d, [n], 6, 16)
    dq = ctx.synth_reads(d, n, R, m, k, True, 43)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    nc, nh = ctx.search_device(dq, 2 * R, m)
    hits = ctx.fetch_hits()
    assert hits.shape[0] == nh and nh >= nc > 0
    order = np.lexsort((hits[:, 3], hits[:, 2], hits[:, 1], hits[:, 0]))
    assert np.array_equal(order, np.arange(nh))
    found = np.zeros(R, bool)
    found[(hits[:, 0] // 2).astype(np.int64)] = True
    assert found.mean() > 0.88  # 90 % of the reads are sampled from the genome, 10 % are random
    sch_h = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=True)
    ctx.set_scheme(sch_h, False)
    ctx.search_device(dq, 2 * R, m)
    ham = ctx.fetch_hits()
    edit_set = set(map(tuple, hits[:, :3].tolist()))
    exact = ham[ham[:, 3] == 0]
    assert exact.shape[0] > 1000 and all(tuple(h) in edit_set for h in exact[:2000, :3].tolist())
    # oracle on a slice of the same reads, through the downloaded index image
    view = ctx.download_view()
    try:
        oix = O.OracleIndex.from_view(view)
    finally:
        ctx.free_view(view)
    q = ctx.to_host(dq, 2 * 500 * m).reshape(-1, m)
    ctx.set_scheme(sch, True)
    assert np.array_equal(ctx.search(q), O.sort_rows(oix.locate(oix.search(q, sch, True))))
    ctx.device_free(dq)
    ctx.device_free(d)


def test_cli_index_and_search_roundtrip(sb, cases, tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sahara_b200", "sahara")
    rng, seqs, ix, path = cases[("multi", 6)]
    fa = os.path.join(tmp_path, "ref.fa")
    with open(fa, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">seq{i}\n")
            txt = "".join("$ACGTN"[c] for c in s)
            for o in range(0, len(txt), 80):
                f.write(txt[o:o + 80] + "\n")
    subprocess.check_call([exe, "index", fa], stdout=subprocess.DEVNULL)
    assert open(fa + ".idx", "rb").read() == open(path, "rb").read()  # same file as the oracle's builder
    m, k = 36, 2
    q = W.sample_reads(rng, seqs, 150, m, k, True)
    qa = os.path.join(tmp_path, "reads.fa")
    with open(qa, "w") as f:
        for i in range(0, q.shape[0], 2):
            f.write(f">r{i // 2}\n" + "".join("$ACGTN"[c] for c in q[i]) + "\n")
    out = os.path.join(tmp_path, "out.txt")
    n_rows = ix.info()["n_rows"]
    for extra, edit, dyn in (([], True, False), (["-d", "ham"], False, False), (["--dynamic_generator"], True, True),
                             (["--dynamic_generator", "-d", "ham"], False, True)):
        res = subprocess.run([exe, "search", "-q", qa, "-i", fa + ".idx", "-e", str(k), "-o", out, "--batch", "100"] + extra,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert "queries per second" in res.stdout and "number of hits" in res.stdout
        got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
        if dyn:  # search.cpp:193-195: parts sized by weighted node count, the partition is printed
            sch, part = sb.SearchScheme.generate_dynamic("h2-k2", 0, k, m, edit, 6, n_rows)
            assert "partition: [" + ", ".join(str(x) for x in part) + "]" in res.stdout
        else:
            sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=not edit)
        want = sorted((int(a), int(b), int(c)) for a, b, c, d in ix.locate(ix.search(q, sch, edit)))
        assert got == want
    # --no-reverse and an odd --limit_queries (the forward strand of the last read only)
    for extra, keep in ((["--no-reverse"], q[0::2]), (["--limit_queries", "51"], q[:51])):
        res = subprocess.run([exe, "search", "-q", qa, "-i", fa + ".idx", "-e", str(k), "-o", out, "--batch", "64"] + extra,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
        sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
        want = sorted((int(a), int(b), int(c)) for a, b, c, d in ix.locate(ix.search(np.ascontiguousarray(keep), sch, True)))
        assert got == want, extra
    # error behaviour of the reference CLI: message + exit code 1
    res = subprocess.run([exe, "search", "-q", qa, "-i", os.path.join(tmp_path, "missing.idx")], capture_output=True, text=True)
    assert res.returncode == 1 and "no valid index path" in res.stderr
    res = subprocess.run([exe, "search", "-q", qa, "-i", fa + ".idx", "-g", "nope"], capture_output=True, text=True)
    assert res.returncode == 1 and "unknown search scheme generetaror" in res.stderr


@pytest.mark.parametrize("key", [("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)])
def test_text_mode_keeps_results(sb, ctx, cases, key):
    rng, seqs, ix, path = cases[key]
    ctx.load_index(path)
    ctx.enable_text(True)
    m = 44
    for edit, k in ((False, 2), (True, 1), (True, 2), (True, 3)):
        q = W.sample_reads(rng, seqs, 200, m, k, edit)
        q[3, 5] = 0  # a query that contains the delimiter stays on the FM path
        for gen in ("h2-k2", "pigeon_opt"):
            sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
            ctx.set_scheme(sch, edit)
            want_cur = O.sort_rows(ix.search(q, sch, edit))
            assert np.array_equal(ctx.search_cursors(q), want_cur)
            assert np.array_equal(ctx.search(q), O.sort_rows(ix.locate(want_cur)))
    ctx.enable_text(False)


def test_pipelined_host_search_with_small_chunks(sb, ctx, cases):
    """sb200_search cuts the batch into chunks that overlap copies and kernels; tiny chunks exercise the
    double buffering, the growth of the pinned result buffer and the global query ids."""
    rng, seqs, ix, path = cases[("repeats", 6)]
    ctx.load_index(path)
    ctx.enable_text(True)
    ctx.build_qgram(5)
    m, k = 40, 2
    q = W.sample_reads(rng, seqs, 333, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
    for chunk in ("7", "64", "100000"):
        ctx.set_option("chunk", int(chunk))
        assert np.array_equal(ctx.search(q), want)
    ctx.set_option("chunk", 0)  # (0 = default)
    ctx.build_qgram(0)
    ctx.enable_text(False)


def test_cli_besthits_follows_search_best(sb, cases, tmp_path):
    """-m besthits: strata of exactly 0..k errors, the first stratum with a hit ends the query
    (fmc::search_ng21::search_best as called at /root/reference/src/sahara/search.cpp:233-240)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sahara_b200", "sahara")
    rng, seqs, ix, path = cases[("repeats", 6)]
    m, k = 36, 2
    q = W.sample_reads(rng, seqs, 120, m, k, True)
    qa = os.path.join(tmp_path, "reads.fa")
    with open(qa, "w") as f:
        for i in range(0, q.shape[0], 2):
            f.write(f">r{i // 2}\n" + "".join("$ACGTN"[c] for c in q[i]) + "\n")
    out = os.path.join(tmp_path, "best.txt")
    res = subprocess.run([exe, "search", "-q", qa, "-i", path, "-e", str(k), "-o", out, "-m", "besthits", "--batch", "50"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
    want = []
    active = list(range(q.shape[0]))
    for j in range(k + 1):
        if not active:
            break
        sch = sb.SearchScheme.generate("h2-k2", j, j, m)
        hits = ix.locate(ix.search(q[active], sch, True))
        found = set()
        for a, b, c, d in hits:
            assert int(d) == j
            found.add(int(a))
            want.append((active[int(a)], int(b), int(c)))
        active = [x for i, x in enumerate(active) if i not in found]
    assert got == sorted(want)
    # search_best_n (search.cpp:240): the limit applies inside the first stratum with a hit
    res = subprocess.run([exe, "search", "-q", qa, "-i", path, "-e", str(k), "-o", out, "-m", "besthits", "--max_hits", "2"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
    want = []
    active = list(range(q.shape[0]))
    for j in range(k + 1):
        if not active:
            break
        sch = sb.SearchScheme.generate("h2-k2", j, j, m)
        hits = ix.locate(ix.search(q[active], sch, True, max_hits=2))
        found = set(int(a) for a in hits[:, 0])
        want += [(active[int(a)], int(b), int(c)) for a, b, c, d in hits]
        active = [x for i, x in enumerate(active) if i not in found]
    assert got == sorted(want)
    res = subprocess.run([exe, "search", "-q", qa, "-i", path, "--max_hits", "-3"], capture_output=True, text=True)
    assert res.returncode == 1 and "max_hits" in res.stderr


@pytest.mark.parametrize("key", [("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)])
@pytest.mark.parametrize("edit,k", [(False, 0), (False, 2), (True, 1), (True, 2), (True, 3)])
def test_max_hits_matches_search_n(sb, ctx, cases, key, edit, k):
    """sb200_set_max_hits (fm_ordered_kernel) against the oracle's search_n (search_ng24::search_n as called at
    /root/reference/src/sahara/search.cpp:228,231): the first n rows of every query in recursion order, the same number
    of cursor extensions; with and without the in-text verification tables / q-gram table loaded (not used here)."""
    rng, seqs, ix, path = cases[key]
    ctx.load_index(path)
    if k == 2:
        ctx.enable_text(True)
        ctx.build_qgram(5)
    m = 44
    q = W.sample_reads(rng, seqs, 400, m, k, edit)
    q[5, 7] = 0  # a query with the delimiter
    try:
        for gen in ("h2-k2", "01*0"):
            sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
            ctx.set_scheme(sch, edit)
            for n in (1, 3, 40, 10**9):
                ctx.set_max_hits(n)
                before = int(ix.counters[0])
                want_cur = O.sort_rows(ix.search(q, sch, edit, max_hits=n))
                nodes_oracle = int(ix.counters[0]) - before
                # the plain search + the ordered walk over the queries above the limit (what runs by default)
                got_cur = ctx.search_cursors(q)
                assert got_cur.shape == want_cur.shape and np.array_equal(got_cur, want_cur)
                # the ordered walk over every query: the extensions of the reference's search_n, one by one
                ctx.set_option("ordered_only", 1)
                ctx.reset_counters()
                got_cur = ctx.search_cursors(q)
                assert got_cur.shape == want_cur.shape and np.array_equal(got_cur, want_cur)
                if k == 2:  # (started from the q-gram table: the leading extensions are skipped)
                    assert ctx.counters()["nodes"] <= nodes_oracle
                else:
                    assert ctx.counters()["nodes"] == nodes_oracle
                ctx.set_option("ordered_only", 0)
                want_hits = O.sort_rows(ix.locate(want_cur))
                assert np.array_equal(ctx.search(q), want_hits)
                per_query = np.bincount(want_hits[:, 0].astype(np.int64), minlength=q.shape[0])
                assert per_query.max() <= n
                if n == 3:  # the reads-in / compact-hits-out call (both strands made on the device)
                    reads = q[0::2]
                    both = np.empty((2 * reads.shape[0], m), np.uint8)
                    both[0::2] = reads
                    both[1::2] = np.stack([W.revcomp(r) for r in reads])
                    want2 = O.sort_rows(ix.locate(ix.search(both, sch, edit, max_hits=n)))
                    assert np.array_equal(ctx.search_reads(reads).astype(np.uint64), want2)
            ctx.set_max_hits(0)
            assert np.array_equal(ctx.search_cursors(q), O.sort_rows(ix.search(q, sch, edit)))
    finally:
        ctx.set_max_hits(0)


def test_cli_max_hits(sb, cases, tmp_path):
    """sahara search --max_hits n (src/sahara/search.cpp:91-96, 228, 231), Hamming and edit distance, 2 batches"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sahara_b200", "sahara")
    rng, seqs, ix, path = cases[("repeats", 6)]
    m, k = 36, 2
    q = W.sample_reads(rng, seqs, 120, m, k, True)
    qa = os.path.join(tmp_path, "reads.fa")
    with open(qa, "w") as f:
        for i in range(0, q.shape[0], 2):
            f.write(f">r{i // 2}\n" + "".join("$ACGTN"[c] for c in q[i]) + "\n")
    reads = q[0::2]
    both = np.empty((2 * reads.shape[0], m), np.uint8)
    both[0::2] = reads
    both[1::2] = np.stack([W.revcomp(r) for r in reads])
    out = os.path.join(tmp_path, "n.txt")
    for metric, edit in (("lev", True), ("ham", False)):
        for n in (1, 4):
            res = subprocess.run([exe, "search", "-q", qa, "-i", path, "-e", str(k), "-o", out, "-d", metric, "--max_hits", str(n),
                                  "--batch", "100"], capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
            got = sorted(tuple(int(x) for x in line.split()) for line in open(out))
            sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=not edit)
            hits = ix.locate(ix.search(both, sch, edit, max_hits=n))
            assert got == sorted((int(a), int(b), int(c)) for a, b, c, d in hits)


def test_search_reads_compact_matches_full_call(sb, ctx, cases):
    """sb200_search_reads: reads only, reverse complements made on the device, 16-byte hits."""
    rng, seqs, ix, path = cases[("multi", 6)]
    ctx.load_index(path)
    ctx.enable_text(True)
    m, k = 33, 2
    q = W.sample_reads(rng, seqs, 211, m, k, True)
    reads = np.ascontiguousarray(q[0::2])
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
    for chunk in ("10", "500000"):
        ctx.set_option("chunk", int(chunk))
        got = ctx.search_reads(reads)
        assert got.dtype == np.uint32 and np.array_equal(got.astype(np.uint64), want)
        fwd_only = ctx.search_reads(reads, with_reverse=False)
        want_fwd = O.sort_rows(ix.locate(ix.search(reads, sch, True)))
        assert np.array_equal(fwd_only.astype(np.uint64), want_fwd)
    ctx.set_option("chunk", 0)
    ctx.enable_text(False)


def test_host_buffer_call_edge_and_middle_chunks(sb, ctx, cases):
    """Large batches are cut into a short first chunk, middle pieces and a short last chunk
    (search_host_pipelined in csrc/capi.cu); the concatenated hit lists must equal the oracle's."""
    rng, seqs, ix, path = cases[("repeats", 6)]
    ctx.load_index(path)
    ctx.enable_text(True)
    m, k = 24, 1
    q = W.sample_reads(rng, seqs, 70001, m, k, True)
    reads = np.ascontiguousarray(q[0::2])
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=True)
    ctx.set_scheme(sch, False)
    want = O.sort_rows(ix.locate(ix.search(q, sch, False, 4), 4))
    for chunk in ("40000", "2000000"):
        ctx.set_option("chunk", int(chunk))
        assert np.array_equal(ctx.search_reads(reads).astype(np.uint64), want)
        assert np.array_equal(ctx.search(q), want)
    ctx.set_option("chunk", 0)
    ctx.enable_text(False)


def test_hit_sort_fused_and_pair_paths_agree(sb, ctx, cases):
    """Hits are sorted by one 64-bit key (query id above position and errors) when the bits fit, else by two
    stable pair sorts (locate_only in csrc/capi.cu): both orders must be the oracle's."""
    rng, seqs, ix, path = cases[("multi", 6)]
    ctx.load_index(path)
    m, k = 30, 2
    q = W.sample_reads(rng, seqs, 150, m, k, True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
    for fused in ("1", "0"):
        ctx.set_option("fused_sort", int(fused))
        ctx.set_option("bucket_sort", 0)  # (the global sort is the path that has the two key layouts)
        assert np.array_equal(ctx.search(q), want)
        assert np.array_equal(ctx.search_reads(np.ascontiguousarray(q[0::2])).astype(np.uint64), want)
    ctx.set_option("fused_sort", -1)
    ctx.set_option("bucket_sort", 1)


@pytest.mark.parametrize("key", [("multi", 6), ("repeats", 6)])
def test_launch_geometry_does_not_change_results(sb, ctx, cases, key):
    """The options that tune launch geometry (blocks per SM of the walk, warps and blocks of the frame pools, run rounds
    per pop) select between equivalent executions: every combination, with and without the q-gram table, reports the
    oracle's cursors and hits."""
    rng, seqs, ix, path = cases[key]
    ctx.load_index(path)
    ctx.enable_text(True)
    m, k = 40, 2
    q = W.sample_reads(rng, seqs, 400, m, k, True)
    q[7, 11] = 0
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    want_cur = O.sort_rows(ix.search(q, sch, True))
    want = O.sort_rows(ix.locate(want_cur))
    try:
        for qlen in (0, 4, 9):
            ctx.build_qgram(qlen)
            for items_blocks, pool_threads, pool_blocks, rounds in ((0, 0, 0, 0), (1, 64, 1, 1), (4, 128, 2, 3), (3, 384, 1, 6)):
                ctx.set_option("items_blocks_per_sm", items_blocks)
                ctx.set_option("pool_threads", pool_threads)
                ctx.set_option("pool_blocks_per_sm", pool_blocks)
                ctx.set_option("run_rounds", rounds)
                assert np.array_equal(ctx.search_cursors(q), want_cur), (qlen, items_blocks, pool_threads)
                assert np.array_equal(ctx.search(q), want), (qlen, items_blocks, pool_threads)
        with pytest.raises(sb.SaharaError, match="unknown option"):
            ctx.set_option("no_such_knob", 1)
    finally:
        for name in ("items_blocks_per_sm", "pool_threads", "pool_blocks_per_sm", "run_rounds"):
            ctx.set_option(name, 0)
        ctx.build_qgram(0)
        ctx.enable_text(False)


def test_bucketed_locate_sort_all_segment_sizes(sb, ctx):
    """Hits are located into per-query buckets and sorted per query (locate.cuh): a thread sorts <= 32 hits, a block
    <= 2048, a query with more falls back to the global radix sort; cursors with more than 8 rows are located by
    warp tasks.  Repeats of 40 / 300 / 3000 copies reach every one of these paths, with the sampled and the complete
    suffix array, with text positions or rows in the cursors."""
    rng = np.random.default_rng(5)
    units = [W.random_genome(rng, 50) for _ in range(3)]
    seqs = [np.concatenate([np.tile(units[0], 40), W.random_genome(rng, 3000), np.tile(units[1], 300)]),
            np.concatenate([W.random_genome(rng, 500), np.tile(units[2], 3000)])]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m, k = 30, 1
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    reads = [W.sample_reads(rng, [np.tile(u, 3)], 6, m, k, True) for u in units[:2]] + [W.sample_reads(rng, seqs, 60, m, k, True)]
    q_small = np.concatenate(reads)
    q_huge = np.concatenate([q_small, W.sample_reads(rng, [np.tile(units[2], 3)], 4, m, k, True)])
    ctx.build_index(seqs, sigma=6, sampling_rate=16)
    ctx.set_scheme(sch, True)
    try:
        for q in (q_small, q_huge):
            want = O.sort_rows(ix.locate(ix.search(q, sch, True)))
            for text in (False, True):
                ctx.enable_text(text)
                for bucket in ("1", "0"):
                    for textpos in ("1", "0"):
                        ctx.set_option("bucket_sort", int(bucket))
                        ctx.set_option("textpos", int(textpos))
                        got = ctx.search(q)
                        assert got.shape == want.shape and np.array_equal(got, want), (text, bucket, textpos)
                        got32 = ctx.search_reads(np.ascontiguousarray(q[0::2])).astype(np.uint64)
                        assert np.array_equal(got32, want), (text, bucket, textpos)
                        # the asynchronous batches over the same segment sizes, records fixed and delta coded (a query
                        # with thousands of hits: long runs of small differences, and the radix fallback inside a batch)
                        for delta in (0, 1):
                            ctx.set_option("delta_records", delta)
                            got_async = ctx.search_reads_async(np.ascontiguousarray(q[0::2]), packed4=bool(delta), batch=40)
                            assert np.array_equal(got_async, want), (text, bucket, textpos, delta)
    finally:
        ctx.set_option("bucket_sort", 1)
        ctx.set_option("textpos", 1)
        ctx.set_option("delta_records", 1)
        ctx.enable_text(False)


@pytest.mark.parametrize("m", [19, 20, 21, 24, 33])
def test_query_validation_reports_the_last_bad_symbol(sb, ctx, cases, m):
    """pack_queries_kernel / pack_reads_kernel verify the ranks while packing them 8 per word (verify_rank,
    search.cpp:118-120); the error names the offset inside the query and the query (lengths around the word size,
    both strands, unaligned query starts)."""
    rng, seqs, ix, path = cases[("random", 5)]  # dna4: rank 5 (N) is invalid
    ctx.load_index(path)
    sch = sb.SearchScheme.generate("h2-k2", 0, 1, m)
    ctx.set_scheme(sch, True)
    q = W.sample_reads(rng, seqs, 6, m, 1, True)
    for qi, pos in ((0, 0), (3, m - 1), (5, 7), (8, 8), (11, m // 2)):
        bad = q.copy()
        bad[qi, pos] = 5
        with pytest.raises(sb.SaharaError, match=f"invalid character at offset {pos} of query {qi}$"):
            ctx.search(bad)
    # the reads call: a bad symbol of read r shows up in query 2r (offset pos) and 2r+1 (offset m-1-pos); the larger one is named
    reads = np.ascontiguousarray(q[0::2])
    for r, pos in ((0, 0), (2, m - 1), (5, 9)):
        bad = reads.copy()
        bad[r, pos] = 5
        with pytest.raises(sb.SaharaError, match=f"invalid character at offset {m - 1 - pos} of query {2 * r + 1}$"):
            ctx.search_reads(bad)
    assert np.array_equal(ctx.search_reads(reads).astype(np.uint64), ctx.search(q))
