"""Randomised parity sweep (GPU): random genomes, read lengths, error counts, generators, alphabets and device
options (in-text verification, q-gram table, densified suffix array, --max_hits limits) against the oracle, bit-exact."""
import numpy as np
import pytest

import oracle as O
import workloads as W

pytestmark = pytest.mark.gpu

GENERATORS = ["h2-k2", "h2-k1", "h2-k3", "pigeon", "pigeon_opt", "suffix", "01*0", "01*0_opt", "optimum", "kianfar", "kucherov-k1",
              "kucherov-k2", "backtracking"]


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed):
    import sahara_b200 as sb
    rng = np.random.default_rng(9000 + seed)
    sigma = 6 if rng.random() < 0.7 else 5
    kind = rng.integers(0, 3)
    if kind == 0:
        seqs = [W.random_genome(rng, int(rng.integers(2000, 40000)), with_n=(sigma == 6))]
    elif kind == 1:
        seqs = [W.repetitive_genome(rng, int(rng.integers(3000, 20000))) for _ in range(int(rng.integers(1, 4)))]
    else:
        seqs = [W.random_genome(rng, int(n), with_n=(sigma == 6)) for n in rng.integers(1, 3000, size=int(rng.integers(2, 40)))]
        seqs.append(W.random_genome(rng, 5000))
    rate = int(rng.choice([1, 4, 16, 32]))
    ix = O.OracleIndex.build(seqs, sigma, rate)
    with sb.Context(0) as ctx:
        ctx.build_index(seqs, sigma=sigma, sampling_rate=rate)
        if rng.random() < 0.6:
            ctx.enable_text(True)
        elif rate > 1 and rng.random() < 0.5:
            ctx.densify(int(rng.choice([r for r in (1, 2, 4, 8, 16) if r <= rate])))
        if rng.random() < 0.5:
            ctx.build_qgram(int(rng.integers(1, 8)))
        for it in range(3):
            edit = bool(rng.random() < 0.65)
            k = int(rng.integers(0, 5 if edit else 4))
            m = int(rng.integers(max(8, k + 3), 40 if k == 4 else 160))
            gen = str(rng.choice(GENERATORS))
            if gen == "backtracking" and (k > 2 or m > 40):
                gen = "h2-k2"
            n_reads = 25 if k >= 3 else 120
            q = W.sample_reads(rng, seqs, n_reads, m, k, edit)
            if rng.random() < 0.3:
                q[int(rng.integers(0, q.shape[0])), int(rng.integers(0, m))] = 0  # a delimiter inside a query
            sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
            ctx.set_scheme(sch, edit)
            before = int(ix.counters[0])
            want_cur = O.sort_rows(ix.search(q, sch, edit))
            nodes = int(ix.counters[0]) - before
            ctx.reset_counters()
            got_cur = ctx.search_cursors(q)
            assert got_cur.shape == want_cur.shape and np.array_equal(got_cur, want_cur), (seed, gen, k, m, edit)
            if ctx.info()["device_bytes"] and not ctx.counters()["nodes"] > nodes:  # the q-gram table only removes nodes
                assert ctx.counters()["nodes"] <= nodes
            want_hits = O.sort_rows(ix.locate(want_cur))
            assert np.array_equal(ctx.search(q), want_hits), (seed, gen, k, m, edit)
            # the asynchronous pair (sb200_submit_reads / sb200_wait_batch): reads as ranks or 4-bit packed, batches of a
            # few reads submitted two deep, CSR records back
            reads = np.ascontiguousarray(q[0::2])
            both = np.empty_like(q)  # (the delimiter above may sit in a reverse strand only: the device derives that strand from the read)
            both[0::2] = reads
            both[1::2] = np.stack([W.revcomp(r) for r in reads])
            want_async = want_hits if np.array_equal(both, q) else O.sort_rows(ix.locate(ix.search(both, sch, edit)))
            # records delta coded or fixed, sorted through the per-query buckets or the global radix sort
            ctx.set_option("delta_records", ((seed + it) >> 1) & 1)
            ctx.set_option("bucket_sort", 0 if (seed + it) % 5 == 4 else 1)
            try:
                # reads in as ranks, 4-bit words or — when they hold A, C, G, T only — 2-bit words
                two_bit = (seed + it) % 3 == 2 and bool(np.all((reads >= 1) & (reads <= 4)))
                got_async = ctx.search_reads_async(reads, packed4=bool((seed + it) & 1), packed2=two_bit, batch=(7, 40, 1000)[(seed + it) % 3])
            finally:
                ctx.set_option("delta_records", 1)
                ctx.set_option("bucket_sort", 1)
            assert got_async.shape == want_async.shape and np.array_equal(got_async, want_async), (seed, gen, k, m, edit, "async")
            # search_n with a random limit: the first rows of every query in the reference's recursion order
            n = (1, 2, 3, 7, 50)[(seed + it) % 5]  # (not drawn from rng: the configurations stay what they were)
            ctx.set_max_hits(n)
            try:
                want_n = O.sort_rows(ix.search(q, sch, edit, max_hits=n))
                got_n = ctx.search_cursors(q)
                assert got_n.shape == want_n.shape and np.array_equal(got_n, want_n), (seed, gen, k, m, edit, n)
                assert np.array_equal(ctx.search(q), O.sort_rows(ix.locate(want_n))), (seed, gen, k, m, edit, n)
            finally:
                ctx.set_max_hits(0)
