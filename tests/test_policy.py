"""The reconstructed rules of the recursion live in ONE table (include/sahara_policy.h) that the CPU oracle, the
scheme expansion and every kernel body consume.  These tests flip each switch and check that (a) oracle and kernels move
together — bit-exact under the flipped rule as well — and (b) the flip changes the result, i.e. the switch is live and
not shadowed by a hand-copied rule somewhere.  CPU part: the kernel SOURCE compiled for the host (tests/host_emu);
GPU part (-m gpu): the CUDA library through the C ABI."""
import os
import sys

import numpy as np
import pytest

import oracle as O
import workloads as W
import sahara_b200 as sb
from sahara_b200 import _native as N

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))
import emu  # noqa: E402

M, S, I, D = 1, 2, 4, 8  # bits of the info codes (SB200_INFO_BIT)


def policy(**kw):
    p = sb.default_policy()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# (name, policy overrides, needs max_hits to show) — every rule of the table in at least two settings
FLIPS = [
    ("del_after=M", dict(del_after=M), False),
    ("del_after=all", dict(del_after=M | S | I | D), False),
    ("ins_after=M", dict(ins_after=M), False),
    ("ins_after=all (no pair frames)", dict(ins_after=M | S | I | D), False),
    ("end_ok=all", dict(end_ok=M | S | I | D), False),
    ("end_ok=M", dict(end_ok=M), False),
    ("end_ok=M,I,S", dict(end_ok=M | I | S), False),
    ("sub before del", dict(child_order=1), True),
    ("ins before symbols", dict(child_order=2), True),
    ("both orders", dict(child_order=3), True),
]


@pytest.fixture(scope="module")
def case():
    rng = np.random.default_rng(4242)
    seqs = [W.repetitive_genome(rng, 12000), W.random_genome(rng, 3000, with_n=True)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    m, k = 32, 2
    q = W.sample_reads(rng, seqs, 80, m, k, True)
    q[5, 3] = 0
    return rng, seqs, ix, emu.text_tables(ix, seqs), q, m, k


@pytest.fixture(autouse=True)
def restore_defaults():
    yield
    O.set_policy(None)
    emu.set_policy(None)
    sb.set_expand_rule(0)


def test_default_table_is_what_the_header_says():
    p = sb.default_policy()
    assert (p.del_after, p.ins_after, p.end_ok, p.child_order, p.expand_lower) == (M | D, M | I, M | I, 0, 0)
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "sahara_policy.h")).read()
    assert "SB200_POLICY_DEFAULT" in text
    # no kernel body or oracle function spells a rule out by hand any more
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in ("sahara_b200/csrc/search.cuh", "oracle/sahara_oracle.cpp"):
        src = open(os.path.join(root, f)).read()
        assert "== INFO_D)" not in src and "== INFO_I)" not in src and "T == 'D'" not in src and "T == 'I'" not in src, f


@pytest.mark.parametrize("name,over,ordered", FLIPS, ids=[f[0] for f in FLIPS])
def test_oracle_and_kernel_source_move_together(case, name, over, ordered):
    rng, seqs, ix, tt, q, m, k = case
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    n = 3 if ordered else 0
    before = int(ix.counters[0])
    base = O.sort_rows(ix.search(q, sch, True, max_hits=n))
    nodes_base = int(ix.counters[0]) - before
    p = policy(**over)
    O.set_policy(p)
    emu.set_policy(p)
    before = int(ix.counters[0])
    want = O.sort_rows(ix.search(q, sch, True, max_hits=n))
    nodes_oracle = int(ix.counters[0]) - before
    # the switch is live: the reported cursors or at least the number of extensions change
    assert want.shape != base.shape or not np.array_equal(want, base) or nodes_oracle != nodes_base, "dead rule"
    for text, small in ((None, False), (tt, False), (tt, True)):
        got, nodes = emu.search(ix, q, sch, True, 0, text, small, max_hits=n)
        assert got.shape == want.shape and np.array_equal(got, want), (name, text is not None, small)
        assert nodes == nodes_oracle
    if not ordered:  # the ordered walk follows the same table
        want_n = O.sort_rows(ix.search(q, sch, True, max_hits=2))
        got, _ = emu.search(ix, q, sch, True, 0, tt, max_hits=2)
        assert np.array_equal(got, want_n)
    # Hamming distance ignores every rule of the table
    sch_h = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=True)
    got, _ = emu.search(ix, q, sch_h, False, 0, tt)
    O.set_policy(None)
    assert np.array_equal(got, O.sort_rows(ix.search(q, sch_h, False)))


def test_expand_rule_is_one_switch(case):
    rng, seqs, ix, tt, q, m, k = case
    a = sb.SearchScheme.generate("h2-k2", 0, k, m)
    sb.set_expand_rule(1)
    b = sb.SearchScheme.generate("h2-k2", 0, k, m)
    assert np.array_equal(a.pi, b.pi) and np.array_equal(a.u, b.u) and not np.array_equal(a.l, b.l)
    assert (b.l >= a.l).all()
    # the stricter rule demands a part's errors before its first character: fewer extensions, and a subset of the cursors
    # (it gives up completeness, which is why rule 0 is the one in force)
    n0 = int(ix.counters[0])
    ra = ix.search(q, a, True)
    n1 = int(ix.counters[0])
    rb = ix.search(q, b, True)
    n2 = int(ix.counters[0])
    sa, sb_ = set(map(tuple, ra.tolist())), set(map(tuple, rb.tolist()))
    assert sb_ < sa and n2 - n1 < n1 - n0
    got, _ = emu.search(ix, q, b, True, 0, tt)
    assert np.array_equal(got, O.sort_rows(rb))


@pytest.mark.gpu
@pytest.mark.parametrize("name,over,ordered", FLIPS, ids=[f[0] for f in FLIPS])
def test_oracle_and_cuda_kernels_move_together(case, name, over, ordered):
    rng, seqs, ix, tt, q, m, k = case
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    n = 3 if ordered else 0
    before = int(ix.counters[0])
    base = O.sort_rows(ix.search(q, sch, True, max_hits=n))
    nodes_base = int(ix.counters[0]) - before
    p = policy(**over)
    with sb.Context(0) as ctx:
        ctx.build_index(seqs, sigma=6, sampling_rate=16)
        ctx.set_scheme(sch, True)
        ctx.set_max_hits(n)
        assert np.array_equal(ctx.search_cursors(q), base)
        ctx.set_policy(p)
        got_p = ctx.get_policy()
        assert (got_p.del_after, got_p.ins_after, got_p.end_ok, got_p.child_order) == (p.del_after, p.ins_after, p.end_ok, p.child_order)
        O.set_policy(p)
        before = int(ix.counters[0])
        want = O.sort_rows(ix.search(q, sch, True, max_hits=n))
        nodes_oracle = int(ix.counters[0]) - before
        assert want.shape != base.shape or not np.array_equal(want, base) or nodes_oracle != nodes_base
        want_hits = O.sort_rows(ix.locate(want))
        for text in (False, True):
            ctx.enable_text(text)
            ctx.build_qgram(4 if text else 0)
            for only in ((0, 1) if ordered else (0,)):
                ctx.set_option("ordered_only", only)
                ctx.reset_counters()
                got = ctx.search_cursors(q)
                assert got.shape == want.shape and np.array_equal(got, want), (name, text, only)
                if not text and (only or not ordered):
                    assert ctx.counters()["nodes"] == nodes_oracle
                assert np.array_equal(ctx.search(q), want_hits)
        ctx.set_policy(None)
        ctx.set_option("ordered_only", 0)
        assert np.array_equal(ctx.search_cursors(q), base)
        bad = policy(del_after=99)
        with pytest.raises(sb.SaharaError, match="invalid search policy"):
            ctx.set_policy(bad)
