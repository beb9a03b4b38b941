import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O, workloads as W
import sahara_b200 as sb
from test_gpu_parity import make_case
ctx = sb.Context(0)
rng, seqs = make_case(102, "repeats", 6)
ix = O.OracleIndex.build(seqs, 6, 16)
ix.save("/tmp/dbg.idx")
ctx.load_index("/tmp/dbg.idx")
m = 48
for edit, k in [(True, 1), (True, 2)]:
    q = W.sample_reads(rng, seqs, 300, m, k, edit)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, edit)
    want = O.sort_rows(ix.search(q, sch, edit))
    for rep in range(3):
        got = ctx.search_cursors(q)
        ok = got.shape == want.shape and np.array_equal(got, want)
        print(edit, k, rep, "OK" if ok else "DIFF", got.shape, want.shape, flush=True)
    # one query at a time
    bad = 0
    for qi in range(0, q.shape[0]):
        w1 = O.sort_rows(ix.search(q[qi:qi+1], sch, edit))
        g1 = ctx.search_cursors(q[qi:qi+1])
        if not (g1.shape == w1.shape and np.array_equal(g1, w1)):
            bad += 1
            if bad <= 3:
                print("query", qi, "".join("$ACGTN"[x] for x in q[qi]))
                print(" want", w1.tolist()[:12]); print(" got ", g1.tolist()[:12])
    print("single-query mismatches:", bad)


if os.environ.get("TRACE_FIRST"):
    k = 1; edit = True
    rng, seqs = make_case(102, "repeats", 6)
    q = W.sample_reads(rng, seqs, 300, m, k, edit)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, edit)
    for qi in range(q.shape[0]):
        w1 = O.sort_rows(ix.search(q[qi:qi+1], sch, edit))
        g1 = ctx.search_cursors(q[qi:qi+1])
        if not (g1.shape == w1.shape and np.array_equal(g1, w1)):
            print("TRACEQ", qi, flush=True)
            np.save("gpurun_out/trace_query.npy", q[qi:qi+1])
            os.environ["SB200_DEBUG"] = "4"
            print("RESULT", ctx.search_cursors(q[qi:qi+1]).tolist(), flush=True)
            print("WANT", w1.tolist(), flush=True)
            break
