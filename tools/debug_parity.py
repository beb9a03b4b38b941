import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O, workloads as W
import sahara_b200 as sb
from test_gpu_parity import make_case

ctx = sb.Context(0)
for i, (kind, sigma) in enumerate([("random", 6), ("multi", 6), ("repeats", 6), ("random", 5)]):
    rng, seqs = make_case(100 + i, kind, sigma)
    ix = O.OracleIndex.build(seqs, sigma, 16)
    ix.save("/tmp/dbg.idx")
    ctx.load_index("/tmp/dbg.idx")
    m = 48
    for edit, k in [(True, 1), (True, 2), (True, 3), (False, 2)]:
        q = W.sample_reads(rng, seqs, 300, m, k, edit)
        for gen in ("h2-k2", "pigeon"):
            sch = sb.SearchScheme.generate(gen, 0, k, m, limit_to_hamming=not edit)
            ctx.set_scheme(sch, edit)
            want = O.sort_rows(ix.search(q, sch, edit))
            got = ctx.search_cursors(q)
            ok = got.shape == want.shape and np.array_equal(got, want)
            print(kind, sigma, edit, k, gen, "OK" if ok else f"DIFF got {got.shape[0]} want {want.shape[0]}", flush=True)
            if not ok:
                ws = set(map(tuple, want.tolist())); gs = set(map(tuple, got.tolist()))
                miss = sorted(ws - gs)[:5]; extra = sorted(gs - ws)[:5]
                print("  missing", miss, "extra", extra)
                for (qid, lb, ln, e) in miss[:2] + extra[:2]:
                    print("  query", qid, "".join("$ACGTN"[x] for x in q[qid]))
                print(sch.to_columba()[:400])
