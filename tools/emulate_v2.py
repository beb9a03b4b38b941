"""CPU emulation of the search kernel's frame/state traversal (search.cuh) on top of the oracle's rank
primitive; used to debug the traversal logic without a GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O, workloads as W
import sahara_b200 as sb

M, S, I, D = 0, 1, 2, 3

def emulate(ix, q, sch, edit, sigma):
    info = ix.info(); C = info["C"]; n_rows = info["n_rows"]
    out = []
    qlen = len(q)
    for j in range(sch.n_searches):
        pi, L, U = sch.pi[j], sch.l[j], sch.u[j]
        right_of = [bool(pi[0] < pi[1]) if i == 0 else bool(pi[i-1] < pi[i]) for i in range(qlen)]
        if L[0] > 1: continue
        stack = [(0, 0, n_rows, 0, 0, M, M, False)]
        while stack:
            lb, lbRev, ln, step, e, Li, Ri, pair = stack.pop()
            right = right_of[step]
            lo = lbRev if right else lb
            r1 = ix.all_ranks(1 if right else 0, [lo])[0]; r2 = ix.all_ranks(1 if right else 0, [lo + ln])[0]
            cnt = [int(r2[s] - r1[s]) for s in range(sigma)]
            own = [int(C[s] + r1[s]) for s in range(sigma)]
            before = [sum(cnt[:s]) for s in range(sigma)]
            base = lb if right else lbRev
            def kid(s):
                return (base + before[s], own[s]) if right else (own[s], base + before[s])
            t = 0
            while True:
                l, u = int(L[step]), int(U[step]); c = int(q[pi[step]])
                last = step + 1 == qlen
                lnext = 0 if last else int(L[step + 1])
                sameDirNext = (not last) and right_of[step + 1] == right
                matchOK = l <= e <= u; mmOK = l <= e + 1 <= u
                T = Ri if right else Li; Oi = Li if right else Ri
                otherEndOK = (not edit) or (Oi & 1) == 0
                def metaFor(ns, ne, side):
                    return (ns, ne, Li, side) if right else (ns, ne, side, Ri)
                if matchOK and cnt[c] != 0:
                    if last:
                        if otherEndOK: out.append((kid(c)[0], cnt[c], e))
                    elif lnext <= e + 1:
                        k = kid(c); ns, ne, a, b = metaFor(step + 1, e, M)
                        stack.append((k[0], k[1], cnt[c], ns, ne, a, b, False))
                if mmOK:
                    delOK = edit and T in (M, D)
                    subAlive = (not last) and lnext <= e + 2
                    asPair = delOK and subAlive and sameDirNext
                    for s in range(1, sigma):
                        live = s != c and cnt[s] != 0
                        if not live: continue
                        k = kid(s)
                        if asPair:
                            ns, ne, a, b = metaFor(step, e + 1, D); stack.append((k[0], k[1], cnt[s], ns, ne, a, b, True))
                        else:
                            if delOK:
                                ns, ne, a, b = metaFor(step, e + 1, D); stack.append((k[0], k[1], cnt[s], ns, ne, a, b, False))
                            if subAlive:
                                ns, ne, a, b = metaFor(step + 1, e + 1, S); stack.append((k[0], k[1], cnt[s], ns, ne, a, b, False))
                            if (not edit) and last: out.append((k[0], cnt[s], e + 1))
                if pair:
                    if t != 0: break
                    step += 1
                    if right: Ri = S
                    else: Li = S
                    t += 1
                    continue
                insOK = edit and mmOK and T in (M, I)
                if not insOK: break
                if last:
                    if otherEndOK: out.append((lb, ln, e + 1))
                    break
                if lnext > e + 2: break
                if not sameDirNext:
                    ns, ne, a, b = metaFor(step + 1, e + 1, I); stack.append((lb, lbRev, ln, ns, ne, a, b, False))
                    break
                step += 1; e += 1
                if right: Ri = I
                else: Li = I
                t += 1
    return sorted(out)

if __name__ == "__main__":
    from test_gpu_parity import make_case
    rng, seqs = make_case(102, "repeats", 6)
    ix = O.OracleIndex.build(seqs, 6, 16)
    m = 48
    for edit, k in [(True, 1), (True, 2)]:
        q = W.sample_reads(rng, seqs, 300, m, k, edit)
        sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
        for qi in ([386] if k == 1 else [292]):
            want = sorted((int(b), int(c), int(d)) for a, b, c, d in ix.search(q[qi:qi+1], sch, edit))
            got = emulate(ix, q[qi], sch, edit, 6)
            print(k, qi, "emulation == oracle:", want == got, len(want), len(got))
            if want != got:
                print(" want", want[:10]); print(" got ", got[:10])
