"""Search-kernel time against the batch size (fixed cost of a launch vs. slope).  GPU only.
python tools/scale_probe.py [reads ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sahara_b200 as sb
n = int(os.environ.get("GENOME", 3100000000)); m = 150; k = 2
sizes = [int(x) for x in sys.argv[1:]] or [100, 1000, 10000, 50000, 166666, 500000, 1000000]
ctx = sb.Context(0)
dg = ctx.synth_genome(n, 42); ctx.build_index_device(dg, [n], 6, 16); ctx.enable_text(True)
ctx.build_qgram(int(os.environ.get("QGRAM", 15)))
ctx.set_scheme(sb.SearchScheme.generate("h2-k2", 0, k, m), True)
for R in sizes:
    dq = ctx.synth_reads(dg, n, R, m, k, True, 43, 0)
    best = None
    for rep in range(4):
        ctx.reset_counters()
        ctx.search_device(dq, 2 * R, m)
        c = ctx.counters()
        t = (c["ms_fm"], c["ms_text"], c["ms_locate"], c["ms_sort"], c["ms_search"])
        best = t if best is None or t[4] < best[4] else best
    print(f"reads {R:8d}: fm {best[0]:.3f} text {best[1]:.3f} locate {best[2]:.3f} sort {best[3]:.3f} search {best[4]:.3f} ms", flush=True)
    ctx.device_free(dq)
