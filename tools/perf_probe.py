"""Quick performance probe on one GPU: build a synthetic index on the device, time search + locate."""
import argparse
import json
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sahara_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=250_000_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--len", type=int, default=150)
ap.add_argument("--k", type=int, default=2)
ap.add_argument("--edit", type=int, default=1)
ap.add_argument("--gen", default="h2-k2")
ap.add_argument("--qgram", type=int, default=0)
ap.add_argument("--densify", type=int, default=0)
ap.add_argument("--text", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--rank-bench", type=int, default=1)
ap.add_argument("--max-hits", default="", help="comma list of --max_hits limits to time after the plain search (search_n, fm_ordered_kernel)")
a = ap.parse_args()

ctx = sb.Context(0)
t = time.time()
dg = ctx.synth_genome(a.genome, 42)
ctx.build_index_device(dg, [a.genome], 6, 16)
print("index build s", round(time.time() - t, 2), ctx.info(), flush=True)
if a.rank_bench:
    for chains in (1 << 18, 1 << 20, 1 << 22):
        ms, cs = ctx.rank_bench(0, chains, 64, 7)
        ops = chains * 64
        print(f"rank_bench chains={chains} {ms:.3f} ms  {ops / ms / 1e6:.2f} Gops/s  {ops * 64 / ms / 1e6:.1f} GB/s(64B/op)", flush=True)
if a.densify:
    t = time.time(); ctx.densify(a.densify); print("densify s", round(time.time() - t, 2), flush=True)
if a.text:
    t = time.time(); ctx.enable_text(True); print("enable_text s", round(time.time() - t, 2), ctx.info()["device_bytes"] / 1e9, "GB", flush=True)
if a.qgram:
    t = time.time(); ctx.build_qgram(a.qgram); print("qgram s", round(time.time() - t, 2), flush=True)
sch = sb.SearchScheme.generate(a.gen, 0, a.k, a.len, limit_to_hamming=not a.edit)
ctx.set_scheme(sch, bool(a.edit))
dq = ctx.synth_reads(dg, a.genome, a.reads, a.len, a.k, a.edit, 43)
for rep in range(a.reps):
    ctx.reset_counters()
    t = time.time()
    nc, nh = ctx.search_device(dq, 2 * a.reads, a.len)
    dt = time.time() - t
    c = ctx.counters()
    print(json.dumps(dict(rep=rep, wall_s=round(dt, 4), reads_per_s=round(a.reads / dt), cursors=nc, hits=nh,
                          nodes=c["nodes"], lf=c["lf_steps"], ms_search=round(c["ms_search"], 2), ms_locate=round(c["ms_locate"], 2),
                          ms_sort=round(c["ms_sort"], 2), Gnodes_s=round(c["nodes"] / c["ms_search"] / 1e6, 2),
                          nodes_per_read=round(c["nodes"] / a.reads, 1))), flush=True)
for n in [int(x) for x in a.max_hits.split(",") if x]:
    ctx.set_max_hits(n)
    for rep in range(2):
        ctx.reset_counters()
        t = time.time()
        nc, nh = ctx.search_device(dq, 2 * a.reads, a.len)
        dt = time.time() - t
        c = ctx.counters()
        print(json.dumps(dict(max_hits=n, rep=rep, wall_s=round(dt, 4), reads_per_s=round(a.reads / dt), cursors=nc, hits=nh, nodes=c["nodes"],
                              ms_search=round(c["ms_search"], 2), ms_locate=round(c["ms_locate"], 2), ms_sort=round(c["ms_sort"], 2),
                              nodes_per_read=round(c["nodes"] / a.reads, 1))), flush=True)
ctx.set_max_hits(0)
