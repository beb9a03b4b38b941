"""One device-resident step of the headline workload for ncu (GPU only): index build, derived tables, `--warm` untimed
batches, then ONE batch whose kernels are the ones to capture (`ncu -k regex:... -s <skip> -c <n>`).  Prints the node
counts of that batch as JSON (the roofline numerators that go with the captured launches).

  python tools/profile_step.py [--genome BP] [--reads R] [--len M] [--errors K] [--warm W] [--out FILE]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sahara_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=3_100_000_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--len", type=int, default=150)
ap.add_argument("--errors", type=int, default=2)
ap.add_argument("--metric", default="lev")
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--text", type=int, default=1)
ap.add_argument("--qgram", type=int, default=-1)
ap.add_argument("--out", default="")
ap.add_argument("--opt", action="append", default=[], help="name=value for sb200_set_option (repeatable)")
a = ap.parse_args()
edit = a.metric == "lev"
ctx = sb.Context(0)
for o in a.opt:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
dg = ctx.synth_genome(a.genome, 42)
ctx.build_index_device(dg, [a.genome], 6, 16)
q = a.qgram if a.qgram >= 0 else max(0, min(15, int(math.log(max(4, ctx.info()["n_rows"]), 4))))
if a.text:
    ctx.enable_text(True)
if q:
    ctx.build_qgram(q)
ctx.set_scheme(sb.SearchScheme.generate("h2-k2", 0, a.errors, a.len, limit_to_hamming=not edit), edit)
batches = [ctx.synth_reads(dg, a.genome, a.reads, a.len, a.errors, edit, 43, b * a.reads) for b in range(a.warm + 1)]
for b in range(a.warm):
    ctx.search_device(batches[b], 2 * a.reads, a.len)
ctx.reset_counters()
nc, nh = ctx.search_device(batches[a.warm], 2 * a.reads, a.len)
c = ctx.counters()
rec = {"workload": f"{a.genome} bp, {a.reads} x {a.len} bp, k={a.errors} {a.metric}, qgram {q}, text {a.text}", "cursors": nc, "hits": nh,
       "nodes": c["nodes"], "nodes_text": c["nodes_text"], "nodes_fm": c["nodes"] - c["nodes_text"], "lf_steps": c["lf_steps"],
       "ms_search": c["ms_search"], "ms_fm": c["ms_fm"], "ms_text": c["ms_text"], "ms_locate": c["ms_locate"], "ms_sort": c["ms_sort"]}
print(json.dumps(rec))
if a.out:
    open(a.out, "w").write(json.dumps(rec) + "\n")
