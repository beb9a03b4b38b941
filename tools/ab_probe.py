"""A/B of library variants on one GPU (GPU only): every variant directory under build/variants/ given on the command line
runs tools/profile_step.py in its own process (SB200_LIB_DIR), the kernel times of the measured batch side by side.

  python tools/ab_probe.py [--args "--genome 3100000000 --warm 2"] base path12 ...
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--args", default="--warm 2")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--opts", default="", help="';'-separated option sets, each a ','-separated list of name=value (one run per set)")
ap.add_argument("variants", nargs="+")
a = ap.parse_args()
optsets = [o for o in a.opts.split(";")] if a.opts else [""]
for v0 in a.variants:
  for oset in optsets:
    v = v0 + (" [" + oset + "]" if oset else "")
    libdir = os.path.join(ROOT, "sahara_b200") if v0 == "." else os.path.join(ROOT, "build", "variants", v0)
    env = dict(os.environ, SB200_LIB_DIR=libdir)
    extra = [x for o in oset.split(",") if o for x in ("--opt", o)]
    for rep in range(a.reps):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_step.py")] + a.args.split() + extra, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(f"{v:32s} FAILED: {p.stderr[-400:]}", flush=True)
            continue
        r = json.loads(p.stdout.strip().splitlines()[-1])
        print(f"{v:32s} text {r['ms_text']:7.3f} ms  fm {r['ms_fm']:6.3f}  search {r['ms_search']:7.3f}  locate {r['ms_locate']:6.3f}  sort {r['ms_sort']:6.3f}  "
              f"nodes {r['nodes']}  hits {r['hits']}  | {r['workload']}", flush=True)
