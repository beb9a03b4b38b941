"""Runs the BASELINE.json configs on one GPU and prints one JSON line per config (reads/s, phases, parity on a
sample against the oracle).  Usage: python tools/run_configs.py [cfg ...]   (cfg in 1 2 3 4 5)"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import sahara_b200 as sb  # noqa: E402
import oracle as O  # noqa: E402

CFG = {
    "1": dict(genome=1_000_000, reads=10_000, m=100, runs=[(2, True)]),
    "2": dict(genome=100_000_000, reads=1_000_000, m=150, runs=[(0, False), (1, False), (2, False)]),
    "3": dict(genome=250_000_000, reads=1_000_000, m=150, runs=[(2, True)]),
    "4": dict(genome=3_100_000_000, reads=1_000_000, m=150, runs=[(2, True)]),
    "5": dict(genome=3_100_000_000, reads=1_000_000, m=250, runs=[(3, True)]),
}


def main():
    which = sys.argv[1:] or ["1", "2", "3"]
    ctx = sb.Context(0)
    for name in which:
        cfg = CFG[name]
        n = cfg["genome"]
        t = time.time()
        dg = ctx.synth_genome(n, 42)
        ctx.build_index_device(dg, [n], 6, 16)
        t_build = time.time() - t
        ctx.enable_text(True)
        ctx.build_qgram(max(0, min(15, int(math.log(n + 1, 4)))))  # the CLI's choice
        view = ctx.download_view()
        try:
            oix = O.OracleIndex.from_view(view)
        finally:
            ctx.free_view(view)
        for k, edit in cfg["runs"]:
            m, R = cfg["m"], cfg["reads"]
            sch = sb.SearchScheme.generate("h2-k2", 0, k, m, limit_to_hamming=not edit)
            ctx.set_scheme(sch, edit)
            dq = ctx.synth_reads(dg, n, R, m, k, edit, 43)
            best = None
            for rep in range(3):
                ctx.reset_counters()
                t = time.time()
                nc, nh = ctx.search_device(dq, 2 * R, m)
                dt = time.time() - t
                c = ctx.counters()
                if best is None or dt < best[0]:
                    best = (dt, c, nc, nh)
            dt, c, nc, nh = best
            # parity on a sample through the host-buffer call, and the oracle's speed on it
            S = min(R, 5000)
            q = ctx.to_host(dq, 2 * S * m).reshape(-1, m)
            threads = O.max_threads()
            t = time.time()
            cur = oix.search(q, sch, edit, threads)
            want = O.sort_rows(oix.locate(cur, threads))
            t_cpu = time.time() - t
            t = time.time()
            cur1 = oix.search(q[: 2 * (S // 8)], sch, edit, 1)
            oix.locate(cur1, 1)
            t_cpu1 = time.time() - t
            got = ctx.search(q)
            print(json.dumps(dict(cfg=name, genome=n, reads=R, len=m, k=k, distance="edit" if edit else "hamming",
                                  index_build_s=round(t_build, 2), reads_per_s=round(R / dt), ms_search=round(c["ms_search"], 2),
                                  ms_fm=round(c["ms_fm"], 2), ms_text=round(c["ms_text"], 2), ms_locate=round(c["ms_locate"], 2),
                                  ms_sort=round(c["ms_sort"], 2), cursors=nc, hits=nh, nodes=c["nodes"],
                                  parity_sample_ok=bool(np.array_equal(got, want)), sample_reads=S,
                                  cpu_reads_per_s=round(S / t_cpu), cpu_threads=threads, cpu_1thread_reads_per_s=round((S // 8) / t_cpu1))),
                  flush=True)
            ctx.device_free(dq)
        ctx.device_free(dg)
        oix.close()


if __name__ == "__main__":
    main()
