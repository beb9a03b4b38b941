"""End-to-end probe of the host-buffer paths on the headline workload (GPU only): PCIe bandwidth, the synchronous
calls, and the asynchronous submit / wait pair at several depths and input formats.  Prints one line per variant.

  python tools/e2e_probe.py [--genome BP] [--reads R] [--steps K] [--len M] [--errors E]
"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sahara_b200 as sb  # noqa: E402
from sahara_b200._native import check, cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=3_100_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--errors", type=int, default=2)
    ap.add_argument("--qgram", type=int, default=-1)
    a = ap.parse_args()
    R, m, k = a.reads, a.len, a.errors

    # PCIe
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        print(f"pcie {name}: {n * 4 / (time.perf_counter() - t) / 1e9:.1f} GB/s", flush=True)
    del h, d

    ctx = sb.Context(0)
    t0 = time.time()
    dg = ctx.synth_genome(a.genome, 42)
    ctx.build_index_device(dg, [a.genome], 6, 16)
    info = ctx.info()
    import math
    q = a.qgram if a.qgram >= 0 else max(0, min(15, int(math.log(max(4, info["n_rows"]), 4))))
    ctx.enable_text(True)
    ctx.build_qgram(q)
    print(f"index ready in {time.time() - t0:.1f} s, qgram {q}", flush=True)
    sch = sb.SearchScheme.generate("h2-k2", 0, k, m)
    ctx.set_scheme(sch, True)
    nb = a.steps + 2
    d_batches = [ctx.synth_reads(dg, a.genome, R, m, k, True, 43, b * R) for b in range(nb)]
    host_q, host_r, host_p = [], [], []
    W = (m + 7) // 8
    for b in range(nb):
        tq = torch.empty(2 * R * m, dtype=torch.uint8, pin_memory=True)
        check(cuda.sb200_copy_to_host(ctx._h, C.c_void_p(tq.data_ptr()), C.c_void_p(d_batches[b]), 2 * R * m))
        host_q.append(tq)
        tr = torch.empty(R * m, dtype=torch.uint8, pin_memory=True)
        tr.view(R, m).copy_(tq.view(R, 2, m)[:, 0, :])
        host_r.append(tr)
        tp = torch.empty(R * W, dtype=torch.int32, pin_memory=True)
        sb.pack_reads4(tr.view(R, m).numpy(), threads=8, out=tp.view(R, W).numpy().view(np.uint32))
        host_p.append(tp)

    def timed(name, fn, warm=2):
        for b in range(warm):
            fn(b)
        torch.cuda.synchronize()
        t = time.perf_counter()
        hits = 0
        for b in range(warm, nb):
            hits += fn(b)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / a.steps
        print(f"{name:46s} {dt * 1e3:8.3f} ms/step  {R / dt / 1e6:8.1f} M reads/s  hits/step {hits // a.steps}", flush=True)

    def pipelined(name, submit, depth, copy):
        def run(first, last):
            tickets, hits = [], 0
            for i in range(first, last + depth):
                if i - depth >= first:
                    res = ctx.wait_batch(tickets[i - depth - first], copy_to_host=copy)
                    hits += res.n_hits
                    if copy and res.n_hits:
                        C.cast(res.records, C.POINTER(C.c_uint8))[0]
                    ctx.release_batch(tickets[i - depth - first])
                if i < last:
                    tickets.append(submit(i))
            return hits
        run(0, 2)
        torch.cuda.synchronize()
        t = time.perf_counter()
        hits = run(2, nb)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / a.steps
        print(f"{name:46s} {dt * 1e3:8.3f} ms/step  {R / dt / 1e6:8.1f} M reads/s  hits/step {hits // a.steps}", flush=True)

    timed("device resident, sb200_search_device", lambda b: ctx.search_device(d_batches[b], 2 * R, m)[1])
    c = ctx.counters()
    print("   last call: search %.3f (fm %.3f text %.3f) locate %.3f sort %.3f ms" % (c["ms_search"], c["ms_fm"], c["ms_text"], c["ms_locate"], c["ms_sort"]))
    pipelined("device resident, submit_device depth 1", lambda i: ctx.submit_device(d_batches[i], 2 * R, m), 1, False)

    def sync_reads(b):
        p, nh = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search_reads(ctx._h, C.c_void_p(host_r[b].data_ptr()), R, m, 1, C.byref(p), C.byref(nh)))
        cuda.sb200_free(p)
        return nh.value

    pipelined("warm slot 2", lambda i: ctx.submit_reads((host_p[i].data_ptr(), R, m), packed4=True), 3, True)
    for overlap in (0, 1, 2):
        ctx.set_option("overlap", overlap)
        pipelined(f"overlap {overlap} device resident, submit_device depth 2", lambda i: ctx.submit_device(d_batches[i], 2 * R, m), 2, False)
        timed(f"overlap {overlap} host: sb200_search_reads (ranks, 16 B)", sync_reads)
        for depth in (2, 3):
            pipelined(f"overlap {overlap} host: submit_reads ranks depth {depth}", lambda i: ctx.submit_reads((host_r[i].data_ptr(), R, m)), depth, True)
            pipelined(f"overlap {overlap} host: submit_reads packed4 depth {depth}", lambda i: ctx.submit_reads((host_p[i].data_ptr(), R, m), packed4=True), depth, True)
    ctx.set_option("overlap", 2)
    # which copy is exposed?  reads from the host but hits left on the device / queries on the device but hits copied out
    pipelined("host: packed4 in, hits stay on device, depth 2", lambda i: ctx.submit_reads((host_p[i].data_ptr(), R, m), packed4=True), 2, False)
    pipelined("device queries, hits copied out, depth 2", lambda i: ctx.submit_device(d_batches[i], 2 * R, m), 2, True)
    # host timeline of the depth-2 loop
    tl = []
    tickets = []
    t00 = time.perf_counter()
    for i in range(2, nb + 2):
        if i - 2 >= 2:
            a0 = time.perf_counter()
            res = ctx.wait_batch(tickets[i - 4])
            a1 = time.perf_counter()
            ctx.release_batch(tickets[i - 4])
            tl.append(("wait", i - 2, (a0 - t00) * 1e3, (a1 - a0) * 1e3, res.ms_search, res.ms_locate + res.ms_sort))
        if i < nb:
            a0 = time.perf_counter()
            tickets.append(ctx.submit_reads((host_p[i].data_ptr(), R, m), packed4=True))
            tl.append(("submit", i, (a0 - t00) * 1e3, (time.perf_counter() - a0) * 1e3, 0, 0))
    for e in tl:
        print("   %-6s batch %2d at %8.3f ms took %7.3f ms  (kernels: search %.3f locate+sort %.3f)" % e)
    t = ctx.submit_reads((host_p[0].data_ptr(), R, m), packed4=True)
    res = ctx.wait_batch(t)
    print(f"bytes per step: h2d {res.h2d_bytes} d2h {res.d2h_bytes} record_bytes {res.record_bytes}; restarts {ctx.counters()['batch_restarts']}")
    ctx.release_batch(t)


if __name__ == "__main__":
    main()
