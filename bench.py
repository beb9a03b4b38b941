#!/usr/bin/env python
"""bench.py — headline benchmark of the sahara_b200 hot path.

Metric (BASELINE.json): queries/s for 150 bp reads, k = 2 edit distance, on a synthetic human-sized genome
(configs[3]: 3.1 Gbp, query-sharded over 1/2/4/8 B200).  One "query" = one read searched on both strands
(SURVEY.md §8d).  A step = search + locate + sort of one batch of reads against the replicated index.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--genome BP] [--reads-per-step R]

`value`  : whole-job reads/s with the queries already in HBM and the hits left in HBM (device timed).
`e2e`    : the same through sb200_search_reads() with pinned HOST buffers: H2D of the reads (the reverse
           complements are made on the device) and D2H of the hits (16 B per hit) inside the timed region.
`e2e_full_tuples`: the same through sb200_search(): both strands from the host, 32-byte hit tuples back.
`roofline`: the search kernel; algorithmic bytes = search nodes x 2 rank-ops x 64 B over
           its CUDA-event duration, against the measured HBM bandwidth of MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (restatement of fmc::search_ng24 + LocateLinear, "port") on the host cores,
           on a bounded sample of the same reads.
`--impl reference` times only that CPU path (the real sahara cannot be built here: its search lives in
fmindex-collection, fetched at configure time).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome", type=int, default=3_100_000_000)
    ap.add_argument("--reads-per-step", type=int, default=1_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--errors", type=int, default=2)
    ap.add_argument("--metric", default="lev", choices=["lev", "ham"])
    ap.add_argument("--generator", default="h2-k2")
    ap.add_argument("--qgram", type=int, default=-1, help="q-gram jump table length (-1 = auto, 0 = off)")
    ap.add_argument("--device-sa-rate", type=int, default=0, help="densify the device suffix array (0 = keep 16)")
    ap.add_argument("--text", type=int, default=1, help="in-text verification of unique cursors (1 = on)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """per-launch DRAM bytes (read + write) of the search kernels from the committed ncu captures, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


_JSON_OUT = None


def emit_json(line):
    """the one JSON line goes to the real stdout; everything libraries print meanwhile (NCCL's version banner ...)
    was sent to stderr by main()"""
    os.write(_JSON_OUT if _JSON_OUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _JSON_OUT
    a = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    edit = a.metric == "lev"

    if a.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 only

    import torch
    import torch.distributed as dist
    import sahara_b200 as sb

    use_dist = world > 1 and a.impl == "b200"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sahara_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION / WARN make NCCL print its version on stdout, in front of the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = sb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    R, m, k = a.reads_per_step, a.len, a.errors
    n_batches = a.steps + a.warmup
    t0 = time.time()
    d_genome = ctx.synth_genome(a.genome, 42)
    ctx.build_index_device(d_genome, [a.genome], 6, 16)
    t_build = time.time() - t0
    info = ctx.info()
    scheme = sb.SearchScheme.generate(a.generator, 0, k, m, limit_to_hamming=not edit)
    ctx.set_scheme(scheme, edit)

    # distinct reads for every step and rank
    def batch_first_read(b):
        return (rank * n_batches + b) * R

    workload = (f"synthetic {a.genome / 1e9:.2f} Gbp random-DNA genome (seed 42), {R} x {m} bp reads per step per GPU "
                f"(seed 43, 90% sampled with <= {k} errors, 10% random, both strands), k={k} "
                f"{'edit' if edit else 'Hamming'} distance, generator {a.generator}")
    config = {"workload": workload, "genome_bp": a.genome, "reads_per_step_per_gpu": R, "read_len": m, "errors": k,
              "distance": "edit" if edit else "hamming", "generator": a.generator, "index_rows": info["n_rows"],
              "index_build_s": round(t_build, 2), "parallelism": f"query-sharded x{world}, index replicated",
              "l2_policy": "inputs larger than L2: index %.1f GB, a fresh 300 MB read batch every step" % (info["device_bytes"] / 1e9)}

    # ---------------- CPU arm / CPU baseline (rank 0) ----------------
    def cpu_baseline(sample_reads, steps, warm):
        import oracle as O
        view = ctx.download_view()
        try:
            oix = O.OracleIndex.from_view(view)
        finally:
            ctx.free_view(view)
        # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what is asked for here)
        threads = max(O.max_threads(), len(os.sched_getaffinity(0)))
        d_q = ctx.synth_reads(d_genome, a.genome, sample_reads * (steps + warm), m, k, edit, 43, batch_first_read(0))
        q = ctx.to_host(d_q, 2 * sample_reads * (steps + warm) * m).reshape(-1, m)
        ctx.device_free(d_q)
        times = []
        hits_total = 0
        for s in range(steps + warm):
            qs = q[2 * sample_reads * s: 2 * sample_reads * (s + 1)]
            t = time.perf_counter()
            cur = oix.search(qs, scheme, edit, threads)
            hits = oix.locate(cur, threads)
            dt = time.perf_counter() - t
            if s >= warm:
                times.append(dt)
                hits_total += hits.shape[0]
        # one-thread run, faithful to the single-threaded reference, on a smaller slice
        n1 = max(1, sample_reads // 8)
        t = time.perf_counter()
        cur1 = oix.search(q[: 2 * n1], scheme, edit, 1)
        oix.locate(cur1, 1)
        dt1 = time.perf_counter() - t
        return {"oix": oix, "threads": threads, "times": times, "reads": sample_reads, "one_thread_reads_s": n1 / dt1,
                "first_batch": q[: 2 * sample_reads], "hits": hits_total}

    if a.impl == "reference":
        sample = a.cpu_sample or 20_000
        res = cpu_baseline(sample, a.steps, a.warmup)
        total_t = sum(res["times"])
        value = sample * a.steps / total_t
        line = {"impl": "reference", "metric": "queries/s (150bp, k=2 edit)", "value": round(value, 1), "unit": "reads/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * total_t / a.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": dict(config, reads_per_step_per_gpu=sample),
                "cpu_baseline": {"value": round(value, 1), "unit": "reads/s", "cores": res["threads"], "kind": "port",
                                 "sample": f"{sample} reads per step (both strands), CPU oracle search+locate, "
                                           f"{res['threads']} OpenMP threads; 1 thread: {res['one_thread_reads_s']:.0f} reads/s"},
                "e2e": {"value": round(value, 1), "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit_json(line)
        return 0

    # ---------------- GPU arm ----------------
    qauto = a.qgram
    if qauto < 0:  # auto (as the CLI): the expected depth at which cursors become unique, floor(log4(rows)); 17 GB at most
        import math
        qauto = max(0, min(15, int(math.log(max(4, info["n_rows"]), 4))))
    if a.device_sa_rate:
        ctx.densify(a.device_sa_rate)
    if a.text:
        ctx.enable_text(True)
    if qauto:
        ctx.build_qgram(qauto)

    d_batches = [ctx.synth_reads(d_genome, a.genome, R, m, k, edit, 43, batch_first_read(b)) for b in range(n_batches)]

    def barrier():
        if use_dist:
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def allmax(x):
        if not use_dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if not use_dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing (`value`) ----
    for b in range(a.warmup):
        ctx.search_device(d_batches[b], 2 * R, m)
    ctx.reset_counters()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ms_search = ms_locate = ms_sort = ms_fm = ms_text = 0.0
    hits_total = cursors_total = 0
    for b in range(a.warmup, n_batches):
        nc, nh = ctx.search_device(d_batches[b], 2 * R, m)
        c = ctx.counters()
        ms_search += c["ms_search"]
        ms_locate += c["ms_locate"]
        ms_sort += c["ms_sort"]
        ms_fm += c["ms_fm"]
        ms_text += c["ms_text"]
        hits_total += nh
        cursors_total += nc
    e1.record(stream)
    barrier()
    dev_ms = allmax(e0.elapsed_time(e1))
    clocks = sampler.finish()
    ct = ctx.counters()
    launches = int(ct["kernel_launches"])
    value = world * R * a.steps / (dev_ms * 1e-3)

    # ---- end-to-end timing through the host-buffer C-ABI call ----
    import ctypes as C
    from sahara_b200._native import check, cuda
    host_batches = []
    for b in range(n_batches):  # queries go device -> pinned host once, outside the timed region
        t = torch.empty(2 * R * m, dtype=torch.uint8, pin_memory=True)
        check(cuda.sb200_copy_to_host(ctx._h, C.c_void_p(t.data_ptr()), C.c_void_p(d_batches[b]), 2 * R * m))
        host_batches.append(t)

    def e2e_step(t):
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search(ctx._h, C.c_void_p(t.data_ptr()), 2 * R, m, C.byref(p), C.byref(n)))
        first = 0
        if n.value:
            first = C.cast(p, C.POINTER(C.c_uint64))[0]  # touch the result on the host
        cuda.sb200_free(p)
        return n.value, first

    for b in range(a.warmup):
        e2e_step(host_batches[b])
    barrier()
    t_start = time.perf_counter()
    e2e_hits = 0
    for b in range(a.warmup, n_batches):
        nh, _ = e2e_step(host_batches[b])
        e2e_hits += nh
    barrier()
    e2e_s = allmax(time.perf_counter() - t_start)
    e2e_value = world * R * a.steps / e2e_s
    h2d = 2 * R * m
    d2h = int(32 * e2e_hits / a.steps)

    # the compact host-buffer call: reads only (reverse complements made on the device), 16-byte hits
    def e2e_reads_step(t):
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search_reads(ctx._h, C.c_void_p(t.data_ptr()), R, m, 1, C.byref(p), C.byref(n)))
        if n.value:
            C.cast(p, C.POINTER(C.c_uint32))[0]
        cuda.sb200_free(p)
        return n.value

    fwd_batches = []
    for b in range(n_batches):  # forward strands = every second query of the batch
        t = torch.empty(R * m, dtype=torch.uint8, pin_memory=True)
        t.view(R, m).copy_(host_batches[b].view(R, 2, m)[:, 0, :])
        fwd_batches.append(t)
    for b in range(a.warmup):
        e2e_reads_step(fwd_batches[b])
    barrier()
    t_start = time.perf_counter()
    compact_hits = 0
    for b in range(a.warmup, n_batches):
        compact_hits += e2e_reads_step(fwd_batches[b])
    barrier()
    e2e_compact_s = allmax(time.perf_counter() - t_start)

    # ---- roofline (SURVEY.md §8d accounting: one search node = 2 rank-ops = 128 B, however it is served) ----
    peak, peak_src = measured_peak()
    nodes = ct["nodes"]            # extensions over the K timed steps == the oracle's extension count (tests assert it)
    nodes_text = ct["nodes_text"]  # of those, verified in the text by text_pool_kernel
    nodes_fm = nodes - nodes_text
    traffic = ncu_traffic() or {}

    def kern(name, n_nodes, ms, what, bound):
        per_launch = n_nodes * 128 / a.steps
        ach = per_launch / (ms / a.steps * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": name, "ms_per_launch": round(ms / a.steps, 3), "nodes_per_launch": int(n_nodes / a.steps),
                "algorithmic_bytes_per_launch": int(per_launch), "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                "traffic": traffic.get(name), "limited_by": bound, "does": what}

    k_fm = kern("fm_items_kernel", nodes_fm, ms_fm, "cursor extensions by rank probes (cursors covering several rows); the time includes "
                "fm_roots_kernel (root frames of every query from the q-gram table)",
                "HBM random access: 38.4 G L2-miss requests/s measured (tools/gather_bench.cu), 1 request per probe")
    k_text = kern("text_pool_kernel" if a.text else "text_kernel", nodes_text, ms_text,
                  "cursor extensions of unique cursors verified in the text (warp-level frame pools)",
                  "instruction issue (76 % issue-active, DRAM 4 % busy): the probes are replaced by cached text symbols")
    dom = k_text if ms_text >= ms_fm else k_fm
    roofline = {"bound": "hbm", "kernel": "sb200::" + dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                "accounting": "SURVEY.md 8d: nodes x 128 B per launch / CUDA-event time of the launch; above 1.0 because "
                              "unique cursors are extended from the text (traffic = measured DRAM bytes, see profiles/)",
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "ms_per_launch": dom["ms_per_launch"],
                "kernels": [k_fm, k_text],
                "search_phase": {"nodes_per_step": int(nodes / a.steps), "ms_per_step": round(ms_search / a.steps, 3),
                                 "rank_ops_per_s": round(2 * nodes / (ms_search * 1e-3), 1),
                                 "achieved": round(nodes * 128 / (ms_search * 1e-3) / 1e9, 1),
                                 "frac": round(nodes * 128 / (ms_search * 1e-3) / 1e9 / peak, 4)},
                "random_access_cap_gbs": 2457.6,
                "phase_ms_per_step": {"search": round(ms_search / a.steps, 3), "locate": round(ms_locate / a.steps, 3),
                                      "sort": round(ms_sort / a.steps, 3)},
                "qgram": qauto, "text_mode": bool(a.text), "lf_steps_per_step": int(ct["lf_steps"] / a.steps)}

    line = {"metric": "queries/s (150bp, k=2 edit)", "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(dev_ms / a.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": round(world * R * a.steps / e2e_compact_s, 1), "unit": "reads/s", "h2d_bytes_per_step": R * m,
                    "d2h_bytes_per_step": int(16 * compact_hits / a.steps), "ms_per_step": round(1e3 * e2e_compact_s / a.steps, 3),
                    "call": "sb200_search_reads (reads in, reverse complements on the device, 16-byte hits out)"},
            "e2e_full_tuples": {"value": round(e2e_value, 1), "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                "ms_per_step": round(1e3 * e2e_s / a.steps, 3), "hits_match": bool(compact_hits == e2e_hits),
                                "call": "sb200_search (both strands in, 32-byte hit tuples out)"},
            "gpu_launches": launches, "roofline": roofline,
            "hits_per_step": int(hits_total / a.steps), "cursors_per_step": int(cursors_total / a.steps)}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:  # (the contract: on rank 0 at N = 1 only)
        sample = a.cpu_sample or 20_000
        res = cpu_baseline(sample, 1, 0)
        cpu_v = sample / res["times"][0]
        # parity of the sample through the C ABI against the oracle (checker, not the measured path)
        oix = res["oix"]
        nodes_before = int(oix.counters[0])
        cur = oix.search(res["first_batch"], scheme, edit, res["threads"])
        oracle_nodes = int(oix.counters[0]) - nodes_before
        import oracle as O
        want = O.sort_rows(oix.locate(cur, res["threads"]))
        got = ctx.search(res["first_batch"])
        line["parity_sample_ok"] = bool(np.array_equal(got, want))
        # the roofline numerator: with the q-gram table off the kernels expand exactly the oracle's extensions
        ctx.build_qgram(0)
        ctx.reset_counters()
        ctx.search_cursors(res["first_batch"])
        line["nodes_sample"] = {"kernels": int(ctx.counters()["nodes"]), "oracle": oracle_nodes,
                                "equal": bool(ctx.counters()["nodes"] == oracle_nodes)}
        if qauto:
            ctx.build_qgram(qauto)
        line["cpu_baseline"] = {"value": round(cpu_v, 1), "unit": "reads/s", "cores": res["threads"], "kind": "port",
                                "sample": f"first {sample} reads of rank 0's first batch (both strands), CPU oracle search+locate, "
                                          f"{res['threads']} OpenMP threads",
                                "one_thread_value": round(res["one_thread_reads_s"], 1)}
    if rank == 0:
        emit_json(line)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
