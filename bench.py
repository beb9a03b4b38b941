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
    ap.add_argument("--opt", action="append", default=[], help="name=value for sb200_set_option (experiments; results never depend on it)")
    ap.add_argument("--total-reads", type=int, default=10_000_000, help="configs[3] as stated: reads sharded over the GPUs (strong scaling leg)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU while the timed region runs: NVML in this process every few
    milliseconds (a timed region of some 60 ms is over before an nvidia-smi child has printed its first line), one sample
    taken synchronously when the region starts; nvidia-smi -lms only when NVML cannot be loaded"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None
        self.interval = float(os.environ.get("SB200_BENCH_SAMPLE_MS", "4")) * 1e-3  # NVML polling period
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:
                import torch
                pr = torch.cuda.get_device_properties(gpu)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            self.nvml = None

    def sample(self):
        n = self.nvml
        try:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            flags = ["Active" if r & b else "Not Active" for b in (0x8, 0x40, 0x20, 0x4)]  # hw, hw thermal, sw thermal, sw power cap
            self.rows.append([str(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), str(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)),
                              "0"] + flags)
        except Exception:
            pass

    def run(self):
        if self.nvml is not None:
            while not self.stop_flag.is_set():
                self.sample()
                time.sleep(self.interval)
            return
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
        if self.nvml is not None:
            self.sample()  # (the GPU is still busy with the last step's tail when the region's closing synchronize returns)
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


def metric_name(a):
    """BASELINE.json's metric for the default shape; another shape (--len / --errors / --metric) is named as what it is"""
    return f"queries/s ({a.len}bp, k={a.errors} {'edit' if a.metric == 'lev' else 'Hamming'})"


def kernel_profile():
    """per-launch counters of the search kernels from the committed ncu captures (profiles/r02_kernel_counters.json,
    written by tools/ncu_counters.py): warp instructions, lanes per instruction, DRAM bytes, L2-miss requests and the node
    count of the profiled launch"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_kernel_counters.json")) as f:
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
    for o in a.opt:
        ctx.set_option(o.split("=")[0], int(o.split("=")[1]))
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
        line = {"impl": "reference", "metric": metric_name(a), "value": round(value, 1), "unit": "reads/s",
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
    t_prep = time.time()
    if a.device_sa_rate:
        ctx.densify(a.device_sa_rate)
    if a.text:
        ctx.enable_text(True)
    if qauto:
        ctx.build_qgram(qauto)
    config["index_prepare_s"] = round(time.time() - t_prep, 2)  # derived tables: complete SA + inverse + packed text, q-gram table
    config["index_device_gb"] = round(ctx.info()["device_bytes"] / 1e9, 1)

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

    import ctypes as C
    from sahara_b200._native import check, cuda

    def pipelined(submit, first, last, copy, depth=2):
        """batches first..last-1 through sb200_submit_* / sb200_wait_batch, `depth` in flight; -> per-batch results"""
        tickets, out = [], []
        for i in range(first, last + depth):
            j = i - depth
            if j >= first:
                res = ctx.wait_batch(tickets[j - first], copy_to_host=copy)
                touched = 0
                if copy and res.n_hits:  # touch the result on the host: first and last record, last end
                    rec = C.cast(res.records, C.POINTER(C.c_uint8))
                    touched = rec[0] + rec[res.n_record_bytes - 1] + C.cast(res.hit_end, C.POINTER(C.c_uint32))[res.n_queries - 1]
                out.append(dict(n_hits=res.n_hits, n_cursors=res.n_cursors, ms_search=res.ms_search, ms_locate=res.ms_locate,
                                ms_sort=res.ms_sort, h2d=res.h2d_bytes, d2h=res.d2h_bytes, touched=touched))
                ctx.release_batch(tickets[j - first])
            if i < last:
                tickets.append(submit(i))
        return out

    # ---- device-resident timing (`value`): queries already in HBM, hits left in HBM; batches submitted two deep ----
    pipelined(lambda b: ctx.submit_device(d_batches[b], 2 * R, m), 0, a.warmup, False)
    ctx.reset_counters()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)  # (submit forks from this stream, wait joins it again: the events bracket every batch)
    res_dev = pipelined(lambda b: ctx.submit_device(d_batches[b], 2 * R, m), a.warmup, n_batches, False)
    e1.record(stream)
    barrier()
    my_ms = e0.elapsed_time(e1)
    dev_ms = allmax(my_ms)
    clocks = sampler.finish()
    ct = ctx.counters()
    # what every rank saw in the timed region: its own time and the batches it had to run again (buffer estimates)
    rank_info = [{"ms_per_step": round(my_ms / a.steps, 3), "batch_restarts": int(ct["batch_restarts"]), "sm_mhz": clocks.get("sm_mhz"),
                  "reasons": clocks.get("reasons"), "hits_per_step": int(sum(r["n_hits"] for r in res_dev) / a.steps),
                  "nodes_per_step": int(ct["nodes"] / a.steps)}]
    if use_dist:
        gathered = [None] * world
        dist.all_gather_object(gathered, rank_info[0])
        rank_info = gathered
    launches = int(ct["kernel_launches"])
    value = world * R * a.steps / (dev_ms * 1e-3)
    hits_total = sum(r["n_hits"] for r in res_dev)
    cursors_total = sum(r["n_cursors"] for r in res_dev)

    # ---- one batch at a time (no overlap between batches): per-kernel CUDA-event times for the roofline ----
    ctx.reset_counters()
    ms_search = ms_locate = ms_sort = ms_fm = ms_text = 0.0
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for b in range(a.warmup, n_batches):
        ctx.search_device(d_batches[b], 2 * R, m)
        c = ctx.counters()
        ms_search += c["ms_search"]
        ms_locate += c["ms_locate"]
        ms_sort += c["ms_sort"]
        ms_fm += c["ms_fm"]
        ms_text += c["ms_text"]
    e3.record(stream)
    torch.cuda.synchronize()
    serial_ms = e2.elapsed_time(e3)
    ct = ctx.counters()

    # ---- end to end through the host-buffer calls: pinned host buffers, H2D and D2H inside the timed region ----
    W4, W2 = (m + 7) // 8, (m + 15) // 16
    host_reads, host_packed, host_packed2 = [], [], []
    for b in range(n_batches):  # queries go device -> pinned host once, outside the timed region
        tq = torch.empty(2 * R * m, dtype=torch.uint8, pin_memory=True)
        check(cuda.sb200_copy_to_host(ctx._h, C.c_void_p(tq.data_ptr()), C.c_void_p(d_batches[b]), 2 * R * m))
        tr = torch.empty(R * m, dtype=torch.uint8, pin_memory=True)
        tr.view(R, m).copy_(tq.view(R, 2, m)[:, 0, :])  # forward strands = every second query of the batch
        host_reads.append(tr)
        tp = torch.empty(R * W4, dtype=torch.int32, pin_memory=True)  # what the host-side reader hands over: 4 bits per base
        sb.pack_reads4(tr.view(R, m).numpy(), threads=8, out=tp.view(R, W4).numpy().view(np.uint32))
        host_packed.append(tp)
        tp2 = torch.empty(R * W2, dtype=torch.int32, pin_memory=True)  # 2 bits per base (the synthetic reads hold A, C, G, T only)
        sb.pack_reads2(tr.view(R, m).numpy(), threads=8, out=tp2.view(R, W2).numpy().view(np.uint32))
        host_packed2.append(tp2)
        del tq

    def e2e_run(submit):
        pipelined(submit, 0, a.warmup, True)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start = time.perf_counter()
        ea.record(stream)
        res = pipelined(submit, a.warmup, n_batches, True)
        eb.record(stream)
        wall = time.perf_counter() - t_start
        barrier()
        return allmax(wall), allmax(ea.elapsed_time(eb) * 1e-3), res

    # the call a user of this library makes: reads 2-bit packed by the host reader, hits back as delta-coded CSR records
    e2e_s, e2e_dev_s, res_e2e = e2e_run(lambda b: ctx.submit_reads((host_packed2[b].data_ptr(), R, m), packed2=True))
    # the same with 4 bits per base in (reads that hold N)
    e2e_p4_s, _, res_p4 = e2e_run(lambda b: ctx.submit_reads((host_packed[b].data_ptr(), R, m), packed4=True))
    # the same with one byte per base in (the reference's std::vector<uint8_t> per read)
    e2e_ranks_s, _, res_ranks = e2e_run(lambda b: ctx.submit_reads((host_reads[b].data_ptr(), R, m), packed4=False))

    # round 1's call for comparison: synchronous, ranks in, 16-byte hits out
    def sync_reads_step(t):
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search_reads(ctx._h, C.c_void_p(t.data_ptr()), R, m, 1, C.byref(p), C.byref(n)))
        if n.value:
            C.cast(p, C.POINTER(C.c_uint32))[0]
        cuda.sb200_free(p)
        return n.value

    for b in range(a.warmup):
        sync_reads_step(host_reads[b])
    barrier()
    t_start = time.perf_counter()
    sync_hits = 0
    for b in range(a.warmup, n_batches):
        sync_hits += sync_reads_step(host_reads[b])
    barrier()
    e2e_sync_s = allmax(time.perf_counter() - t_start)

    # ---- BASELINE.json configs[3] as stated: 10 M reads in total, query-sharded over the N GPUs (strong scaling) ----
    total10 = a.total_reads
    share = (total10 + world - 1) // world          # reads of this rank
    n10 = max(1, (share + R - 1) // R)              # batches of at most R reads
    sizes10 = [min(R, share - i * R) for i in range(n10)]
    barrier()
    t_start = time.perf_counter()
    hits10 = 0
    tickets = []
    for i in range(n10 + 2):
        if i >= 2:
            res = ctx.wait_batch(tickets[i - 2], copy_to_host=True)
            hits10 += res.n_hits
            ctx.release_batch(tickets[i - 2])
        if i < n10:  # (the reads of the timed batches, reused round robin: the index is far larger than L2)
            tickets.append(ctx.submit_reads((host_packed2[i % n_batches].data_ptr(), sizes10[i], m), packed2=True))
    barrier()
    strong_s = allmax(time.perf_counter() - t_start)

    # ---- roofline ----
    peak, peak_src = measured_peak()
    nodes = ct["nodes"]            # extensions over the K timed steps == the oracle's extension count (tests assert it)
    nodes_text = ct["nodes_text"]  # of those, verified in the text by text_pool_kernel
    nodes_fm = nodes - nodes_text
    prof = kernel_profile() or {}
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions/s: 4 schedulers per SM, one instruction per clock each

    def kern(name, n_nodes, ms, what):
        ms1 = ms / a.steps
        n1 = n_nodes / a.steps
        p = prof.get(name, {})
        # the committed capture belongs to one workload shape; its per-node instruction and sector counts do not carry over
        if p and f" x {a.len} bp, k={a.errors} {a.metric}" not in p.get("workload", ""):
            p = {}
        out = {"kernel": "sb200::" + name, "ms_per_launch": round(ms1, 3), "nodes_per_launch": int(n1), "does": what}
        # scaled from the committed ncu capture by the node count (the workload is seeded: same launches, same counts)
        scale = n1 / p["nodes_per_launch"] if p.get("nodes_per_launch") else None
        if scale:
            inst = p["inst_executed"] * scale
            out["issue"] = {"warp_inst_per_launch": int(inst), "achieved": round(inst / (ms1 * 1e-3) / 1e9, 1), "peak": round(issue_peak, 1),
                            "unit": "G warp-inst/s", "frac": round(inst / (ms1 * 1e-3) / 1e9 / issue_peak, 3),
                            "lanes_per_inst": p.get("lanes_per_inst"), "sm_mhz": sm_mhz}
            dram = p["dram_bytes"] * scale
            out["hbm"] = {"traffic": int(dram), "achieved": round(dram / (ms1 * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(dram / (ms1 * 1e-3) / 1e9 / peak, 3)}
            if p.get("l2_miss_requests"):
                req = p["l2_miss_requests"] * scale
                out["random_access"] = {"l2_miss_requests_per_launch": int(req), "achieved": round(req / (ms1 * 1e-3) / 1e9, 2), "peak": 38.4,
                                        "unit": "G requests/s", "frac": round(req / (ms1 * 1e-3) / 1e9 / 38.4, 3),
                                        "peak_source": "profiles/r01_gather_microbench.txt (dependent random gather, 3 GiB table)"}
            out["profile"] = p.get("source")
        # SURVEY.md 8d books a node at 2 rank-ops = 128 B "however it is served": kept as a speed-up over the reference's
        # memory behaviour, NOT as a fraction of a hardware ceiling
        out["algorithmic_speedup"] = {"booked_bytes_per_launch": int(n1 * 128), "booked_gbs": round(n1 * 128 / (ms1 * 1e-3) / 1e9, 1),
                                      "over_hbm_peak": round(n1 * 128 / (ms1 * 1e-3) / 1e9 / peak, 2)}
        return out

    k_fm = kern("fm_items_kernel", nodes_fm, ms_fm, "cursor extensions by rank probes (cursors covering several rows); the time includes "
                "fm_roots_kernel (root frames of every query from the q-gram table)")
    k_text = kern("text_pool_kernel", nodes_text, ms_text, "cursor extensions of unique cursors verified in the text (warp-level frame pools)")
    dom = k_text if ms_text >= ms_fm else k_fm
    if "issue" in dom and dom is k_text:
        roofline = {"bound": "issue", "kernel": dom["kernel"], "achieved": dom["issue"]["achieved"], "peak": dom["issue"]["peak"],
                    "unit": "G warp-inst/s", "frac": dom["issue"]["frac"], "traffic": dom["hbm"]["traffic"],
                    "why": "text_pool_kernel replaces the probes of the occurrence table by cached text symbols: DRAM is nearly idle "
                           "(hbm.frac), what limits it is instruction issue — warp instructions per launch (ncu, profiles/) over the "
                           "CUDA-event time against 148 SMs x 4 schedulers x the SM clock sampled during the run; lanes_per_inst of 32 "
                           "says how much of each issued instruction does work"}
    elif "random_access" in dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["random_access"]["achieved"] * 64, "peak": 38.4 * 64,
                    "unit": "GB/s", "frac": dom["random_access"]["frac"], "traffic": dom["hbm"]["traffic"],
                    "why": "random 32-byte gathers: L2-miss requests per launch over the CUDA-event time against the measured request rate"}
    else:  # no committed profile for this kernel: the HBM view with the SURVEY booking, flagged as such
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["algorithmic_speedup"]["booked_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": None, "traffic": None, "why": "no ncu capture committed for this configuration: frac withheld"}
    roofline.update({"peak_source": peak_src, "ms_per_launch": dom["ms_per_launch"], "kernels": [k_fm, k_text],
                     "search_phase": {"nodes_per_step": int(nodes / a.steps), "ms_per_step": round(ms_search / a.steps, 3),
                                      "rank_ops_per_s_booked": round(2 * nodes / (ms_search * 1e-3), 1)},
                     "phase_ms_per_step": {"search": round(ms_search / a.steps, 3), "locate": round(ms_locate / a.steps, 3),
                                           "sort": round(ms_sort / a.steps, 3), "one_batch_at_a_time": round(serial_ms / a.steps, 3)},
                     "qgram": qauto, "text_mode": bool(a.text), "lf_steps_per_step": int(ct["lf_steps"] / a.steps)})

    def brk(res, secs):
        n = len(res)
        return {"ms_per_step": round(1e3 * secs / a.steps, 3), "kernel_ms_per_step": round(sum(r["ms_search"] + r["ms_locate"] + r["ms_sort"] for r in res) / n, 3),
                "h2d_ms_at_55GBs": round(res[0]["h2d"] / 55e9 * 1e3, 3), "d2h_ms_at_55GBs": round(sum(r["d2h"] for r in res) / n / 55e9 * 1e3, 3)}

    line = {"metric": metric_name(a), "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(dev_ms / a.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": dict({"value": round(world * R * a.steps / e2e_s, 1), "unit": "reads/s", "h2d_bytes_per_step": int(res_e2e[0]["h2d"]),
                         "d2h_bytes_per_step": int(sum(r["d2h"] for r in res_e2e) / len(res_e2e)),
                         "device_event_ms_per_step": round(1e3 * e2e_dev_s / a.steps, 3),
                         "call": "sb200_submit_reads(SB200_READS_PACKED2) / sb200_wait_batch, 2 batches in flight: 2-bit packed reads in "
                                 "(reverse complements on the device), hits out as delta-coded CSR records (5 bytes for the first hit of a query, a varint difference for the others) + 4 bytes per query",
                         "hits_match_device_run": bool(sum(r["n_hits"] for r in res_e2e) == hits_total)}, **brk(res_e2e, e2e_s)),
            "e2e_packed4_in": dict({"value": round(world * R * a.steps / e2e_p4_s, 1), "unit": "reads/s", "h2d_bytes_per_step": int(res_p4[0]["h2d"]),
                                    "d2h_bytes_per_step": int(sum(r["d2h"] for r in res_p4) / len(res_p4)),
                                    "call": "sb200_submit_reads(SB200_READS_PACKED4): 4 bits per base in (what reads with N need), delta-coded CSR records out"},
                                   **brk(res_p4, e2e_p4_s)),
            "e2e_rank_bytes_in": dict({"value": round(world * R * a.steps / e2e_ranks_s, 1), "unit": "reads/s", "h2d_bytes_per_step": int(res_ranks[0]["h2d"]),
                                       "d2h_bytes_per_step": int(sum(r["d2h"] for r in res_ranks) / len(res_ranks)),
                                       "call": "sb200_submit_reads(SB200_READS_RANKS): one byte per base in, CSR records out"}, **brk(res_ranks, e2e_ranks_s)),
            "e2e_round1_call": {"value": round(world * R * a.steps / e2e_sync_s, 1), "unit": "reads/s", "h2d_bytes_per_step": R * m,
                                "d2h_bytes_per_step": int(16 * sync_hits / a.steps), "ms_per_step": round(1e3 * e2e_sync_s / a.steps, 3),
                                "call": "sb200_search_reads (synchronous; ranks in, 16-byte hits out), chunks pipelined inside the call"},
            "cfg4_10M_reads_sharded": {"total_reads": total10, "reads_per_gpu": share, "seconds": round(strong_s, 4),
                                       "value": round(total10 / strong_s, 1), "unit": "reads/s", "scaling": "strong", "hits_this_rank": int(hits10),
                                       "what": "BASELINE.json configs[3] as stated: 10 M x 150 bp reads, k=2 edit, sharded over the GPUs; "
                                               "end to end through sb200_submit_reads with host buffers (wall clock, max over ranks)"},
            "gpu_launches": launches, "roofline": roofline,
            "hits_per_step": int(hits_total / a.steps), "cursors_per_step": int(cursors_total / a.steps),
            "batch_restarts": sum(r["batch_restarts"] for r in rank_info), "ranks": rank_info}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:  # (the contract: on rank 0 at N = 1 only)
        sample = a.cpu_sample or 20_000
        res = cpu_baseline(sample, 1, 0)
        cpu_v = sample / res["times"][0]
        # parity of samples through the C ABI against the oracle (checker, not the measured path): the first reads of the
        # first batch and the LAST reads of the last timed batch
        oix = res["oix"]
        import oracle as O

        def parity(qs):
            before = int(oix.counters[0])
            cur = oix.search(qs, scheme, edit, res["threads"])
            oracle_nodes = int(oix.counters[0]) - before
            want = O.sort_rows(oix.locate(cur, res["threads"]))
            got = ctx.search(qs)
            reads = np.ascontiguousarray(qs[0::2])
            got_async = ctx.search_reads_async(reads, packed2=True, batch=reads.shape[0] // 2 + 1)
            return bool(np.array_equal(got, want) and np.array_equal(got_async, want)), oracle_nodes, int(want.shape[0])

        ok_first, oracle_nodes, n_first = parity(res["first_batch"])
        last = ctx.to_host(d_batches[n_batches - 1] + 2 * (R - sample) * m, 2 * sample * m).reshape(-1, m)
        ok_last, _, n_last = parity(last)
        line["parity_sample_ok"] = bool(ok_first and ok_last)
        line["parity_samples"] = {"first_batch_first_reads": {"reads": sample, "hits": n_first, "ok": ok_first},
                                  "last_batch_last_reads": {"reads": sample, "hits": n_last, "ok": ok_last},
                                  "checked": "sb200_search and sb200_submit_reads(2-bit packed reads, delta-coded records) hit lists == oracle search + locate, bit-exact"}
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
        # ---- the two other kernels of the path, timed alone on rank 0 (BASELINE.json: "occ-rank ops/s vs HBM roofline";
        # north_star kernel 3, the locate that walks LF to sampled suffix-array rows — the search above never needs it,
        # its verified occurrences carry text positions) ----
        try:
            chains, iters = 1 << 22, 64
            ms_r, _ = ctx.rank_bench(0, chains, iters, 7)
            ops = chains * iters
            line["rank_ops"] = {"value": round(ops / (ms_r * 1e-3), 1), "unit": "all-symbol rank ops/s", "ms": round(ms_r, 3),
                                "what": f"{chains} independent chains of {iters} dependent all_ranks probes at pseudo-random rows of the "
                                        "occurrence table (one 32-byte block per probe, the superblock from L2)",
                                "booked_gbs_at_64B_per_op": round(ops * 64 / (ms_r * 1e-3) / 1e9, 1), "hbm_peak_gbs": peak,
                                "frac_of_hbm_at_64B_per_op": round(ops * 64 / (ms_r * 1e-3) / 1e9 / peak, 3),
                                "frac_of_random_access_ceiling": round(ops / (ms_r * 1e-3) / 1e9 / 38.4, 3),
                                "random_access_ceiling": "38.4 G requests/s (profiles/r01_gather_microbench.txt)"}
            # LF-walking locate: without the verification tables every row is walked to a sampled row (rate 16)
            n_loc = 4_000_000
            rng = np.random.default_rng(11)
            cur = np.zeros((n_loc, 4), dtype=np.uint64)
            cur[:, 0] = np.arange(n_loc) // 8
            cur[:, 1] = rng.integers(0, info["n_rows"] - 1, size=n_loc)
            cur[:, 2] = 1
            ctx.build_qgram(0)
            ctx.enable_text(False)
            # (an index built on the device keeps its complete suffix array; the file image has the reference's samples:
            # download it and upload it again, as `sahara search` does with X.idx)
            view = ctx.download_view()
            try:
                ctx.upload_view(view)
            finally:
                ctx.free_view(view)
            ctx.locate(cur)  # (untimed: sizes the work buffers)
            ctx.reset_counters()
            t_l = time.perf_counter()
            located = ctx.locate(cur)
            wall_l = time.perf_counter() - t_l
            cl = ctx.counters()
            ms_l = cl["ms_locate"]
            line["lf_locate"] = {"rows": n_loc, "hits": int(located.shape[0]), "lf_steps": int(cl["lf_steps"]), "ms_kernel": round(ms_l, 3),
                                 "rows_per_s": round(n_loc / (ms_l * 1e-3), 1), "lf_steps_per_s": round(cl["lf_steps"] / (ms_l * 1e-3), 1),
                                 "requests_per_step": 2, "frac_of_random_access_ceiling": round(2 * cl["lf_steps"] / (ms_l * 1e-3) / 1e9 / 38.4, 3),
                                 "wall_s_with_copies": round(wall_l, 3),
                                 "what": "sb200_locate of random single rows of the 3.1 Gbp index with the sampled suffix array only (rate 16): "
                                         "one marker record + one occurrence block per LF step, issued together"}
        except Exception as ex:  # (never fails the bench line)
            line["extra_legs_error"] = str(ex)[:200]
        emit_json(line)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
