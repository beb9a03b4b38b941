"""Per-chunk timeline of the host-buffer call (SB200_DEBUG lines) for a few chunk sizes.  GPU only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import torch
import sahara_b200 as sb
from sahara_b200._native import cuda, check
n = int(os.environ.get("GENOME", 3100000000)); R = 1000000; m = 150; k = 2
ctx = sb.Context(0)
dg = ctx.synth_genome(n, 42); ctx.build_index_device(dg, [n], 6, 16); ctx.enable_text(True); ctx.build_qgram(int(os.environ.get('QGRAM', 15)))
ctx.set_scheme(sb.SearchScheme.generate("h2-k2", 0, k, m), True)
bufs = []
for b in range(3):
    dq = ctx.synth_reads(dg, n, R, m, k, True, 43, b * R)
    t = torch.empty(2 * R * m, dtype=torch.uint8, pin_memory=True)
    check(cuda.sb200_copy_to_host(ctx._h, C.c_void_p(t.data_ptr()), C.c_void_p(dq), 2 * R * m)); ctx.device_free(dq)
    f = torch.empty(R * m, dtype=torch.uint8, pin_memory=True)
    f.view(R, m).copy_(t.view(R, 2, m)[:, 0, :]); bufs.append(f)
def call(p):
    out, nh = C.c_void_p(), C.c_uint64()
    t = time.perf_counter()
    check(cuda.sb200_search_reads(ctx._h, C.c_void_p(p.data_ptr()), R, m, 1, C.byref(out), C.byref(nh)))
    dt = time.perf_counter() - t
    cuda.sb200_free(out)
    return dt
for spec in os.environ.get("CHUNKS", "2000000:8,2000000:6,2000000:12,750000:8,1000000:8").split(","):
    chunk, div = spec.split(":")
    os.environ["SB200_CHUNK"] = chunk; os.environ["SB200_EDGE_DIV"] = div
    os.environ.pop("SB200_DEBUG", None)
    call(bufs[0]); call(bufs[1])
    ts = [call(bufs[i % 3]) for i in range(6)]
    print("chunk", spec, "ms per call", [round(t * 1e3, 2) for t in ts], flush=True)
    os.environ["SB200_DEBUG"] = "0"
    call(bufs[2])
