import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sahara_b200 as sb
from sahara_b200._native import cuda, check
n = int(os.environ.get("GENOME", 3100000000)); R = 1000000; m = 150; k = 2
ctx = sb.Context(0)
dg = ctx.synth_genome(n, 42); ctx.build_index_device(dg, [n], 6, 16); ctx.enable_text(True); ctx.build_qgram(12)
ctx.set_scheme(sb.SearchScheme.generate("h2-k2", 0, k, m), True)
bufs = []
for b in range(4):
    dq = ctx.synth_reads(dg, n, R, m, k, True, 43, b * R)
    p = C.c_void_p(); check(cuda.sb200_host_alloc(2 * R * m, C.byref(p)))
    check(cuda.sb200_copy_to_host(ctx._h, p, C.c_void_p(dq), 2 * R * m)); bufs.append(p); ctx.device_free(dq)
for i, p in enumerate(bufs):
    if i == 3: os.environ["SB200_DEBUG"] = "0"
    out, nh = C.c_void_p(), C.c_uint64()
    t = time.perf_counter()
    check(cuda.sb200_search(ctx._h, p, 2 * R, m, C.byref(out), C.byref(nh)))
    dt = time.perf_counter() - t
    cuda.sb200_free(out)
    print("call", i, round(dt * 1e3, 2), "ms", nh.value, flush=True)
