"""ctypes wrapper of the host emulation of the CUDA search kernel (tests/host_emu/search_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "search_emu.cpp")


def _load(name, flags):
    so = os.path.join(HERE, name)
    deps = [SRC] + [os.path.join(HERE, "..", "..", "sahara_b200", "csrc", f) for f in ("search.cuh", "layout.cuh")]
    deps.append(os.path.join(HERE, "..", "..", "include", "sahara_policy.h"))
    if not (os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps)):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=undefined",
                               "-fno-sanitize-recover=undefined", *flags, SRC, "-o", so])
    lib = C.CDLL(so)
    lib.emu_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.emu_free.argtypes = [C.c_void_p]
    lib.emu_set_max_hits.argtypes = [C.c_uint32]
    lib.emu_set_policy.argtypes = [C.c_void_p]
    lib.emu_ordered_max_depth.restype = C.c_uint32
    return lib


# EMU_FLAGS (analysis runs only): extra -D flags for the kernel source, compiled into a library of its own
_extra = os.environ.get("EMU_FLAGS", "").split()
_lib = _load("libsearch_emu_%08x.so" % (hash(tuple(_extra)) & 0xffffffff), _extra) if _extra else _load("libsearch_emu.so", [])
# the same source with a 24-frame pool and 120 spill frames: the pooled text kernel then spills all the time and
# pops narrowly / depth first most of the time (head room = the private-stack bound of 96 frames)
_lib_small_pool = _load("libsearch_emu_smallpool.so", ["-DSB200_POOL_CAP=24", "-DSB200_SPILL_CAP=120"])

def set_policy(policy=None):
    """the table of reconstructed rules the emulated kernels get (ctypes structure laid out like sb200_policy);
    None = the default"""
    for lib in (_lib, _lib_small_pool):
        lib.emu_set_policy(C.byref(policy) if policy is not None else None)



def QGRAM(q):
    """debug flag bits: start the searches from a q-gram jump table (q <= 6 here)"""
    return (q & 0xF) << 8


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def text_tables(oracle_index, seqs):
    """suffix array (global text positions) and the delimited text, for the in-text verification mode"""
    n = oracle_index.info()["n_rows"]
    starts = np.cumsum([0] + [len(s) + 1 for s in seqs])
    loc = oracle_index.locate_rows(np.arange(n, dtype=np.uint64))
    sa = (starts[loc[:, 0].astype(np.int64)] + loc[:, 1].astype(np.int64)).astype(np.uint32)
    text = np.concatenate([np.concatenate([np.asarray(s, np.uint8), np.zeros(1, np.uint8)]) for s in seqs])
    return np.ascontiguousarray(sa), np.ascontiguousarray(text)


def search(oracle_index, queries, scheme, edit, debug_flags=0, text=None, small_pool=False, max_hits=0):
    """runs the kernel body on the host over the oracle index's BWTs -> (sorted cursors uint64 [n,4], nodes).
    text = (sa32, text symbols) from text_tables() enables the in-text verification mode.
    max_hits > 0 runs the ordered walk of search_n (fm_ordered_kernel) instead."""
    info = oracle_index.info()
    bwt = np.ascontiguousarray(oracle_index.bwt(0))
    rev = np.ascontiguousarray(oracle_index.bwt(1))
    Carr = np.array(info["C"], dtype=np.uint64)
    q = np.ascontiguousarray(queries, dtype=np.uint8)
    out, n, nodes = C.c_void_p(), C.c_uint64(), C.c_uint64()
    lib = _lib_small_pool if small_pool else _lib
    lib.emu_set_max_hits(int(max_hits))
    rc = lib.emu_search(_p(bwt), _p(rev), info["n_rows"], info["sigma"], _p(Carr), _p(q), q.shape[0], q.shape[1], scheme.n_searches,
                         _p(scheme.pi), _p(scheme.l), _p(scheme.u), int(edit), debug_flags,
                         _p(text[0]) if text else None, _p(text[1]) if text else None, C.byref(out), C.byref(n), C.byref(nodes))
    if rc != 0:
        raise RuntimeError(f"emu_search failed with code {rc}")
    try:
        a = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(max(1, n.value) * 4,))[: n.value * 4].copy()
    finally:
        lib.emu_free(out)
    a = a.reshape(-1, 4).astype(np.uint64)
    if a.shape[0]:
        a = a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
    return a, nodes.value
