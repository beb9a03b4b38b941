"""ctypes binding of oracle/libsahara_oracle.so — the CPU restatement used as the checker.
Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = C.CDLL(os.path.join(ROOT, "oracle", "libsahara_oracle.so"))

u64p = C.POINTER(C.c_uint64)
_lib.orc_last_error.restype = C.c_char_p
_lib.orc_max_threads.restype = C.c_int
_lib.orc_index_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
_lib.orc_index_free.argtypes = [C.c_void_p]
_lib.orc_index_from_view.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
_lib.orc_index_save.argtypes = [C.c_void_p, C.c_char_p]
_lib.orc_index_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
_lib.orc_index_info.argtypes = [C.c_void_p, C.c_void_p]
_lib.orc_all_ranks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
_lib.orc_bwt_symbols.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
_lib.orc_locate_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
_lib.orc_index_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
_lib.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                            C.c_int, C.POINTER(C.c_void_p), u64p, C.c_void_p]
_lib.orc_search_n.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int, C.c_uint64, C.POINTER(C.c_void_p), u64p, C.c_void_p]
_lib.orc_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), u64p, C.c_void_p]
_lib.orc_free.argtypes = [C.c_void_p]
_lib.orc_set_policy.argtypes = [C.c_void_p]
_lib.orc_bf_hamming.restype = C.c_uint64
_lib.orc_bf_hamming.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
_lib.orc_bf_edit_starts.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(_lib.orc_last_error().decode(errors="replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def set_policy(policy=None):
    """replaces the oracle's table of reconstructed rules (include/sahara_policy.h; a ctypes structure with the layout of
    sb200_policy); None restores the default"""
    _check(_lib.orc_set_policy(C.byref(policy) if policy is not None else None))


def max_threads():
    return _lib.orc_max_threads()


class OracleIndex:
    def __init__(self, handle):
        self._h = handle
        self.counters = np.zeros(8, dtype=np.uint64)

    @staticmethod
    def build(seqs, sigma=6, sampling_rate=16):
        seqs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        lens = np.array([s.size for s in seqs], dtype=np.uint64)
        cat = np.ascontiguousarray(np.concatenate(seqs)) if seqs else np.zeros(1, np.uint8)
        h = C.c_void_p()
        _check(_lib.orc_index_build(_ptr(cat), _ptr(lens), len(seqs), sigma, sampling_rate, C.byref(h)))
        return OracleIndex(h)

    @staticmethod
    def from_view(view):
        """view: ctypes structure with the layout of sb200_index_view (arrays are copied)"""
        h = C.c_void_p()
        _check(_lib.orc_index_from_view(C.byref(view), C.byref(h)))
        return OracleIndex(h)

    @staticmethod
    def load(path):
        h = C.c_void_p()
        _check(_lib.orc_index_load(str(path).encode(), C.byref(h)))
        return OracleIndex(h)

    def save(self, path):
        _check(_lib.orc_index_save(self._h, str(path).encode()))

    def close(self):
        if self._h:
            _lib.orc_index_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def info(self):
        a = np.zeros(16, dtype=np.uint64)
        _check(_lib.orc_index_info(self._h, _ptr(a)))
        sigma = int(a[0])
        return dict(sigma=sigma, n_rows=int(a[1]), n_ssa=int(a[2]), sampling_rate=int(a[3]), bits_for_position=int(a[4]),
                    C=[int(x) for x in a[5:6 + sigma]])

    def all_ranks(self, which, positions):
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        out = np.zeros((pos.size, self.info()["sigma"]), dtype=np.uint64)
        _check(_lib.orc_all_ranks(self._h, which, _ptr(pos), pos.size, _ptr(out)))
        return out

    def bwt(self, which):
        out = np.zeros(self.info()["n_rows"], dtype=np.uint8)
        _check(_lib.orc_bwt_symbols(self._h, which, _ptr(out)))
        return out

    def locate_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros((rows.size, 3), dtype=np.uint64)
        _check(_lib.orc_locate_rows(self._h, _ptr(rows), rows.size, _ptr(out)))
        return out

    def search(self, queries, scheme, edit, threads=1, max_hits=0):
        """scheme: object with pi (uint16 [S, m]), l, u (uint8).  -> cursors uint64 [n, 4] in reference order.
        max_hits > 0: search_n (a query ends after max_hits suffix-array rows)."""
        q = np.ascontiguousarray(queries, dtype=np.uint8)
        assert q.ndim == 2 and q.shape[1] == scheme.pi.shape[1]
        p, n = C.c_void_p(), C.c_uint64()
        _check(_lib.orc_search_n(self._h, _ptr(q), q.shape[0], q.shape[1], scheme.pi.shape[0], _ptr(scheme.pi), _ptr(scheme.l),
                                 _ptr(scheme.u), int(edit), threads, int(max_hits), C.byref(p), C.byref(n), _ptr(self.counters)))
        try:
            if n.value == 0:
                return np.zeros((0, 4), dtype=np.uint64)
            return np.ctypeslib.as_array(C.cast(p, u64p), shape=(n.value * 4,)).copy().reshape(-1, 4)
        finally:
            _lib.orc_free(p)

    def locate(self, cursors, threads=1):
        cur = np.ascontiguousarray(cursors, dtype=np.uint64).reshape(-1, 4)
        p, n = C.c_void_p(), C.c_uint64()
        _check(_lib.orc_locate(self._h, _ptr(cur), cur.shape[0], threads, C.byref(p), C.byref(n), _ptr(self.counters)))
        try:
            if n.value == 0:
                return np.zeros((0, 4), dtype=np.uint64)
            return np.ctypeslib.as_array(C.cast(p, u64p), shape=(n.value * 4,)).copy().reshape(-1, 4)
        finally:
            _lib.orc_free(p)


def sort_rows(a):
    """lexicographic sort of the rows of an [n, 4] array"""
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    if a.shape[0] == 0:
        return a
    order = np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
    return a[order]


def bf_hamming(seq, query, k):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    query = np.ascontiguousarray(query, dtype=np.uint8)
    cap = max(16, seq.size)
    out = np.zeros((cap, 2), dtype=np.uint64)
    n = _lib.orc_bf_hamming(_ptr(seq), seq.size, _ptr(query), query.size, k, _ptr(out), cap)
    return out[:n]


def bf_edit_starts(seq, query, k):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    query = np.ascontiguousarray(query, dtype=np.uint8)
    out = np.zeros(seq.size + 1, dtype=np.uint8)
    _lib.orc_bf_edit_starts(_ptr(seq), seq.size, _ptr(query), query.size, k, _ptr(out))
    return out
