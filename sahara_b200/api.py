"""Python mirror of the reference-facing interface of the hot path.

Names follow /root/reference/src/sahara/search.cpp: a `SearchScheme` is what
`fmc::search_scheme::generator::all[name].generator(minK, maxK, 0, 0)` + `expand(scheme, len)` (+
`limitToHamming`) produce (search.cpp:174-212, 226); `Context.search` is
`fmc::search_ng24::search<Edit>` followed by the `LocateLinear` loop (search.cpp:227-250).
Every call goes through the C ABI (include/sahara_b200.h); nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import check, check_host, cuda, host


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class SearchScheme:
    """Tables pi/l/u of shape [n_searches, n_entries] (n_entries = parts, or query length once expanded)."""

    def __init__(self, pi, l, u):
        self.pi = np.ascontiguousarray(pi, dtype=np.uint16)
        self.l = np.ascontiguousarray(l, dtype=np.uint8)
        self.u = np.ascontiguousarray(u, dtype=np.uint8)
        assert self.pi.shape == self.l.shape == self.u.shape and self.pi.ndim == 2

    @property
    def n_searches(self):
        return self.pi.shape[0]

    @property
    def n_entries(self):
        return self.pi.shape[1]

    @staticmethod
    def names():
        return host.sbh_scheme_names().decode().split(",")

    @staticmethod
    def _take(ns, ne, ppi, pl, pu):
        S, E = ns.value, ne.value
        try:
            pi = np.ctypeslib.as_array(C.cast(ppi, N.u16p), shape=(S * E,)).copy().reshape(S, E)
            l = np.ctypeslib.as_array(C.cast(pl, N.u8p), shape=(S * E,)).copy().reshape(S, E)
            u = np.ctypeslib.as_array(C.cast(pu, N.u8p), shape=(S * E,)).copy().reshape(S, E)
        finally:
            for p in (ppi, pl, pu):
                host.sbh_free(p)
        return SearchScheme(pi, l, u)

    @staticmethod
    def generate(name, min_k, max_k, length=0, limit_to_hamming=False):
        ns, ne = C.c_uint32(), C.c_uint32()
        ppi, pl, pu = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check_host(host.sbh_scheme_generate(name.encode(), min_k, max_k, length, int(limit_to_hamming), C.byref(ns), C.byref(ne),
                                            C.byref(ppi), C.byref(pl), C.byref(pu)))
        return SearchScheme._take(ns, ne, ppi, pl, pu)

    @staticmethod
    def generate_dynamic(name, min_k, max_k, length, edit, sigma, ref_len):
        """`--dynamic_generator`: part sizes by weighted node count -> (scheme, partition)"""
        ns, ne, npart = C.c_uint32(), C.c_uint32(), C.c_uint32()
        ppi, pl, pu = C.c_void_p(), C.c_void_p(), C.c_void_p()
        part = (C.c_uint32 * 64)()
        check_host(host.sbh_scheme_generate_dynamic(name.encode(), min_k, max_k, length, int(edit), sigma, ref_len, C.byref(ns),
                                                    C.byref(ne), C.byref(ppi), C.byref(pl), C.byref(pu), part, 64, C.byref(npart)))
        return SearchScheme._take(ns, ne, ppi, pl, pu), [int(part[i]) for i in range(min(64, npart.value))]

    @staticmethod
    def from_columba(text, length=0, limit_to_hamming=False):
        ns, ne = C.c_uint32(), C.c_uint32()
        ppi, pl, pu = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check_host(host.sbh_scheme_from_columba(text.encode(), length, int(limit_to_hamming), C.byref(ns), C.byref(ne),
                                                C.byref(ppi), C.byref(pl), C.byref(pu)))
        return SearchScheme._take(ns, ne, ppi, pl, pu)

    def to_columba(self):
        rows = []
        for j in range(self.n_searches):
            rows.append(" ".join("{" + ",".join(str(int(x)) for x in t[j]) + "}" for t in (self.pi, self.l, self.u)))
        return "\n".join(rows) + "\n"

    def check(self, min_k, max_k):
        v, c, n = C.c_int(), C.c_int(), C.c_int()
        check_host(host.sbh_scheme_check(self.n_searches, self.n_entries, _ptr(self.pi), _ptr(self.l), _ptr(self.u), min_k, max_k,
                                         C.byref(v), C.byref(c), C.byref(n)))
        return bool(v.value), bool(c.value), bool(n.value)

    def node_count(self, edit, sigma, ref_len):
        a, b = C.c_double(), C.c_double()
        check_host(host.sbh_scheme_node_count(self.n_searches, self.n_entries, _ptr(self.pi), _ptr(self.l), _ptr(self.u), int(edit),
                                              sigma, ref_len, C.byref(a), C.byref(b)))
        return a.value, b.value


def load_fasta_reads(path, sigma=6, threads=4):
    """read set of `sahara search` -> uint8 array [n_reads, len] of ranks (parallel reader of the CLI)"""
    pr, n, ln = C.c_void_p(), C.c_uint64(), C.c_uint64()
    check_host(host.sbh_fasta_load_reads(str(path).encode(), sigma, threads, C.byref(pr), C.byref(n), C.byref(ln)))
    try:
        total = n.value * ln.value
        a = np.ctypeslib.as_array(C.cast(pr, C.POINTER(C.c_uint8)), shape=(max(1, total),))[:total].copy()
    finally:
        host.sbh_free(pr)
    return a.reshape(n.value, ln.value)


def load_fasta_ranks(path, sigma=6, with_revcomp=False):
    """FASTA -> list of uint8 rank arrays (0='$', 1..4=ACGT, 5=N)."""
    pr, pl, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    check_host(host.sbh_fasta_load_ranks(str(path).encode(), sigma, int(with_revcomp), C.byref(pr), C.byref(pl), C.byref(n)))
    try:
        lens = np.ctypeslib.as_array(C.cast(pl, N.u64p), shape=(max(1, n.value),))[: n.value].copy()
        total = int(lens.sum())
        ranks = np.ctypeslib.as_array(C.cast(pr, N.u8p), shape=(max(1, total),))[:total].copy()
    finally:
        host.sbh_free(pr)
        host.sbh_free(pl)
    out, o = [], 0
    for ln in lens:
        out.append(ranks[o:o + int(ln)])
        o += int(ln)
    return out


def pack_reads4(reads, threads=4, out=None):
    """ranks [n_reads, len] -> SB200_READS_PACKED4 words [n_reads, (len + 7) // 8] (uint32)"""
    r = np.ascontiguousarray(reads, dtype=np.uint8)
    if r.ndim != 2:
        raise ValueError("reads must be a dense [n_reads, length] array of ranks")
    W = (r.shape[1] + 7) // 8
    if out is None:
        out = np.empty((r.shape[0], W), dtype=np.uint32)
    check_host(host.sbh_pack_reads4(_ptr(r), r.shape[0], r.shape[1], threads, C.c_void_p(out.ctypes.data)))
    return out


def pack_reads2(reads, threads=4, out=None):
    """ranks [n_reads, len] of A, C, G, T only -> SB200_READS_PACKED2 words [n_reads, (len + 15) // 16] (uint32); raises
    SaharaError when a read holds another symbol"""
    r = np.ascontiguousarray(reads, dtype=np.uint8)
    if r.ndim != 2:
        raise ValueError("reads must be a dense [n_reads, length] array of ranks")
    W = (r.shape[1] + 15) // 16
    if out is None:
        out = np.empty((r.shape[0], W), dtype=np.uint32)
    check_host(host.sbh_pack_reads2(_ptr(r), r.shape[0], r.shape[1], threads, C.c_void_p(out.ctypes.data)))
    return out


def device_count():
    """CUDA devices the library sees (sb200_device_count)"""
    n = C.c_int()
    check(cuda.sb200_device_count(C.byref(n)))
    return n.value


def default_policy():
    """SB200_POLICY_DEFAULT of include/sahara_policy.h"""
    return N.Policy(del_after=0b1001, ins_after=0b0101, end_ok=0b0101, child_order=0, expand_lower=0)


def set_expand_rule(rule):
    """rule `expand_lower` of the policy table for every SearchScheme.generate() that follows"""
    check_host(host.sbh_set_expand_rule(int(rule)))


def decode_batch(res, first_query=0):
    """sb200_batch_result -> uint64 [n_hits, 4] = (queryId, seqId, pos, errors) in the order of the records
    (fixed or delta-coded records: sbh_decode_records of the host library)"""
    out = np.empty((res.n_hits, 4), dtype=np.uint64)
    if res.n_hits == 0:
        return out
    check_host(host.sbh_decode_records(res.hit_end, res.records, res.n_queries, res.n_hits, res.record_bytes, res.bits_for_position,
                                       int(res.delta_coded), int(first_query), _ptr(out)))
    return out


def revcomp_ranks(r):
    r = np.ascontiguousarray(r, dtype=np.uint8)
    out = np.empty_like(r)
    check_host(host.sbh_revcomp_ranks(_ptr(r), r.size, _ptr(out)))
    return out


class Context:
    """One GPU: index + scheme + work buffers (sb200_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(cuda.sb200_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            cuda.sb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- index ----
    def load_index(self, path):
        view, handle = N.IndexView(), C.c_void_p()
        check_host(host.sbh_idx_load(str(path).encode(), C.byref(view), C.byref(handle)))
        try:
            check(cuda.sb200_index_upload(self._h, C.byref(view)))
        finally:
            host.sbh_idx_free(handle)

    def upload_view(self, view):
        check(cuda.sb200_index_upload(self._h, C.byref(view)))

    def build_index(self, seqs, sigma=6, sampling_rate=16):
        seqs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        lens = np.array([s.size for s in seqs], dtype=np.uint64)
        cat = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
        cat = np.ascontiguousarray(cat)
        check(cuda.sb200_index_build(self._h, _ptr(cat), _ptr(lens), len(seqs), sigma, sampling_rate))

    def build_index_device(self, d_ptr, lens, sigma=6, sampling_rate=16):
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        check(cuda.sb200_index_build_device(self._h, C.c_void_p(d_ptr), _ptr(lens), lens.size, sigma, sampling_rate))

    def download_view(self):
        view = N.IndexView()
        check(cuda.sb200_index_download(self._h, C.byref(view)))
        return view

    def free_view(self, view):
        cuda.sb200_index_view_free(C.byref(view))

    def save_index(self, path):
        view = self.download_view()
        try:
            check_host(host.sbh_idx_save(str(path).encode(), C.byref(view)))
        finally:
            self.free_view(view)

    def info(self):
        i = N.IndexInfo()
        check(cuda.sb200_index_info_get(self._h, C.byref(i)))
        return {k: (list(getattr(i, k)) if k == "C" else getattr(i, k)) for k, _ in N.IndexInfo._fields_}

    def densify(self, rate):
        check(cuda.sb200_index_densify(self._h, rate))

    def enable_text(self, enable=True):
        """in-text verification for cursors that hold a single row (results unchanged)"""
        check(cuda.sb200_index_enable_text(self._h, int(enable)))

    def build_qgram(self, q):
        check(cuda.sb200_index_build_qgram(self._h, q))

    # ---- scheme ----
    def set_scheme(self, scheme, edit):
        check(cuda.sb200_set_scheme(self._h, scheme.n_searches, scheme.n_entries, _ptr(scheme.pi), _ptr(scheme.l), _ptr(scheme.u),
                                    int(edit)))

    def set_policy(self, policy=None):
        """replaces the table of reconstructed rules (include/sahara_policy.h); None = the default"""
        p = policy if policy is not None else default_policy()
        check(cuda.sb200_set_policy(self._h, C.byref(p)))

    def get_policy(self):
        p = N.Policy()
        check(cuda.sb200_get_policy(self._h, C.byref(p)))
        return p

    def clone_index_from(self, other):
        """replicates the index of `other` (a Context on another GPU, or the same one) with GPU-to-GPU copies (sb200_index_clone)"""
        check(cuda.sb200_index_clone(self._h, other._h))

    def set_option(self, name, value):
        """knobs of the host orchestration (sb200_set_option); results never depend on them"""
        check(cuda.sb200_set_option(self._h, name.encode(), int(value)))

    def set_max_hits(self, max_hits):
        """search_n (src/sahara/search.cpp:228,231): at most max_hits rows per query, the first ones in the order
        of the reference recursion; 0 = unlimited"""
        check(cuda.sb200_set_max_hits(self._h, int(max_hits)))

    # ---- search ----
    @staticmethod
    def _take4(p, n):
        try:
            if n.value == 0:
                return np.zeros((0, 4), dtype=np.uint64)
            return np.ctypeslib.as_array(C.cast(p, N.u64p), shape=(n.value * 4,)).copy().reshape(-1, 4)
        finally:
            cuda.sb200_free(p)

    @staticmethod
    def _queries(q):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        if q.ndim != 2:
            raise ValueError("queries must be a dense [n_queries, length] array of ranks")
        return q

    def search(self, queries):
        """-> uint64 [n_hits, 4] = (queryId, seqId, pos, errors), sorted."""
        q = self._queries(queries)
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search(self._h, _ptr(q), q.shape[0], q.shape[1], C.byref(p), C.byref(n)))
        return self._take4(p, n)

    def search_reads(self, reads, with_reverse=True):
        """forward reads only -> uint32 [n_hits, 4] = (queryId, seqId, pos, errors); reverse complements are made on
        the device, query ids count both strands like the reference (2i, 2i+1)."""
        r = self._queries(reads)
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search_reads(self._h, _ptr(r), r.shape[0], r.shape[1], int(with_reverse), C.byref(p), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros((0, 4), dtype=np.uint32)
            return np.ctypeslib.as_array(C.cast(p, N.u32p), shape=(n.value * 4,)).copy().reshape(-1, 4)
        finally:
            cuda.sb200_free(p)

    # ---- asynchronous batches (sb200_submit_reads / sb200_wait_batch / sb200_release_batch) ----
    def submit_reads(self, reads, length=None, packed4=False, with_reverse=True, packed2=False):
        """reads: ranks [n, len] (uint8), or with packed4 / packed2 the words of pack_reads4 / pack_reads2 (uint32) — or an
        int address of such a (page-locked) buffer together with (n_reads, length) in `length`.  Returns a ticket; the
        buffer must stay alive until wait_batch."""
        if isinstance(reads, tuple):
            addr, n_reads, ln = reads
        else:
            a = np.ascontiguousarray(reads)
            n_reads = a.shape[0]
            ln = length if length is not None else a.shape[1]
            addr = a.ctypes.data
            self._keep = getattr(self, "_keep", {})
        t = C.c_uint64()
        check(cuda.sb200_submit_reads(self._h, C.c_void_p(addr), n_reads, ln, 2 if packed2 else 1 if packed4 else 0, int(with_reverse), C.byref(t)))
        if not isinstance(reads, tuple):
            self._keep[t.value] = a
        return t.value

    def submit_device(self, d_queries, n_queries, length):
        t = C.c_uint64()
        check(cuda.sb200_submit_device(self._h, C.c_void_p(d_queries), n_queries, length, C.byref(t)))
        return t.value

    def wait_batch(self, ticket, copy_to_host=True):
        res = N.BatchResult()
        check(cuda.sb200_wait_batch(self._h, ticket, int(copy_to_host), C.byref(res)))
        return res

    def release_batch(self, ticket):
        check(cuda.sb200_release_batch(self._h, ticket))
        getattr(self, "_keep", {}).pop(ticket, None)

    def search_reads_async(self, reads, packed4=False, with_reverse=True, batch=None, depth=2, packed2=False):
        """the reads cut into batches that are submitted `depth` deep -> uint64 [n_hits, 4] like search()"""
        r = self._queries(reads)
        n, m = r.shape
        batch = batch or n
        per = 2 if with_reverse else 1
        pieces = [(o, min(n, o + batch)) for o in range(0, n, batch)]
        bufs = [pack_reads2(r[a:b]) if packed2 else pack_reads4(r[a:b]) if packed4 else np.ascontiguousarray(r[a:b]) for a, b in pieces]
        out, tickets = [], []
        for i in range(len(pieces) + depth):
            if i >= depth:
                j = i - depth
                res = self.wait_batch(tickets[j])
                out.append(decode_batch(res, first_query=pieces[j][0] * per))
                self.release_batch(tickets[j])
            if i < len(pieces):
                tickets.append(self.submit_reads(bufs[i], length=m, packed4=packed4, with_reverse=with_reverse, packed2=packed2))
        return np.concatenate(out) if out else np.zeros((0, 4), dtype=np.uint64)

    def search_cursors(self, queries):
        """-> uint64 [n, 4] = (queryId, lb, len, errors), sorted."""
        q = self._queries(queries)
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_search_cursors(self._h, _ptr(q), q.shape[0], q.shape[1], C.byref(p), C.byref(n)))
        return self._take4(p, n)

    def locate(self, cursors):
        cur = np.ascontiguousarray(cursors, dtype=np.uint64).reshape(-1, 4)
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_locate(self._h, _ptr(cur), cur.shape[0], C.byref(p), C.byref(n)))
        return self._take4(p, n)

    def search_device(self, d_queries, n_queries, length, locate=True):
        nc, nh = C.c_uint64(), C.c_uint64()
        check(cuda.sb200_search_device(self._h, C.c_void_p(d_queries), n_queries, length, C.byref(nc),
                                       C.byref(nh) if locate else None))
        return nc.value, nh.value

    def fetch_hits(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(cuda.sb200_fetch_hits(self._h, C.byref(p), C.byref(n)))
        return self._take4(p, n)

    # ---- rank ----
    def rank_probe(self, which, positions):
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        sigma = self.info()["sigma"]
        out = np.zeros((pos.size, sigma), dtype=np.uint64)
        check(cuda.sb200_rank_probe(self._h, which, _ptr(pos), pos.size, _ptr(out)))
        return out

    def rank_bench(self, which, n_chains, iters, seed=1):
        ms, cs = C.c_float(), C.c_uint64()
        check(cuda.sb200_rank_bench(self._h, which, n_chains, iters, seed, C.byref(ms), C.byref(cs)))
        return ms.value, cs.value

    # ---- misc ----
    def counters(self):
        c = N.Counters()
        check(cuda.sb200_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in N.Counters._fields_}

    def reset_counters(self):
        check(cuda.sb200_reset_counters(self._h))

    def set_stream(self, stream_ptr):
        check(cuda.sb200_set_stream(self._h, C.c_void_p(stream_ptr)))

    def synchronize(self):
        check(cuda.sb200_synchronize(self._h))

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        check(cuda.sb200_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, d_ptr):
        check(cuda.sb200_device_free(self._h, C.c_void_p(d_ptr)))

    def to_host(self, d_ptr, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        check(cuda.sb200_copy_to_host(self._h, _ptr(out), C.c_void_p(d_ptr), nbytes))
        return out

    def to_device(self, d_ptr, arr):
        arr = np.ascontiguousarray(arr)
        check(cuda.sb200_copy_to_device(self._h, C.c_void_p(d_ptr), _ptr(arr), arr.nbytes))

    def synth_genome(self, n_bases, seed):
        d = self.device_alloc(n_bases + 64)
        check(cuda.sb200_synth_genome_device(self._h, n_bases, seed, C.c_void_p(d)))
        return d

    def synth_reads(self, d_genome, n_bases, n_reads, length, k, edit, seed, first_read=0, d_out=None):
        if d_out is None:
            d_out = self.device_alloc(2 * n_reads * length)
        check(cuda.sb200_synth_reads_device(self._h, C.c_void_p(d_genome), n_bases, n_reads, length, k, int(edit), seed, first_read,
                                            C.c_void_p(d_out)))
        return d_out
