"""ctypes bindings of the two in-tree shared libraries.

libsahara_b200.so  — the CUDA C ABI (include/sahara_b200.h); no CPU fallback exists.
libsahara_host.so  — host-side helpers (include/sahara_host.h): schemes, index file, FASTA.

Importing this module fails loudly when a library is missing: build with `make` or
`python -c "import __graft_entry__ as g; g.build()"`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    # SB200_LIB_DIR: load a differently configured build of the same library (kernel tuning experiments)
    path = os.path.join(os.environ.get("SB200_LIB_DIR", _HERE), name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension must be built (make, or __graft_entry__.build()); "
            "sahara_b200 has no CPU fallback")
    return C.CDLL(path)


u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class IndexView(C.Structure):
    _fields_ = [
        ("sigma", C.c_uint64), ("n_rows", C.c_uint64), ("n_blocks", C.c_uint64),
        ("bwt_blocks", C.c_void_p), ("bwt_super", C.c_void_p),
        ("bwtrev_blocks", C.c_void_p), ("bwtrev_super", C.c_void_p),
        ("C", C.c_void_p), ("ssa", C.c_void_p), ("n_ssa", C.c_uint64),
        ("mark_bits", C.c_void_p), ("sampling_rate", C.c_uint64), ("bits_for_position", C.c_uint64),
    ]


class IndexInfo(C.Structure):
    _fields_ = [
        ("sigma", C.c_uint64), ("n_rows", C.c_uint64), ("n_ssa", C.c_uint64), ("sampling_rate", C.c_uint64),
        ("bits_for_position", C.c_uint64), ("device_sampling_rate", C.c_uint64), ("device_bytes", C.c_uint64),
        ("C", C.c_uint64 * 8),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("nodes", C.c_uint64), ("rank_ops", C.c_uint64), ("cursors", C.c_uint64), ("lf_steps", C.c_uint64),
        ("hits", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_search", C.c_float), ("ms_locate", C.c_float), ("ms_sort", C.c_float), ("ms_h2d", C.c_float),
        ("ms_d2h", C.c_float), ("ms_fm", C.c_float), ("ms_text", C.c_float), ("nodes_text", C.c_uint64),
        ("batch_restarts", C.c_uint64),
    ]


class Policy(C.Structure):
    """sb200_policy (include/sahara_policy.h): the reconstructed rules of the recursion"""
    _fields_ = [
        ("del_after", C.c_uint32), ("ins_after", C.c_uint32), ("end_ok", C.c_uint32), ("child_order", C.c_uint32),
        ("expand_lower", C.c_uint32), ("reserved", C.c_uint32 * 3),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("n_queries", C.c_uint64), ("n_hits", C.c_uint64), ("n_cursors", C.c_uint64),
        ("hit_end", C.c_void_p), ("records", C.c_void_p),
        ("record_bytes", C.c_uint32), ("bits_for_position", C.c_uint32),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("ms_search", C.c_float), ("ms_locate", C.c_float), ("ms_sort", C.c_float),
        ("delta_coded", C.c_uint32), ("n_record_bytes", C.c_uint64),
    ]


# every symbol include/sahara_b200.h declares: name -> (restype, argtypes)
SB200_SYMBOLS = {
    "sb200_abi_version": (C.c_int, []),
    "sb200_last_error": (C.c_char_p, []),
    "sb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sb200_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sb200_destroy": (C.c_int, [C.c_void_p]),
    "sb200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb200_synchronize": (C.c_int, [C.c_void_p]),
    "sb200_index_upload": (C.c_int, [C.c_void_p, C.POINTER(IndexView)]),
    "sb200_index_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]),
    "sb200_index_build_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]),
    "sb200_index_download": (C.c_int, [C.c_void_p, C.POINTER(IndexView)]),
    "sb200_index_view_free": (None, [C.POINTER(IndexView)]),
    "sb200_index_info_get": (C.c_int, [C.c_void_p, C.POINTER(IndexInfo)]),
    "sb200_index_densify": (C.c_int, [C.c_void_p, C.c_uint32]),
    "sb200_index_build_qgram": (C.c_int, [C.c_void_p, C.c_uint32]),
    "sb200_index_enable_text": (C.c_int, [C.c_void_p, C.c_int]),
    "sb200_set_scheme": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "sb200_set_max_hits": (C.c_int, [C.c_void_p, C.c_uint64]),
    "sb200_set_policy": (C.c_int, [C.c_void_p, C.POINTER(Policy)]),
    "sb200_get_policy": (C.c_int, [C.c_void_p, C.POINTER(Policy)]),
    "sb200_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sb200_index_clone": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb200_submit_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, u64p]),
    "sb200_submit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, u64p]),
    "sb200_wait_batch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(BatchResult)]),
    "sb200_release_batch": (C.c_int, [C.c_void_p, C.c_uint64]),
    "sb200_search_cursors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p), u64p]),
    "sb200_locate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), u64p]),
    "sb200_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p), u64p]),
    "sb200_search_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), u64p]),
    "sb200_search_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, u64p, u64p]),
    "sb200_fetch_hits": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), u64p]),
    "sb200_free": (None, [C.c_void_p]),
    "sb200_rank_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]),
    "sb200_rank_bench": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_float), u64p]),
    "sb200_get_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "sb200_reset_counters": (C.c_int, [C.c_void_p]),
    "sb200_synth_genome_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "sb200_synth_reads_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                           C.c_uint64, C.c_uint64, C.c_void_p]),
    "sb200_device_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "sb200_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb200_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "sb200_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "sb200_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
}

SBH_SYMBOLS = {
    "sbh_last_error": (C.c_char_p, []),
    "sbh_scheme_names": (C.c_char_p, []),
    "sbh_scheme_generate": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_int, u32p, u32p, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "sbh_scheme_generate_dynamic": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, u32p, u32p,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_uint32,
                                              u32p]),
    "sbh_scheme_from_columba": (C.c_int, [C.c_char_p, C.c_uint32, C.c_int, u32p, u32p, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "sbh_scheme_check": (C.c_int, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sbh_scheme_node_count": (C.c_int, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sbh_idx_peek_sigma": (C.c_int, [C.c_char_p, u64p]),
    "sbh_idx_load": (C.c_int, [C.c_char_p, C.POINTER(IndexView), C.POINTER(C.c_void_p)]),
    "sbh_idx_free": (None, [C.c_void_p]),
    "sbh_idx_save": (C.c_int, [C.c_char_p, C.POINTER(IndexView)]),
    "sbh_fasta_load_ranks": (C.c_int, [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p]),
    "sbh_fasta_load_reads": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p), u64p, u64p]),
    "sbh_revcomp_ranks": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "sbh_pack_reads4": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sbh_pack_reads2": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sbh_set_expand_rule": (C.c_int, [C.c_uint32]),
    "sbh_decode_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_void_p]),
    "sbh_free": (None, [C.c_void_p]),
}


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib


cuda = _bind(_load("libsahara_b200.so"), SB200_SYMBOLS)
host = _bind(_load("libsahara_host.so"), SBH_SYMBOLS)


class SaharaError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise SaharaError(cuda.sb200_last_error().decode(errors="replace"))


def check_host(rc):
    if rc != 0:
        raise SaharaError(host.sbh_last_error().decode(errors="replace"))
