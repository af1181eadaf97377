"""ctypes binding of libstrainscan_b200.so (C ABI in include/strainscan_b200.h).

The shared object is built in-tree by strainscan_b200/csrc/Makefile (see __graft_entry__.build()).
There is no fallback: if the library is missing, or no sm_100 GPU is present at Engine() time, the
caller gets an exception -- never a CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SS_LIB_PATH: another build of the SAME library (tools/build_variant.sh, A/B runs of compile-time knobs); never a fallback
LIB_PATH = os.environ.get("SS_LIB_PATH") or os.path.join(_HERE, "libstrainscan_b200.so")

SS_OK = 0
SS_ERR_NO_DEVICE, SS_ERR_CUDA, SS_ERR_IO, SS_ERR_FORMAT, SS_ERR_ARG, SS_ERR_NOMEM, SS_ERR_UNSUPPORTED = range(1, 8)
SS_REC_IN_SET, SS_REC_IS_LAST, SS_REC_RAW_UPPER = 1, 2, 4
SS_SYNTH_MAX_SOURCES = 8


class StrainScanB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[ss error %d] %s" % (code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("text_bytes", C.c_uint64), ("n_reads", C.c_uint64), ("n_kmers", C.c_uint64),
                ("n_hits", C.c_uint64), ("n_second_probe", C.c_uint64),
                ("ms_index", C.c_double), ("ms_probe", C.c_double), ("ms_gather", C.c_double),
                ("ms_h2d", C.c_double), ("ms_total", C.c_double),
                ("probe_launches", C.c_uint32), ("total_launches", C.c_uint32),
                ("n_table_probes", C.c_uint64), ("binned_rounds", C.c_uint32), ("bins", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_leaves", C.c_uint32), ("genome_len", C.c_uint32),
                ("block_len", C.c_uint32), ("k", C.c_uint32), ("read_len", C.c_uint32),
                ("header_len", C.c_uint32), ("n_sources", C.c_uint32),
                ("source_leaf", C.c_uint32 * SS_SYNTH_MAX_SOURCES),
                ("source_strain", C.c_uint32 * SS_SYNTH_MAX_SOURCES),
                ("source_cum", C.c_uint32 * SS_SYNTH_MAX_SOURCES),
                ("p_offtarget", C.c_uint32), ("p_sub", C.c_uint32), ("p_n", C.c_uint32),
                ("snp_rate", C.c_uint32)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_CSTRS = C.POINTER(C.c_char_p)
_SIZES = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); every symbol include/strainscan_b200.h declares
SIGNATURES = {
    "ss_init": (C.c_int, [C.c_int, _PP]),
    "ss_shutdown": (C.c_int, [_P]),
    "ss_last_error": (C.c_char_p, []),
    "ss_set_stream": (C.c_int, [_P, _P]),
    "ss_device_info": (C.c_int, [_P, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "ss_kmerset_from_fasta": (C.c_int, [_P, C.c_char_p, C.c_int, _PP]),
    "ss_kmerset_from_text": (C.c_int, [_P, C.c_char_p, C.c_size_t, C.c_int, _PP]),
    "ss_kmerset_from_fasta_cached": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int), _PP]),
    "ss_kmerset_free": (C.c_int, [_P]),
    "ss_kmerset_records": (C.c_uint64, [_P]),
    "ss_kmerset_distinct": (C.c_uint64, [_P]),
    "ss_kmerset_k": (C.c_int, [_P]),
    "ss_kmerset_table_bytes": (C.c_uint64, [_P]),
    "ss_kmerset_flags": (C.c_int, [_P, _P]),
    "ss_kmerset_header_ids": (C.c_int, [_P, _P]),
    "ss_reads_from_files": (C.c_int, [_P, _CSTRS, C.c_int, C.c_int, C.c_int, _PP]),
    "ss_fastq_shard_range": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_int, _SIZES, _SIZES]),
    "ss_ingest_files_host": (C.c_int, [_CSTRS, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, _P, C.c_size_t, _SIZES,
                                       C.POINTER(C.c_uint32)]),
    "ss_dgz_inflate_host": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, _P,
                                      C.c_size_t, _SIZES, _SIZES, C.POINTER(C.c_uint64)]),
    "ss_dense_nonzero_lists": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ss_node_lists_parse": (C.c_int, [_CSTRS, C.c_uint32, C.c_int, _P, _P, C.c_uint64, _P]),
    "ss_dgz_plan_host": (C.c_int, [C.c_size_t, C.c_int, C.c_double, C.POINTER(C.c_uint64)]),
    "ss_dgz_tables_selftest_host": (C.c_int, [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "ss_reads_from_host": (C.c_int, [_P, _CSTRS, _SIZES, C.c_int, _PP]),
    "ss_reads_from_device": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _PP]),
    "ss_reads_device_capacity": (C.c_size_t, [C.c_size_t]),
    "ss_reads_bytes": (C.c_uint64, [_P]),
    "ss_reads_free": (C.c_int, [_P]),
    "ss_reads_drop_index": (C.c_int, [_P]),
    "ss_count": (C.c_int, [_P, _P, _P, _P, C.POINTER(Stats)]),
    "ss_count_device": (C.c_int, [_P, _P, _P, _P, C.POINTER(Stats)]),
    "ss_count_host": (C.c_int, [_P, _P, C.POINTER(C.c_void_p), _SIZES, C.c_int, _P, C.POINTER(Stats)]),
    "ss_count_files": (C.c_int, [_P, _P, _CSTRS, C.c_int, C.c_int, C.c_int, _P, C.POINTER(Stats)]),
    "ss_l2_finalize": (C.c_int, [_P, _P, _P, _P]),
    "ss_node_reduce": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, _P, _P, _P]),
    "ss_strain_reduce": (C.c_int, [_P, _P, _P, C.c_uint32, _P, _P, C.c_uint64, _P, _P, _P]),
    "ss_node_index_create": (C.c_int, [_P, _P, _P, C.c_uint32, _PP]),
    "ss_node_index_free": (C.c_int, [_P]),
    "ss_node_index_reduce": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "ss_strain_matrix_create": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint64, _PP]),
    "ss_strain_matrix_free": (C.c_int, [_P]),
    "ss_strain_matrix_reduce": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "ss_bench_random_gather": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "ss_synth_read_record_bytes": (C.c_size_t, [C.POINTER(SynthParams)]),
    "ss_synth_db_record_bytes": (C.c_size_t, [C.POINTER(SynthParams)]),
    "ss_synth_reads_device": (C.c_int, [_P, C.POINTER(SynthParams), _P, C.c_uint64, C.c_uint64]),
    "ss_synth_db_host": (C.c_int, [_P, C.POINTER(SynthParams), _P, C.c_uint32, _P, _P]),
}

_lib = None


def load():
    """Load the shared object and bind every symbol; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "strainscan_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C strainscan_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != SS_OK:
        msg = load().ss_last_error()
        raise StrainScanB200Error(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
