"""ctypes wrapper of tools/libss_synth_host.so (tools/synth_host.cpp): the synthetic workload generated on host
threads, byte-identical to the device generators, without a GPU and without loading the product library.
Used by bench.py --impl reference (so that arm never touches libstrainscan_b200.so) and by tests."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libss_synth_host.so")
_lib = None


def build():
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", LIB_PATH,
                           os.path.join(_HERE, "synth_host.cpp")])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.ssh_last_error.restype = C.c_char_p
        lib.ssh_read_record_bytes.restype = C.c_size_t
        lib.ssh_db_record_bytes.restype = C.c_size_t
        lib.ssh_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
        lib.ssh_db.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(load().ssh_last_error().decode())


def read_record_bytes(params):
    return int(load().ssh_read_record_bytes(C.byref(params)))


def db_record_bytes(params):
    return int(load().ssh_db_record_bytes(C.byref(params)))


def reads(params, n_reads, first_read=0, threads=None, out=None):
    """uint8 array of 4-line FASTQ text for reads [first_read, first_read + n_reads)."""
    n = int(n_reads) * read_record_bytes(params)
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags.c_contiguous
    _check(load().ssh_reads(C.addressof(params), out.ctypes.data, int(n_reads), int(first_read),
                            int(threads or os.cpu_count() or 1)))
    return out[:n]


def db(params, node_sizes, want_nodes=True, threads=None):
    """(kmer.fa text as uint8, owner node of every record or None)."""
    node_sizes = np.ascontiguousarray(node_sizes, dtype=np.uint32)
    n = int(node_sizes.sum())
    text = np.empty(n * db_record_bytes(params), dtype=np.uint8)
    node_of = np.empty(n, dtype=np.uint32) if want_nodes else None
    _check(load().ssh_db(C.addressof(params), node_sizes.ctypes.data, node_sizes.size, text.ctypes.data,
                         node_of.ctypes.data if want_nodes else None, int(threads or os.cpu_count() or 1)))
    return text, node_of
