"""Thin object layer over the C ABI: Engine (one GPU), KmerSet (probe table), Reads (read cache).

PyTorch is used by callers only to own device buffers (tensor.data_ptr()) and for the NCCL plumbing
in strainscan_b200/dist.py; nothing here computes with torch.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Stats, check


def _cstr_array(paths):
    arr = (C.c_char_p * max(len(paths), 1))()
    for i, p in enumerate(paths):
        arr[i] = p.encode() if isinstance(p, str) else p
    return arr


def _as_bytes_list(bufs):
    if isinstance(bufs, (bytes, bytearray, memoryview)):
        bufs = [bufs]
    return [bytes(b) if not isinstance(b, bytes) else b for b in bufs]


class Engine:
    """One CUDA context on one B200.  Fails loudly when there is no sm_100 device."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.ss_init(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ss_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the handle that NAMES the default stream (0 means "the context's own")

    def set_stream(self, cuda_stream_ptr):
        """Launch on the caller's stream (a cudaStream_t handle); 0 = the context's own non-blocking stream."""
        check(self.lib.ss_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr) or None)))

    def follow_torch_stream(self, device=None):
        """Launch on torch's CURRENT stream, so that torch work issued next (an NCCL all-reduce of the count vector,
        a reduction, an event) is ordered after the kernels AND the next pass is ordered after that work.  torch's
        default stream has handle 0, which ss_set_stream reads as "own stream": it is passed as cudaStreamLegacy.
        (With the own stream a pass started right after an all-reduce would overwrite the vector while NCCL still
        reads it -- seen at 8 GPUs with sub-millisecond passes.)"""
        import torch
        h = int(torch.cuda.current_stream(device if device is not None else self.device).cuda_stream)
        self.set_stream(h if h else self.CUDA_STREAM_LEGACY)

    def device_info(self):
        name = C.create_string_buffer(256)
        n_sm = C.c_int()
        mem = C.c_uint64()
        check(self.lib.ss_device_info(self.h, name, 256, C.byref(n_sm), C.byref(mem)))
        return {"name": name.value.decode(), "n_sm": n_sm.value, "mem_bytes": mem.value}

    # ---- k-mer sets -------------------------------------------------------------------------
    def kmerset_from_fasta(self, path, k):
        h = C.c_void_p()
        check(self.lib.ss_kmerset_from_fasta(self.h, path.encode(), int(k), C.byref(h)))
        return KmerSet(self, h)

    def kmerset_from_fasta_cached(self, path, k, cache_path):
        """Binary cache of the parsed records next to the text FASTA; returns (KmerSet, cache_hit)."""
        h = C.c_void_p()
        hit = C.c_int(0)
        check(self.lib.ss_kmerset_from_fasta_cached(self.h, path.encode(), int(k), cache_path.encode(), C.byref(hit), C.byref(h)))
        return KmerSet(self, h), bool(hit.value)

    def kmerset_from_text(self, text, k):
        if isinstance(text, str):
            text = text.encode()
        if isinstance(text, np.ndarray):
            ptr, n = C.cast(text.ctypes.data, C.c_char_p), text.nbytes
        else:
            ptr, n = text, len(text)
        h = C.c_void_p()
        check(self.lib.ss_kmerset_from_text(self.h, ptr, n, int(k), C.byref(h)))
        return KmerSet(self, h)

    # ---- reads ------------------------------------------------------------------------------
    def reads_from_files(self, paths, shard=0, n_shards=1):
        paths = [p for p in ([paths] if isinstance(paths, str) else list(paths)) if p]
        h = C.c_void_p()
        check(self.lib.ss_reads_from_files(self.h, _cstr_array(paths), len(paths), int(shard), int(n_shards),
                                           C.byref(h)))
        return Reads(self, h)

    def reads_from_host(self, bufs):
        bufs = _as_bytes_list(bufs)
        arr = _cstr_array(bufs)
        lens = (C.c_size_t * max(len(bufs), 1))(*[len(b) for b in bufs])
        h = C.c_void_p()
        check(self.lib.ss_reads_from_host(self.h, arr, lens, len(bufs), C.byref(h)))
        return Reads(self, h)

    def reads_device_capacity(self, n_bytes):
        return int(self.lib.ss_reads_device_capacity(int(n_bytes)))

    def reads_from_device(self, dev_ptr, n_bytes, capacity, keepalive=None):
        h = C.c_void_p()
        check(self.lib.ss_reads_from_device(self.h, C.c_void_p(int(dev_ptr)), int(n_bytes), int(capacity), C.byref(h)))
        r = Reads(self, h)
        r._keepalive = keepalive
        return r

    # ---- match + count ----------------------------------------------------------------------
    def count(self, kset, reads, out=None):
        """Dense uint32 vector by record ordinal (host), and the call's Stats."""
        n = kset.n_records
        if out is None:
            out = np.zeros(n, dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size >= n and out.flags.c_contiguous
        st = Stats()
        check(self.lib.ss_count(self.h, kset.h, reads.h, C.c_void_p(out.ctypes.data), C.byref(st)))
        return out, st

    def count_device(self, kset, reads, dev_ptr):
        st = Stats()
        check(self.lib.ss_count_device(self.h, kset.h, reads.h, C.c_void_p(int(dev_ptr)), C.byref(st)))
        return st

    def count_host(self, kset, bufs, out=None, out_ptr=None):
        """End to end from host text.  bufs: bytes objects, or (ptr, nbytes) tuples for pinned memory."""
        n = kset.n_records
        if out_ptr is None:
            if out is None:
                out = np.zeros(n, dtype=np.uint32)
            out_ptr = out.ctypes.data
        if not isinstance(bufs, (list, tuple)) or (len(bufs) == 2 and isinstance(bufs[0], int)):
            bufs = [bufs]
        keep, ptrs, lens = [], [], []
        for b in bufs:
            if isinstance(b, tuple):
                ptrs.append(int(b[0])); lens.append(int(b[1]))
            else:
                b = bytes(b) if not isinstance(b, bytes) else b
                keep.append(b)
                ptrs.append(C.cast(C.c_char_p(b), C.c_void_p).value or 0); lens.append(len(b))
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        ln = (C.c_size_t * max(len(lens), 1))(*lens)
        st = Stats()
        check(self.lib.ss_count_host(self.h, kset.h, arr, ln, len(ptrs), C.c_void_p(int(out_ptr)), C.byref(st)))
        return out, st

    def count_files(self, kset, paths, shard=0, n_shards=1, out=None, out_ptr=None):
        """End to end from files.  The dense vector goes to `out` (host uint32 array) or to `out_ptr` (host or device)."""
        paths = [p for p in ([paths] if isinstance(paths, str) else list(paths)) if p]
        n = kset.n_records
        if out_ptr is None:
            if out is None:
                out = np.zeros(n, dtype=np.uint32)
            out_ptr = out.ctypes.data
        st = Stats()
        check(self.lib.ss_count_files(self.h, kset.h, _cstr_array(paths), len(paths), int(shard), int(n_shards),
                                      C.c_void_p(int(out_ptr)), C.byref(st)))
        return out, st

    def l2_finalize(self, kset, dev_ptr):
        out = np.zeros(kset.n_records, dtype=np.int64)
        check(self.lib.ss_l2_finalize(self.h, kset.h, C.c_void_p(int(dev_ptr)), C.c_void_p(out.ctypes.data)))
        return out

    # ---- reducers ---------------------------------------------------------------------------
    def node_reduce(self, kset, dev_counts_ptr, node_ptr, ordinals):
        node_ptr = np.ascontiguousarray(node_ptr, dtype=np.uint64)
        ordinals = np.ascontiguousarray(ordinals, dtype=np.uint32)
        n = node_ptr.size - 1
        length = np.zeros(n, dtype=np.uint32)
        covered = np.zeros(n, dtype=np.uint32)
        total = np.zeros(n, dtype=np.uint64)
        check(self.lib.ss_node_reduce(self.h, kset.h, C.c_void_p(int(dev_counts_ptr)), C.c_void_p(node_ptr.ctypes.data),
                                      C.c_void_p(ordinals.ctypes.data), n, C.c_void_p(length.ctypes.data),
                                      C.c_void_p(covered.ctypes.data), C.c_void_p(total.ctypes.data)))
        return length, covered, total

    def strain_reduce(self, col_ptr, rows, y, row_mask=None):
        col_ptr = np.ascontiguousarray(col_ptr, dtype=np.uint64)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.int64)
        n = col_ptr.size - 1
        if row_mask is not None:
            row_mask = np.ascontiguousarray(row_mask, dtype=np.uint8)
        total = np.zeros(n, dtype=np.uint64)
        covered = np.zeros(n, dtype=np.uint64)
        ssum = np.zeros(n, dtype=np.uint64)
        check(self.lib.ss_strain_reduce(self.h, C.c_void_p(col_ptr.ctypes.data), C.c_void_p(rows.ctypes.data), n,
                                        C.c_void_p(y.ctypes.data),
                                        C.c_void_p(row_mask.ctypes.data) if row_mask is not None else None,
                                        y.size, C.c_void_p(total.ctypes.data), C.c_void_p(covered.ctypes.data),
                                        C.c_void_p(ssum.ctypes.data)))
        return total, covered, ssum

    # ---- persistent reducer handles ---------------------------------------------------------
    def node_index_create(self, node_ptr, ordinals):
        node_ptr = np.ascontiguousarray(node_ptr, dtype=np.uint64)
        ordinals = np.ascontiguousarray(ordinals, dtype=np.uint32)
        h = C.c_void_p()
        check(self.lib.ss_node_index_create(self.h, C.c_void_p(node_ptr.ctypes.data), C.c_void_p(ordinals.ctypes.data),
                                            node_ptr.size - 1, C.byref(h)))
        return NodeIndex(self, h, node_ptr.size - 1)

    def node_index_reduce(self, kset, index, dev_counts_ptr):
        """(length, covered, sum, max) per node of a NodeIndex over a dense DEVICE count vector (one K4 launch)."""
        n = index.n_nodes
        length, covered = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        total, mx = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32)
        check(self.lib.ss_node_index_reduce(self.h, kset.h, index.h, C.c_void_p(int(dev_counts_ptr)),
                                            C.c_void_p(length.ctypes.data), C.c_void_p(covered.ctypes.data),
                                            C.c_void_p(total.ctypes.data), C.c_void_p(mx.ctypes.data)))
        return length, covered, total, mx

    def strain_matrix_create(self, col_ptr, rows, n_rows):
        col_ptr = np.ascontiguousarray(col_ptr, dtype=np.uint64)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        h = C.c_void_p()
        check(self.lib.ss_strain_matrix_create(self.h, C.c_void_p(col_ptr.ctypes.data), C.c_void_p(rows.ctypes.data),
                                               col_ptr.size - 1, int(n_rows), C.byref(h)))
        return (h, col_ptr.size - 1, int(n_rows))

    def strain_matrix_free(self, handle):
        if self.h and handle[0]:
            self.lib.ss_strain_matrix_free(handle[0])

    def strain_matrix_reduce(self, handle, y, row_mask=None):
        h, n, n_rows = handle
        y = np.ascontiguousarray(y, dtype=np.int64)
        assert y.size == n_rows
        if row_mask is not None:
            row_mask = np.ascontiguousarray(row_mask, dtype=np.uint8)
            assert row_mask.size == n_rows
        total, covered, ssum = (np.zeros(n, dtype=np.uint64) for _ in range(3))
        check(self.lib.ss_strain_matrix_reduce(self.h, h, C.c_void_p(y.ctypes.data),
                                               C.c_void_p(row_mask.ctypes.data) if row_mask is not None else None,
                                               C.c_void_p(total.ctypes.data), C.c_void_p(covered.ctypes.data),
                                               C.c_void_p(ssum.ctypes.data)))
        return total, covered, ssum

    # ---- measurement / synthetic workload ---------------------------------------------------
    def random_gather_gbps(self, n_bytes, n_probes, iters=3):
        g = C.c_double()
        check(self.lib.ss_bench_random_gather(self.h, int(n_bytes), int(n_probes), int(iters), C.byref(g)))
        return g.value

    def synth_read_record_bytes(self, params):
        return int(self.lib.ss_synth_read_record_bytes(C.byref(params)))

    def synth_db_record_bytes(self, params):
        return int(self.lib.ss_synth_db_record_bytes(C.byref(params)))

    def synth_reads_device(self, params, dev_ptr, n_reads, first_read=0):
        check(self.lib.ss_synth_reads_device(self.h, C.byref(params), C.c_void_p(int(dev_ptr)), int(n_reads),
                                             int(first_read)))

    def synth_db_host(self, params, node_sizes, want_nodes=True):
        node_sizes = np.ascontiguousarray(node_sizes, dtype=np.uint32)
        n = int(node_sizes.sum())
        text = np.empty(n * self.synth_db_record_bytes(params), dtype=np.uint8)
        node_of = np.empty(n, dtype=np.uint32) if want_nodes else None
        check(self.lib.ss_synth_db_host(self.h, C.byref(params), C.c_void_p(node_sizes.ctypes.data), node_sizes.size,
                                        C.c_void_p(text.ctypes.data),
                                        C.c_void_p(node_of.ctypes.data) if want_nodes else None))
        return text, node_of


class KmerSet:
    def __init__(self, eng, h):
        self.eng, self.h = eng, h
        lib = eng.lib
        self.n_records = int(lib.ss_kmerset_records(h))
        self.n_distinct = int(lib.ss_kmerset_distinct(h))
        self.k = int(lib.ss_kmerset_k(h))
        self.table_bytes = int(lib.ss_kmerset_table_bytes(h))
        self._flags = None
        self._hid = None

    @property
    def flags(self):
        if self._flags is None:
            f = np.zeros(self.n_records, dtype=np.uint8)
            check(self.eng.lib.ss_kmerset_flags(self.h, C.c_void_p(f.ctypes.data)))
            self._flags = f
        return self._flags

    @property
    def header_ids(self):
        if self._hid is None:
            f = np.zeros(self.n_records, dtype=np.uint64)
            check(self.eng.lib.ss_kmerset_header_ids(self.h, C.c_void_p(f.ctypes.data)))
            self._hid = f
        return self._hid

    @property
    def valid(self):
        """valid_kmers of identify.py:410 as a boolean mask over record ordinals."""
        need = _lib.SS_REC_IN_SET | _lib.SS_REC_IS_LAST
        return (self.flags & need) == need

    def free(self):
        if self.h:
            self.eng.lib.ss_kmerset_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.eng.h:
                self.free()
        except Exception:
            pass


class NodeIndex:
    """Persistent device CSR node -> record ordinals (ss_node_index)."""

    def __init__(self, eng, h, n_nodes):
        self.eng, self.h, self.n_nodes = eng, h, n_nodes

    def free(self):
        if self.h:
            self.eng.lib.ss_node_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.eng.h:
                self.free()
        except Exception:
            pass


class Reads:
    def __init__(self, eng, h):
        self.eng, self.h = eng, h
        self.n_bytes = int(eng.lib.ss_reads_bytes(h))
        self._keepalive = None

    def drop_index(self):
        """Forget the line index: the next counting pass rebuilds it (and pays for it) as a first pass does."""
        check(self.eng.lib.ss_reads_drop_index(self.h))

    def free(self):
        if self.h:
            self.eng.lib.ss_reads_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.eng.h:
                self.free()
        except Exception:
            pass
