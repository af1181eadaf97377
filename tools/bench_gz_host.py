#!/usr/bin/env python
"""Host-only timing of the gzip ingest (no GPU work): ss_ingest_files_host on a gzip'ed synthetic FASTQ file,
sequential vs parallel decoding.  Usage: python tools/bench_gz_host.py file.fq.gz [threads ...]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strainscan_b200 import _lib


def main():
    lib = _lib.load()
    path = sys.argv[1]
    arr = (C.c_char_p * 1)(path.encode())
    n, nch = C.c_size_t(), C.c_uint32()
    for t in [int(x) for x in sys.argv[2:]] or [1, 4, 8, 12]:
        os.environ["SS_PGZ_THREADS"] = str(t)
        for span in ((2 << 20), (4 << 20)) if t > 1 else ((2 << 20),):
            os.environ["SS_PGZ_SPAN"] = str(span)
            t0 = time.perf_counter()
            rc = lib.ss_ingest_files_host(arr, 1, 0, 1, 32 << 20, 4, None, 0, C.byref(n), C.byref(nch))
            dt = time.perf_counter() - t0
            assert rc == 0, lib.ss_last_error()
            print("threads %2d span %d MiB: %.3f s, %.2f GB/s of text (%d chunks)" % (t, span >> 20, dt, n.value / dt / 1e9, nch.value), flush=True)


if __name__ == "__main__":
    main()
