"""Multi-GPU form of the match+count pass: reads shard across ranks (one process per GPU,
torch.distributed), every rank probes a full replica of the table, and the dense integer count
vectors are summed with ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests of the
host-side logic).  Nothing else is exchanged.

Order matters for L2: remove_1 (count == 1 -> 0, Vote_Strain_L2_Lasso_new_sp.py:312-322) is applied
AFTER the sum, never per shard -- a k-mer seen once on each of two ranks has count 2.
"""
import ctypes as C

import numpy as np

from . import _lib


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bind_to_gpu_cpus(device_index):
    """Pin this process (one per GPU) to the CPU cores NVML reports as local to its GPU, so that the pinned
    staging buffers it allocates afterwards are first-touched on that NUMA node and host->device copies of
    several ranks do not all cross one socket's memory controllers.  Returns the core list, or None when
    NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        # a rank's ingest runs a dozen producer threads (pread / inflate into pinned chunks): pinning it to a handful of
        # cores costs more than remote memory does (seen on a 4-GPU lease: two ranks uploaded at a fifth of the others'
        # rate), so the binding is only taken when it leaves the rank a fair share of the box
        world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        if len(cpus) < max(4, len(allowed) // (2 * world)):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_range(buf, shard, n_shards):
    """Record-aligned byte range of FASTQ text owned by `shard` (host-only C ABI helper)."""
    lib = _lib.load()
    lo, hi = C.c_size_t(), C.c_size_t()
    _lib.check(lib.ss_fastq_shard_range(buf, len(buf), int(shard), int(n_shards), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def allreduce_counts(t, group=None):
    """In-place integer SUM of a dense count tensor over all ranks (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def count_sharded(engine, kset, fq_paths, group=None):
    """Every rank counts its record-aligned shard of the read files; returns the summed dense vector
    as a torch.int32 CUDA tensor (identical on all ranks) and this rank's Stats."""
    import torch
    from .identify_shim import cached_reads
    rank, n = world()
    reads = cached_reads(engine, fq_paths, rank, n)
    dev = torch.device("cuda", engine.device)
    out = torch.zeros(max(kset.n_records, 1), dtype=torch.int32, device=dev)
    engine.follow_torch_stream(dev)          # the all-reduce below and the next pass are ordered after these kernels
    st = engine.count_device(kset, reads, out.data_ptr())
    allreduce_counts(out, group)
    return out[:kset.n_records], st


def jellyfish_count_sharded(engine, fq_path, db_dir, k=31, group=None):
    """Multi-GPU jellyfish_count (identify.py:73-103): same CountVector on every rank."""
    import os
    from .identify_shim import CountVector, cached_kmerset
    kset = cached_kmerset(engine, os.path.join(db_dir, "kmer.fa"), k)
    dense, st = count_sharded(engine, kset, fq_path, group)
    return CountVector(dense.cpu().numpy().view(np.uint32), kset.valid, st)


def count_cluster_sharded(engine, input_fq, fq2, db_dir, ksize, group=None):
    """Multi-GPU L2 count block (Vote_...:348-403): sum first, then remove_1 + kid order."""
    import os
    from .identify_shim import cached_kmerset
    kset = cached_kmerset(engine, os.path.join(db_dir, "all_kmer.fasta"), int(ksize))
    dense, _ = count_sharded(engine, kset, (input_fq, fq2), group)
    return engine.l2_finalize(kset, dense.data_ptr())


def reduce_then_remove_1(local_counts, raw_upper, header_ids=None, group=None):
    """Host-side statement of the L2 multi-rank rule on a torch CPU/CUDA integer tensor of raw
    per-shard counts: all-reduce, then mask + remove_1 (+ kid order).  Returns np.int64."""
    t = allreduce_counts(local_counts, group)
    c = t.cpu().numpy().astype(np.int64)
    c = np.where(np.asarray(raw_upper).astype(bool), c, 0)
    c[c == 1] = 0
    if header_ids is not None:
        from .l2_shim import kid_row_order
        order = kid_row_order(header_ids)                 # the same rule as remove_1 / ss_l2_finalize
        if order is not None:
            c = c[order]
    return c
