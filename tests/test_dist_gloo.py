"""world_size = 2 gloo test of the multi-rank host logic on CPU: record-aligned shard ranges (C-ABI
host helper), integer all-reduce of the dense vectors, and remove_1 applied AFTER the sum.  The
per-shard counter here is the oracle (no GPU in this container); on the GPU box the same dist.py
functions run with the CUDA path (tests/test_gpu_dist.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adapters
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    rng = np.random.default_rng(8)
    G = util.rand_genome(rng, 30_000)
    fa = util.make_db(rng, G, 21, 4000, both_strands=True, header=None, lower_frac=0.01)
    fq = util.make_reads(rng, G, 1500, 100, var_len=True)
    return fa, fq


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from strainscan_b200 import dist as ssd
        fa, fq = _inputs()
        lo, hi = ssd.shard_range(fq, rank, world)
        d = adapters.count_dense(fa, 21, [fq[lo:hi]])
        local = torch.from_numpy(d.cnt.astype(np.int64))
        wrong = local.clone()                                   # remove_1 BEFORE the sum (the bug to avoid)
        wrong[wrong == 1] = 0
        py_o = ssd.reduce_then_remove_1(local, d.raw_upper, d.header_id)
        ssd.allreduce_counts(wrong)
        r, n = ssd.world()
        q.put((rank, r, n, lo, hi, py_o, wrong.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sum_then_remove_1():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fa, fq = _inputs()
    whole = adapters.count_dense(fa, 21, [fq])
    expect = adapters.l2_py_o(whole)
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == len(fq)
    for rank, r, n, lo, hi, py_o, wrong in res:
        assert (r, n) == (rank, world)
        assert np.array_equal(py_o, expect)
    # the order of operations matters on this input: per-shard remove_1 loses k-mers seen once per shard
    w = np.where(whole.raw_upper.astype(bool), res[0][6], 0)
    w[w == 1] = 0
    assert not np.array_equal(w[np.argsort(whole.header_id, kind="stable")], expect)


def test_single_process_is_a_noop():
    from strainscan_b200 import dist as ssd
    t = torch.arange(5)
    assert ssd.allreduce_counts(t.clone()).tolist() == t.tolist()
    assert ssd.world() == (0, 1)
