"""Multi-GPU form on real GPUs (skipped with fewer than 2): two ranks over NCCL each count a
record-aligned shard of the same files (one BGZF, one ordinary gzip) through strainscan_b200.dist, and every rank must end up with
the single-GPU vectors (L1 CountVector and L2 py_o with remove_1 applied after the sum)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import adapters
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from strainscan_b200 import Engine, dist as ssd
    rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    eng = Engine(lr)
    d = sys.argv[1]
    cv = ssd.jellyfish_count_sharded(eng, (d + "/r1.fq.gz", d + "/r2.fq.gz"), d + "/Tree_database")
    py_o = ssd.count_cluster_sharded(eng, d + "/r1.fq.gz", d + "/r2.fq.gz", d + "/C1", 31)
    np.save(d + "/l1_counts_%%d.npy" %% rank, cv.counts); np.save(d + "/l1_valid_%%d.npy" %% rank, cv.valid_mask)
    np.save(d + "/l2_%%d.npy" %% rank, py_o)
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_nccl_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gzip
    rng = np.random.default_rng(21)
    G = util.rand_genome(rng, 120_000)
    d = str(tmp_path)
    os.makedirs(d + "/Tree_database"); os.makedirs(d + "/C1")
    fa1 = util.make_db(rng, G, 31, 20_000, both_strands=True, junk=16)
    fa2 = util.make_db(rng, G, 31, 8_000, both_strands=True, header=None)
    open(d + "/Tree_database/kmer.fa", "wb").write(fa1)
    open(d + "/C1/all_kmer.fasta", "wb").write(fa2)
    fq1 = util.make_reads(rng, G, 6000, 150)
    fq2 = util.make_reads(rng, G, 5000, 120, var_len=True)
    open(d + "/r1.fq.gz", "wb").write(util.bgzf_compress(fq1, block=9000))     # blocked gzip: batches dealt to the ranks,
    with gzip.open(d + "/r2.fq.gz", "wb") as f:
        f.write(fq2)
    open(d + "/worker.py", "w").write(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), d + "/worker.py", d], timeout=600,
                          env=dict(os.environ, SS_CHUNK_BYTES=str(1 << 20), SS_BGZF_OUT_CAP=str(1 << 20)))   # inflated on each device
    o1 = adapters.count_dense(fa1, 31, [fq1, fq2])
    o2 = adapters.count_dense(fa2, 31, [fq1, fq2])
    for r in (0, 1):
        assert np.array_equal(np.load(d + "/l1_counts_%d.npy" % r).astype(np.uint64), o1.cnt)
        assert np.array_equal(np.load(d + "/l1_valid_%d.npy" % r), (o1.in_set & o1.is_last).astype(bool))
        assert np.array_equal(np.load(d + "/l2_%d.npy" % r), adapters.l2_py_o(o2))
