#!/usr/bin/env python
"""bench.py -- StrainScan identification hot path (match+count) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference engine on host cores

Workload (BASELINE.json configs[1]): synthetic E. coli-scale cluster search tree (823 clusters ->
1645 nodes, U[1000,30000] node-specific 31-mers per node, both strands as separate records,
~2.6e7 records in Tree_database/kmer.fa form) and 10 M synthetic 150 bp single-end reads PER GPU
(weak scaling: every rank scans its own 10 M-read shard against a full replica of the table; the
dense count vectors are summed with one NCCL all-reduce, as north_star prescribes).

A step = one L1 match+count pass (what library/identify.py:73-103 does per call): zero the
counters, scan/encode/probe/count every read k-mer, gather the dense per-record vector, all-reduce.
  value : k-mers/s with the FASTQ text already resident in HBM.
  e2e   : the same pass through the C ABI from HOST pinned FASTQ text (chunked H2D inside the
          timed region) with the dense vector copied back to the host.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "k-mers/s match+count (L1 pass, k=31)"
UNIT = "k-mers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--leaves", type=int, default=823, help="clusters in the synthetic search tree")
    ap.add_argument("--sample-reads", type=int, default=3_000_000,
                    help="reads in the bounded CPU sample (3 M: the increment over the seeding time is then well above its noise)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=1)
    return ap.parse_args()


def workload_name(a):
    return ("synthetic E. coli-scale DB (%d clusters / %d tree nodes, U[1000,30000] 31-mers per node, both strands) "
            "+ %d synthetic 150bp SE reads per GPU" % (a.leaves, 2 * a.leaves - 1, a.reads))


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed regions
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is None:
            return
        self._stop.clear()
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own engine (oracle/_ref/jellyfish-linux) on the host cores
# ---------------------------------------------------------------------------------------------
def jellyfish_path():
    p = os.path.join(ROOT, "oracle", "_ref", "jellyfish-linux")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def run_jellyfish(jf, fa, fq, threads, workdir):
    """`count` + `dump -c` with the reference's literal argv (identify.py:86-87), all host threads."""
    out = os.path.join(workdir, "o.jf")
    t0 = time.perf_counter()
    subprocess.check_call([jf, "count", "-m", "31", "-s", "100M", "-t", str(threads), "--if", fa, "-o", out, fq])
    t1 = time.perf_counter()
    with open(os.path.join(workdir, "o.fa"), "wb") as f:
        subprocess.check_call([jf, "dump", "-c", out], stdout=f)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def measure_reference(jf, fa, fq, empty, cores, workdir, warmup, steps):
    """Timings of the reference engine that the whole-pass extrapolation rests on.  The fixed per-pass cost (seeding
    the --if set, dumping it) is the MINIMUM of two runs with no reads, taken after one untimed run (a cold first
    run would inflate it and, with it, the marginal rate); the sample's count time is the median over the steps."""
    run_jellyfish(jf, fa, empty, cores, workdir)
    e = [run_jellyfish(jf, fa, empty, cores, workdir) for _ in range(2)]
    tc0, td0 = min(x[0] for x in e), min(x[1] for x in e)
    runs = [run_jellyfish(jf, fa, fq, cores, workdir) for _ in range(warmup + steps)][warmup:]
    tc, td = statistics.median(x[0] for x in runs), statistics.median(x[1] for x in runs)
    return tc0, td0, tc, td, [x[0] + x[1] for x in runs]


def reference_sample(eng, params, db_text, sample_reads, workdir):
    """Write kmer.fa and a bounded FASTQ sample (the first `sample_reads` reads of rank 0's shard)."""
    import torch
    fa = os.path.join(workdir, "kmer.fa")
    db_text.tofile(fa)
    rec = eng.synth_read_record_bytes(params)
    buf = torch.empty(sample_reads * rec, dtype=torch.uint8, device="cuda")
    eng.synth_reads_device(params, buf.data_ptr(), sample_reads, 0)
    fq = os.path.join(workdir, "sample.fq")
    buf.cpu().numpy().tofile(fq)
    empty = os.path.join(workdir, "empty.fq")
    open(empty, "wb").close()
    del buf
    return fa, fq, empty


def cpu_baseline(eng, params, db_text, args, kmers_per_read, full_kmers):
    """Bounded CPU sample: the reference engine with all host threads.  value = whole-pass k-mers/s
    extrapolated as full / (T_fixed + full / marginal_rate), T_fixed = seeding + dump with no reads."""
    cores = os.cpu_count() or 1
    jf = jellyfish_path()
    tmp = tempfile.mkdtemp(prefix="ssb200_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if jf is None:
            return cpu_baseline_port(eng, params, db_text, args, kmers_per_read)
        fa, fq, empty = reference_sample(eng, params, db_text, args.sample_reads, tmp)
        tc0, td0, tc1, td1, _ = measure_reference(jf, fa, fq, empty, cores, tmp, 0, 2)
        sample_kmers = args.sample_reads * kmers_per_read
        marginal = sample_kmers / max(tc1 - tc0, 1e-3)
        fixed = tc0 + td0
        est = full_kmers / (fixed + full_kmers / marginal)
        return {"value": est, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/jellyfish-linux 2.3.0 count -m 31 -s 100M -t %d --if kmer.fa (+ dump -c) on "
                          "the first %d reads of the workload: count %.2fs (empty input: %.2fs seeding), dump %.2fs; "
                          "marginal %.3g k-mers/s; value = whole-pass rate incl. per-pass seeding+dump, "
                          "Python dump parse (identify.py:90-101) excluded" % (
                              cores, args.sample_reads, tc1, tc0, td1, marginal),
                "marginal_kmers_per_s": marginal, "fixed_s": fixed}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_port(eng, params, db_text, args, kmers_per_read):
    import torch
    from oracle import adapters
    n = min(args.sample_reads, 100_000)
    rec = eng.synth_read_record_bytes(params)
    buf = torch.empty(n * rec, dtype=torch.uint8, device="cuda")
    eng.synth_reads_device(params, buf.data_ptr(), n, 0)
    fq = buf.cpu().numpy().tobytes()
    fa = db_text.tobytes()
    t0 = time.perf_counter()
    adapters.count_dense(fa, 31, [fq])
    dt = time.perf_counter() - t0
    return {"value": n * kmers_per_read / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/kmer_count_oracle.c (scalar C port) on the first %d reads incl. seeding: %.2fs" % (n, dt)}


def run_reference_arm(args, rank, world):
    """--impl reference: rank 0 alone times the reference engine; other ranks exit 0."""
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from strainscan_b200 import Engine, synth
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    params = synth.default_params(n_leaves=args.leaves, seed=args.seed)
    sizes = synth.node_sizes(params, seed=args.seed)
    db_text, _ = eng.synth_db_host(params, sizes, want_nodes=False)
    kpr = params.read_len - params.k + 1
    cores = os.cpu_count() or 1
    jf = jellyfish_path()
    tmp = tempfile.mkdtemp(prefix="ssb200_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if jf is None:
            base = cpu_baseline_port(eng, params, db_text, args, kpr)
            times = [args.sample_reads * kpr / base["value"]]
            kind, sample_reads = "port", min(args.sample_reads, 100_000)
        else:
            fa, fq, empty = reference_sample(eng, params, db_text, args.sample_reads, tmp)
            # fixed per-pass cost (seeding the --if set + dumping it) with no reads, then the sample
            tc0, td0, tc_med, _, times = measure_reference(jf, fa, fq, empty, cores, tmp, min(args.warmup, 1), min(args.steps, 5))
            kind, sample_reads = "reference", args.sample_reads
        ms = 1e3 * sum(times) / len(times)
        sample_kmers = sample_reads * kpr
        full_kmers = args.reads * kpr
        if kind == "reference":
            marginal = sample_kmers / max(tc_med - tc0, 1e-3)
            fixed = tc0 + td0
            value = full_kmers / (fixed + full_kmers / marginal)     # whole 10 M-read pass, per-pass seeding+dump included
            how = ("each step = jellyfish count -t %d + dump -c on a bounded sample (%d reads, full %d-record --if set): "
                   "%.2fs; fixed seeding+dump with no reads %.2fs; marginal %.3g k-mers/s; value = k-mers of the "
                   "whole %d-read pass / (fixed + k-mers/marginal); Python dump parse (identify.py:90-101) excluded"
                   % (cores, sample_reads, int(sizes.sum()), ms / 1e3, fixed, marginal, args.reads))
        else:
            value = base["value"]
            how = base["sample"]
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "k": 31, "step": how},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores if kind == "reference" else 1, "kind": kind,
                             "sample": how},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else that lands on fd 1 (NCCL's version banner,
    library chatter) was re-pointed at stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from strainscan_b200 import Engine, synth

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the match+count path has no CPU fallback")
    from strainscan_b200 import dist as ssd
    local_cpus = ssd.bind_to_gpu_cpus(local_rank) if world > 1 else None     # NUMA-local pinned staging per rank
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Engine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- synthetic database (replicated per GPU) and this rank's read shard (untimed) ---------
    params = synth.default_params(n_leaves=args.leaves, seed=args.seed)
    sizes = synth.node_sizes(params, seed=args.seed)
    t0 = time.perf_counter()
    db_text, node_of = eng.synth_db_host(params, sizes, want_nodes=(rank == 0))
    kset = eng.kmerset_from_text(db_text, params.k)
    t_db = time.perf_counter() - t0
    rec = eng.synth_read_record_bytes(params)
    n_bytes = args.reads * rec
    cap = eng.reads_device_capacity(n_bytes)
    text = torch.empty(cap, dtype=torch.uint8, device=dev)
    eng.synth_reads_device(params, text.data_ptr(), args.reads, rank * args.reads)
    reads = eng.reads_from_device(text.data_ptr(), n_bytes, cap, keepalive=text)
    counts = torch.zeros(kset.n_records, dtype=torch.int32, device=dev)
    kpr = params.read_len - params.k + 1

    sampler = ClockSampler(local_rank)

    def step_resident():
        st = eng.count_device(kset, reads, counts.data_ptr())
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        return st

    # ---- value: text resident in HBM ----------------------------------------------------------
    for _ in range(args.warmup):
        st = step_resident()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_ms, gather_ms, launches = [], [], 0
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        st = step_resident()
        probe_ms.append(st.ms_probe)
        gather_ms.append(st.ms_gather)
        launches += st.total_launches
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    n_kmers = torch.tensor([st.n_kmers, st.n_hits, st.n_second_probe, st.n_reads], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_kmers, op=dist.ReduceOp.SUM)
    ms_per_step = float(t[0]) / args.steps
    tot_kmers, tot_hits, tot_second, tot_reads = (int(x) for x in n_kmers.tolist())
    value = tot_kmers / (ms_per_step * 1e-3)

    # size-independent parity properties at full size (cheap, outside the timed region)
    total_counts = int(counts.to(torch.int64)[torch.from_numpy(kset.valid).to(dev)].sum())
    assert total_counts == tot_hits, "sum of valid-record counts %d != hits %d" % (total_counts, tot_hits)
    assert tot_reads == args.reads * world

    # ---- per-node hit vectors of the whole search tree (K4 = match_node for every node, identify.py:115-127;
    # what identify_low_depth.identify_ranks asks for), outside the timed region, with its own full-size check
    node_reduce = None
    if rank == 0:
        order = np.argsort(node_of, kind="stable").astype(np.uint32)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(node_of, minlength=sizes.size))]).astype(np.uint64)
        t0 = time.perf_counter()
        length, covered, total = eng.node_reduce(kset, counts.data_ptr(), ptr, order)
        t_nr = (time.perf_counter() - t0) * 1e3
        assert int(length.sum()) == int(kset.valid.sum()) and int(total.sum()) == tot_hits
        node_reduce = {"nodes": int(sizes.size), "list_entries": int(order.size), "ms_incl_csr_upload": t_nr,
                       "nodes_with_hits": int((covered > 0).sum())}
        del order, node_of

    # ---- roofline of the dominant kernel (K3 probe), this rank ---------------------------------
    p2 = st.n_second_probe / max(st.n_kmers, 1)
    h = st.n_hits / max(st.n_kmers, 1)
    bytes_per_kmer = 33.0 + 32.0 * p2 + 8.0 * h
    probe_avg_ms = sum(probe_ms) / len(probe_ms)
    achieved = st.n_kmers * bytes_per_kmer / (probe_avg_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "probe_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    rand_gbps = eng.random_gather_gbps(kset.table_bytes, 1 << 28, iters=3) if rank == 0 else None
    roofline = {"bound": "hbm", "kernel": "ss_probe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_kmer": bytes_per_kmer, "kmers_per_launch": st.n_kmers, "launch_ms": probe_avg_ms,
                "note": "achieved = algorithmic bytes (SURVEY 8d: 32 B table sector + 1 B text per k-mer, + second sectors and "
                        "counter updates) / launch time; an L2-resident filter answers ~97 % of the probes, so the DRAM traffic "
                        "(`traffic`, ncu) is ~4x smaller than the algorithmic bytes; what binds the kernel is issue slots (64 % busy) "
                        "and the latency of the L2 filter loads -- halving the filter requests did not speed it up (DESIGN.md section 3)",
                "random_sector_gather_gbps": rand_gbps,
                "frac_of_random_gather": (st.n_kmers * 32.0 * (1 + p2) / (probe_avg_ms * 1e-3) / 1e9 / rand_gbps)
                if rand_gbps else None}

    # ---- e2e: HOST pinned FASTQ text -> C ABI -> dense vector back on the host -----------------
    e2e = None
    if not args.no_e2e:
        host_text = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
        host_text.copy_(text[:n_bytes])
        host_counts = torch.zeros(kset.n_records, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def step_e2e():
            if world > 1:
                s = eng.count_host(kset, [(host_text.data_ptr(), n_bytes)], out_ptr=counts.data_ptr())[1]
                dist.all_reduce(counts, op=dist.ReduceOp.SUM)
                host_counts.copy_(counts, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            else:
                s = eng.count_host(kset, [(host_text.data_ptr(), n_bytes)], out_ptr=host_counts.data_ptr())[1]
            return s

        for _ in range(max(1, min(args.warmup, 2))):
            se = step_e2e()
        barrier()
        sampler.start()
        ev0.record()
        t0 = time.perf_counter()
        e2e_launches = 0
        for _ in range(args.steps):
            se = step_e2e()
            e2e_launches += se.total_launches
        ev1.record()
        barrier()
        e_wall = (time.perf_counter() - t0) * 1e3
        sampler.stop()
        te = torch.tensor([ev0.elapsed_time(ev1), e_wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te[0]) / args.steps
        assert se.n_kmers == st.n_kmers and se.n_hits == st.n_hits, "e2e pass disagrees with the resident pass"
        assert int(host_counts.to(torch.int64)[torch.from_numpy(kset.valid)].sum()) == tot_hits
        e2e = {"value": tot_kmers / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_bytes,
               "d2h_bytes_per_step": 4 * kset.n_records, "ms_per_step": e_ms,
               "reads_per_s": tot_reads / (e_ms * 1e-3), "wall_ms_per_step": float(te[1]) / args.steps,
               "api": "ss_count_host (C ABI) from pinned host FASTQ text, 32 MiB chunks over 4 device slots, copy/compute overlapped; bytes are per rank"}
        # what the platform gives a bare copy of the same pinned buffer, all ranks copying at once: e2e is bound by it
        dst = torch.empty(min(n_bytes, 1 << 30), dtype=torch.uint8, device=dev)
        dst.copy_(host_text[:dst.numel()], non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(3):
            dst.copy_(host_text[:dst.numel()], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        tc = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        e2e["bare_h2d_gbps_per_rank"] = 3 * dst.numel() / (float(tc[0]) * 1e-3) / 1e9
        e2e["h2d_gbps_per_rank"] = n_bytes / (e_ms * 1e-3) / 1e9
        del host_text, dst

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(eng, params, db_text, args, kpr, args.reads * kpr)

    if rank == 0:
        info = eng.device_info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "k": params.k, "records": kset.n_records,
                       "distinct_kmers": kset.n_distinct, "table_bytes": kset.table_bytes,
                       "text_bytes_per_gpu": n_bytes, "reads_total": tot_reads,
                       "l2": "inputs exceed L2: %.2f GB text + %.2f GB table per GPU vs 126 MB" % (
                           n_bytes / 1e9, kset.table_bytes / 1e9),
                       "hit_rate": h, "second_sector_rate": p2,
                       "table_probe_rate": st.n_table_probes / max(st.n_kmers, 1), "db_build_s": t_db, "seed": args.seed,
                       "collective": "NCCL all-reduce(sum) of the dense int32 count vector" if world > 1 else "none (1 GPU)",
                       "rank0_cpu_affinity": ("%d GPU-local cores" % len(local_cpus)) if local_cpus else "unchanged"},
            "reads_per_s": tot_reads / (ms_per_step * 1e-3),
            "wall_ms_per_step": float(t[1]) / args.steps,
            "kernel_ms": {"probe": probe_avg_ms, "gather": sum(gather_ms) / len(gather_ms)},
            "node_reduce": node_reduce, "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": launches, "clocks": sampler.summary(),
            "device": info["name"], "n_sm": info["n_sm"],
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
