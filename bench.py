#!/usr/bin/env python
"""bench.py -- StrainScan identification hot path (match+count) on B200.

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference engine on host cores
    python bench.py --config c3 ...                                  # BASELINE.json configs[2]
    python bench.py --config c5 ...                                  # BASELINE.json configs[4]

Default workload (--config c2, BASELINE.json configs[1]): synthetic E. coli-scale cluster search tree
(823 clusters -> 1645 nodes, U[1000,30000] node-specific 31-mers per node, both strands as separate
records, ~2.6e7 records in Tree_database/kmer.fa form) and 10 M synthetic 150 bp single-end reads PER
GPU (weak scaling: every rank scans its own 10 M-read shard against a full replica of the table; the
dense count vectors are summed with one NCCL all-reduce, as north_star prescribes).

A step = one L1 match+count pass (what library/identify.py:73-103 does per call): zero the counters,
build the line index, scan/encode/probe/count every read k-mer, gather the dense per-record vector,
all-reduce.
  value        : k-mers/s with the FASTQ text already resident in HBM, line index built inside the step.
  value_rescan : the same pass with the line index kept from an earlier pass (what each L2 pass costs).
  e2e          : the same pass through the C ABI from HOST pinned FASTQ text (chunked H2D inside the
                 timed region) with the dense vector copied back to the host.

--config c3 (S. aureus-scale DB, 50 M gzipped 2x150 bp pairs, ONE sample sharded over the ranks: strong
scaling): value = resident shard per rank, e2e = ss_count_files from the two .fq.gz files.
--config c5: read-count sweep 1 M .. 500 M reads sharded over the ranks (reads generated on the device).

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
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="c2: reads per GPU")
    ap.add_argument("--leaves", type=int, default=None, help="clusters in the synthetic search tree (c2/c5: 823, c3: 202)")
    ap.add_argument("--pairs", type=int, default=50_000_000, help="c3: read pairs of the sample (two .fq.gz files)")
    ap.add_argument("--member-reads", type=int, default=200_000, help="c3: reads per gzip member of the generated files")
    ap.add_argument("--gz-level", type=int, default=1, help="c3: deflate level of the generated files")
    ap.add_argument("--sweep", default="1,3,10,30,100,500", help="c5: total reads of every point, in millions")
    ap.add_argument("--sample-reads", type=int, default=3_000_000,
                    help="reads of the bounded CPU sample beside the GPU line (cpu_baseline)")
    ap.add_argument("--ref-reads", type=int, default=None,
                    help="reference arm: reads per step (default: c2 the full 10 M-read pass; c3 a bounded 10 M-read sample)")
    ap.add_argument("--ref-max-steps", type=int, default=3, help="reference arm: timed steps actually run (each ~20 s)")
    ap.add_argument("--no-ref-extras", action="store_true", help="reference arm: skip the -t 8 run and the adapter parse")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the N>1 full parity check on rank 0")
    ap.add_argument("--e2e-variants", default="", help="c3: 'NAME=VAL[,NAME=VAL];...' -- the file pass timed again under "
                                                       "each environment setting (same files; A/B of library knobs)")
    ap.add_argument("--tmp", default=None, help="scratch directory for generated files (default /dev/shm)")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    if a.leaves is None:
        a.leaves = 202 if a.config == "c3" else 823
    return a


def workload_name(a):
    if a.config == "c3":
        return ("synthetic S. aureus-scale DB (%d clusters / %d tree nodes, U[1000,30000] 31-mers per node, both strands) "
                "+ %d gzipped 2x150bp PE read pairs (two multi-member .fq.gz files), ONE sample sharded over the GPUs"
                % (a.leaves, 2 * a.leaves - 1, a.pairs))
    if a.config == "c5":
        return ("read-count sweep %s M synthetic 150bp reads sharded over the GPUs, synthetic E. coli-scale DB "
                "(%d clusters / %d tree nodes) replicated, NCCL count all-reduce" % (a.sweep, a.leaves, 2 * a.leaves - 1))
    return ("synthetic E. coli-scale DB (%d clusters / %d tree nodes, U[1000,30000] 31-mers per node, both strands) "
            "+ %d synthetic 150bp SE reads per GPU" % (a.leaves, 2 * a.leaves - 1, a.reads))


def scratch_dir(a, prefix):
    base = a.tmp or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    return tempfile.mkdtemp(prefix=prefix, dir=base)


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
# the reference's own engine (oracle/_ref/jellyfish-linux) on the host cores
# ---------------------------------------------------------------------------------------------
def jellyfish_path():
    p = os.path.join(ROOT, "oracle", "_ref", "jellyfish-linux")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def run_jellyfish(jf, fa, fqs, threads, workdir, gz=False):
    """`count` + `dump -c` with the reference's argv (identify.py:82-87): plain files as arguments, gzip'ed ones
    through `zcat a b | jellyfish count /dev/fd/0`.  Returns (count seconds, dump seconds)."""
    out = os.path.join(workdir, "o.jf")
    t0 = time.perf_counter()
    if gz:
        cmd = "zcat %s | %s count /dev/fd/0 -m 31 -s 100M -t %d --if %s -o %s" % (" ".join(fqs), jf, threads, fa, out)
        subprocess.check_call(cmd, shell=True)
    else:
        subprocess.check_call([jf, "count", "-m", "31", "-s", "100M", "-t", str(threads), "--if", fa, "-o", out] + list(fqs))
    t1 = time.perf_counter()
    with open(os.path.join(workdir, "o.fa"), "wb") as f:
        subprocess.check_call([jf, "dump", "-c", out], stdout=f)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def adapter_parse_seconds(fa, dump, frac=0.1):
    """The Python side of jellyfish_count (identify.py:90-101: kmer.fa -> {KMER: ordinal}, dump -> {ordinal: count}),
    restated and timed on the first `frac` of both files, scaled to the whole (the cost is linear in the records)."""
    n_fa = max(2, int(os.path.getsize(fa) * frac))
    n_dump = max(2, int(os.path.getsize(dump) * frac))
    t0 = time.perf_counter()
    with open(fa, "r") as f:
        lines = f.read(n_fa).split("\n")
    kmer_index_dict = {}
    for i in range(len(lines) // 2 - 1):
        kmer_index_dict[lines[2 * i + 1].rstrip().upper()] = i
    match_results = {}
    with open(dump, "r") as f:
        for line in f.read(n_dump).split("\n")[:-1]:
            k, c = line.split()
            match_results[kmer_index_dict.get(k, -1)] = int(c)
    return (time.perf_counter() - t0) / frac


def host_workload(a, n_reads, first_read=0):
    """Synthetic kmer.fa text and FASTQ text generated on host threads by tools/libss_synth_host.so -- the product
    library is not loaded."""
    from strainscan_b200 import synth              # parameter struct + node sizes only (no .so behind it)
    from tools import synth_host
    params = synth.default_params(n_leaves=a.leaves, seed=a.seed)
    sizes = synth.node_sizes(params, seed=a.seed)
    db_text, _ = synth_host.db(params, sizes, want_nodes=False)
    fq = synth_host.reads(params, n_reads, first_read) if n_reads else None
    return params, sizes, db_text, fq


def cpu_baseline(a, kpr):
    """Bounded CPU sample beside the GPU line: the reference engine with all host threads on the first
    `sample_reads` reads of the workload (directly measured, fixed per-pass seeding + dump included)."""
    cores = os.cpu_count() or 1
    jf = jellyfish_path()
    tmp = scratch_dir(a, "ssb200_cpu_")
    try:
        params, sizes, db_text, fq = host_workload(a, a.sample_reads)
        if jf is None:
            return cpu_baseline_port(db_text, fq, a.sample_reads, kpr)
        fa = os.path.join(tmp, "kmer.fa")
        db_text.tofile(fa)
        fqp = os.path.join(tmp, "sample.fq")
        fq.tofile(fqp)
        del fq, db_text
        tc, td = run_jellyfish(jf, fa, [fqp], cores, tmp)
        kmers = a.sample_reads * kpr
        return {"value": kmers / (tc + td), "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/jellyfish-linux 2.3.0 count -m 31 -s 100M -t %d --if kmer.fa (%d records) + dump -c on "
                          "the first %d reads of the workload, one run: count %.2fs, dump %.2fs; directly measured, the "
                          "per-pass seeding of the --if set is inside (so a whole 10 M-read pass runs at a higher rate: "
                          "see bench.py --impl reference); Python dump parse (identify.py:90-101) excluded"
                          % (cores, int(sizes.sum()), a.sample_reads, tc, td),
                "count_s": tc, "dump_s": td}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_port(db_text, fq, n_reads, kpr):
    from oracle import adapters
    n = min(n_reads, 100_000)
    rec = fq.size // n_reads
    t0 = time.perf_counter()
    adapters.count_dense(db_text.tobytes(), 31, [fq[:n * rec].tobytes()])
    dt = time.perf_counter() - t0
    return {"value": n * kpr / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/kmer_count_oracle.c (scalar C port) on the first %d reads incl. seeding: %.2fs" % (n, dt)}


def write_gz_members(path, text, member_bytes, level, threads):
    """`text` (uint8 array, whole FASTQ records) as a multi-member gzip file (one member per `member_bytes`)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def member(i):
        co = zlib.compressobj(level, zlib.DEFLATED, 31)
        return co.compress(memoryview(text[i:i + member_bytes])) + co.flush()

    with ThreadPoolExecutor(max(1, threads)) as ex, open(path, "wb") as f:
        for blob in ex.map(member, range(0, text.size, member_bytes)):
            f.write(blob)


def run_reference_arm(args, rank, world):
    """--impl reference: rank 0 alone times the reference's own engine on the host cores (Jellyfish 2.3.0 with the
    reference's argv, all host threads); other ranks exit 0.  Inputs come from the host generator: the product
    library is never loaded.  `steps` / `warmup` in the line are the numbers actually run."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    jf = jellyfish_path()
    gz = args.config == "c3"
    full_reads = 2 * args.pairs if gz else args.reads
    n_reads = args.ref_reads or (min(full_reads, 10_000_000) if gz else full_reads)
    tmp = scratch_dir(args, "ssb200_ref_")
    try:
        t_gen = time.perf_counter()
        params, sizes, db_text, fq = host_workload(args, n_reads)
        kpr = params.read_len - params.k + 1
        kmers = n_reads * kpr
        if jf is None:
            base = cpu_baseline_port(db_text, fq, n_reads, kpr)
            line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                    "steps": 1, "warmup": 0, "ms_per_step": 1e3 * min(n_reads, 100_000) * kpr / base["value"],
                    "higher_is_better": True, "scaling": "strong" if gz else "weak", "vs_baseline": None, "dtype": "u64",
                    "data": "synthetic", "config": {"workload": workload_name(args), "k": 31, "step": base["sample"]},
                    "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}, "gpu_launches": 0}
            emit(line)
            return
        fa = os.path.join(tmp, "kmer.fa")
        db_text.tofile(fa)
        n_records = int(sizes.sum())
        del db_text
        if gz:
            rec = fq.size // n_reads
            half = (n_reads // 2) * rec
            fqs = [os.path.join(tmp, "r1.fq.gz"), os.path.join(tmp, "r2.fq.gz")]
            write_gz_members(fqs[0], fq[:half], args.member_reads * rec, args.gz_level, cores)
            write_gz_members(fqs[1], fq[half:], args.member_reads * rec, args.gz_level, cores)
        else:
            fqs = [os.path.join(tmp, "reads.fq")]
            fq.tofile(fqs[0])
        text_bytes = int(fq.size)
        del fq
        t_gen = time.perf_counter() - t_gen
        warm = min(max(args.warmup, 0), 1)
        steps = max(1, min(args.steps, args.ref_max_steps))
        runs = [run_jellyfish(jf, fa, fqs, cores, tmp, gz) for _ in range(warm + steps)][warm:]
        times = [c + d for c, d in runs]
        ms = 1e3 * sum(times) / len(times)
        value = kmers / (ms * 1e-3)
        extras = {"count_s": [round(c, 3) for c, _ in runs], "dump_s": [round(d, 3) for _, d in runs],
                  "input_generation_s": round(t_gen, 2), "text_bytes": text_bytes}
        if not args.no_ref_extras:
            # the reference's literal thread count (identify.py:82,86: -t 8), one run, and the Python side of
            # jellyfish_count() (identify.py:90-101) -- reported beside the headline, not inside it
            c8, d8 = run_jellyfish(jf, fa, fqs, 8, tmp, gz)
            extras["t8_literal_argv"] = {"count_s": c8, "dump_s": d8, "kmers_per_s": kmers / (c8 + d8)}
            extras["adapter_s"] = adapter_parse_seconds(fa, os.path.join(tmp, "o.fa"))
            extras["adapter_note"] = ("identify.py:90-101 restated (kmer.fa -> dict, dump -> dict), timed on the first 10 % of "
                                      "both files and scaled by 10; a whole jellyfish_count() call = step + adapter_s")
        if gz:
            how = ("each step = `zcat r1.fq.gz r2.fq.gz | jellyfish count /dev/fd/0 -m 31 -s 100M -t %d --if kmer.fa` + dump -c "
                   "(identify.py:82-84) on a BOUNDED sample: the first %d of the %d reads of the sample, full %d-record "
                   "--if set; measured directly" % (cores, n_reads, full_reads, n_records))
        else:
            how = ("each step = jellyfish count -m 31 -s 100M -t %d --if kmer.fa + dump -c (identify.py:86-87) on %d reads "
                   "(%s), full %d-record --if set; measured directly, no extrapolation; at N > 1 the GPU arm scans this many "
                   "reads per GPU" % (cores, n_reads, "the whole pass of one GPU" if n_reads == args.reads else "bounded sample",
                                      n_records))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if gz else "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "reads_per_s": n_reads / (ms * 1e-3),
            "config": {"workload": workload_name(args), "k": 31, "step": how, "reads_per_step": n_reads},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": how},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "reference_detail": extras,
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


class Ctx:
    """What every GPU configuration needs: process group, engine, database replica, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from strainscan_b200 import Engine, synth
        from strainscan_b200 import dist as ssd
        self.args, self.torch, self.dist = args, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            sys.exit("bench.py: no CUDA device; the match+count path has no CPU fallback")
        self.local_cpus = ssd.bind_to_gpu_cpus(self.local_rank) if self.world > 1 else None   # NUMA-local pinned staging
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = Engine(self.local_rank)
        # One non-default stream carries everything -- this library's kernels, the NCCL collectives torch issues and the
        # timing events -- so each is ordered after the other.  (Not the default stream: as cudaStreamLegacy it cost the
        # file -> counts pass of config 3 about a fifth of its time.)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.eng.follow_torch_stream()
        self.params = synth.default_params(n_leaves=args.leaves, seed=args.seed)
        self.sizes = synth.node_sizes(self.params, seed=args.seed)
        t0 = time.perf_counter()
        self.db_text, self.node_of = self.eng.synth_db_host(self.params, self.sizes, want_nodes=(self.rank == 0))
        self.kset = self.eng.kmerset_from_text(self.db_text, self.params.k)
        self.t_db = time.perf_counter() - t0
        self.rec = self.eng.synth_read_record_bytes(self.params)
        self.kpr = self.params.read_len - self.params.k + 1
        self.counts = torch.zeros(self.kset.n_records, dtype=torch.int32, device=self.dev)
        self.valid_dev = torch.from_numpy(self.kset.valid).to(self.dev)
        self.sampler = ClockSampler(self.local_rank)
        # cores every rank may run on (the ingest of config 3 is host work: pread / inflate threads per rank)
        n_cpu = torch.tensor([len(os.sched_getaffinity(0))], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            all_cpu = [torch.zeros_like(n_cpu) for _ in range(self.world)]
            dist.all_gather(all_cpu, n_cpu)
            self.cpus_per_rank = [int(x) for x in all_cpu]
        else:
            self.cpus_per_rank = [int(n_cpu)]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allreduce(self, t, op=None):
        if self.world > 1:
            self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM)
        return t

    def device_text(self, n_reads, first_read):
        """n_reads synthetic reads generated on the device + the read-cache handle over them."""
        n_bytes = n_reads * self.rec
        cap = self.eng.reads_device_capacity(n_bytes)
        text = self.torch.empty(cap, dtype=self.torch.uint8, device=self.dev)
        self.eng.synth_reads_device(self.params, text.data_ptr(), n_reads, first_read)
        return text, self.eng.reads_from_device(text.data_ptr(), n_bytes, cap, keepalive=text)

    def timed(self, step, warmup, steps):
        """W untimed + K timed steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over
        ranks.  Returns (ms per step, wall ms per step, list of step results)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        self.sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = []
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            out.append(step())
        ev1.record()
        self.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        self.sampler.stop()
        t = torch.tensor([ev0.elapsed_time(ev1), wall], dtype=torch.float64, device=self.dev)
        self.allreduce(t, self.dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps, out

    def sum_stats(self, st):
        t = self.torch.tensor([st.n_kmers, st.n_hits, st.n_second_probe, st.n_reads, st.n_table_probes],
                              dtype=self.torch.int64, device=self.dev)
        self.allreduce(t)
        return [int(x) for x in t.tolist()]

    def valid_sum(self, counts):
        return int(counts.to(self.torch.int64)[self.valid_dev].sum())

    def roofline(self, st, probe_ms):
        p2 = st.n_second_probe / max(st.n_kmers, 1)
        h = st.n_hits / max(st.n_kmers, 1)
        bytes_per_kmer = 33.0 + 32.0 * p2 + 8.0 * h
        achieved = st.n_kmers * bytes_per_kmer / (probe_ms * 1e-3) / 1e9
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
        rand_gbps = self.eng.random_gather_gbps(self.kset.table_bytes, 1 << 28, iters=3) if self.rank == 0 else None
        return {"bound": "hbm", "kernel": "ss_probe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_kmer": bytes_per_kmer, "kmers_per_launch": st.n_kmers, "launch_ms": probe_ms,
                "note": "achieved = algorithmic bytes (SURVEY 8d: 32 B table sector + 1 B text per k-mer, + second sectors and "
                        "counter updates) / launch time; an L2-resident filter answers ~97 % of the probes, so the DRAM traffic "
                        "(`traffic`, ncu, per launch) is ~4x smaller than the algorithmic bytes and `frac` is NOT a statement about "
                        "DRAM saturation: what binds the kernel is issue slots and the latency of the L2 filter loads (DESIGN.md "
                        "section 3); `frac_of_random_gather` compares the k-mer rate with the measured random-sector ceiling",
                "dram_traffic_gbps": (traffic / (probe_ms * 1e-3) / 1e9) if traffic else None,
                "dram_frac": (traffic / (probe_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "random_sector_gather_gbps": rand_gbps,
                "frac_of_random_gather": (st.n_kmers * 32.0 * (1 + p2) / (probe_ms * 1e-3) / 1e9 / rand_gbps)
                if rand_gbps else None}, h, p2

    def base_line(self, value, ms_per_step, scaling):
        a = self.args
        info = self.eng.device_info()
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": self.world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "k": self.params.k, "records": self.kset.n_records,
                       "distinct_kmers": self.kset.n_distinct, "table_bytes": self.kset.table_bytes,
                       "db_build_s": self.t_db, "seed": a.seed,
                       "collective": "NCCL all-reduce(sum) of the dense int32 count vector" if self.world > 1 else "none (1 GPU)",
                       "rank0_cpu_affinity": ("%d GPU-local cores" % len(self.local_cpus)) if self.local_cpus else "unchanged"},
            "device": info["name"], "n_sm": info["n_sm"],
        }

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_c2(args):
    import numpy as np
    import torch
    cx = Ctx(args)
    eng, kset, counts, world, rank, dev = cx.eng, cx.kset, cx.counts, cx.world, cx.rank, cx.dev
    n_bytes = args.reads * cx.rec
    text, reads = cx.device_text(args.reads, rank * args.reads)

    def step_first():
        reads.drop_index()                       # a one-shot L1 pass pays for the line index (K1) too
        st = eng.count_device(kset, reads, counts.data_ptr())
        cx.allreduce(counts)
        return st

    def step_rescan():
        st = eng.count_device(kset, reads, counts.data_ptr())
        cx.allreduce(counts)
        return st

    # ---- value: text resident in HBM, index + probe + gather + all-reduce ----------------------
    ms_per_step, wall_ms, sts = cx.timed(step_first, args.warmup, args.steps)
    st = sts[-1]
    probe_ms = sum(s.ms_probe for s in sts) / len(sts)
    index_ms = sum(s.ms_index for s in sts) / len(sts)
    gather_ms = sum(s.ms_gather for s in sts) / len(sts)
    launches = sum(s.total_launches for s in sts)
    tot_kmers, tot_hits, tot_second, tot_reads, tot_table = cx.sum_stats(st)
    value = tot_kmers / (ms_per_step * 1e-3)
    ms_rescan, _, sts2 = cx.timed(step_rescan, 1, args.steps)
    assert sts2[-1].n_hits == st.n_hits and sts2[-1].ms_index == 0.0

    # size-independent parity properties at full size (cheap, outside the timed region)
    total_counts = cx.valid_sum(counts)
    assert total_counts == tot_hits, "sum of valid-record counts %d != hits %d" % (total_counts, tot_hits)
    assert tot_reads == args.reads * world

    # ---- N > 1: the all-reduced vector against rank 0 counting EVERY shard by itself ------------
    check = None
    if world > 1 and not args.no_check:
        if rank == 0:
            t0 = time.perf_counter()
            acc = torch.zeros_like(counts)
            one = torch.zeros_like(counts)
            for r in range(world):
                if r == 0:
                    eng.count_device(kset, reads, one.data_ptr())
                else:
                    t_r, reads_r = cx.device_text(args.reads, r * args.reads)
                    eng.count_device(kset, reads_r, one.data_ptr())
                    reads_r.free()
                    del t_r
                acc += one
            same = bool(torch.equal(acc, counts))
            check = {"all_reduced_equals_single_rank_count_of_all_shards": same, "shards": world,
                     "records_compared": int(counts.numel()), "seconds": time.perf_counter() - t0}
            assert same, "all-reduced count vector differs from rank 0's own count of all %d shards" % world
            del acc, one
        cx.barrier()

    # ---- per-node hit vectors of the whole search tree (K4 = match_node for every node) ---------
    node_reduce = None
    if rank == 0:
        node_of = cx.node_of
        order = np.argsort(node_of, kind="stable").astype(np.uint32)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(node_of, minlength=cx.sizes.size))]).astype(np.uint64)
        t0 = time.perf_counter()
        length, covered, total = eng.node_reduce(kset, counts.data_ptr(), ptr, order)
        t_nr = (time.perf_counter() - t0) * 1e3
        assert int(length.sum()) == int(kset.valid.sum()) and int(total.sum()) == tot_hits
        node_reduce = {"nodes": int(cx.sizes.size), "list_entries": int(order.size), "ms_incl_csr_upload": t_nr,
                       "nodes_with_hits": int((covered > 0).sum())}
        del order

    roofline, h, p2 = cx.roofline(st, probe_ms)

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
                cx.allreduce(counts)
                host_counts.copy_(counts, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            else:
                s = eng.count_host(kset, [(host_text.data_ptr(), n_bytes)], out_ptr=host_counts.data_ptr())[1]
            return s

        e_ms, e_wall, ses = cx.timed(step_e2e, max(1, min(args.warmup, 2)), args.steps)
        se = ses[-1]
        assert se.n_kmers == st.n_kmers and se.n_hits == st.n_hits, "e2e pass disagrees with the resident pass"
        assert int(host_counts.to(torch.int64)[torch.from_numpy(kset.valid)].sum()) == tot_hits
        e2e = {"value": tot_kmers / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_bytes,
               "d2h_bytes_per_step": 4 * kset.n_records, "ms_per_step": e_ms,
               "reads_per_s": tot_reads / (e_ms * 1e-3), "wall_ms_per_step": e_wall,
               "gpu_launches": sum(s.total_launches for s in ses),
               "api": "ss_count_host (C ABI) from pinned host FASTQ text, 32 MiB chunks over 4 device slots, copy/compute overlapped; bytes are per rank"}
        # what the platform gives a bare copy of the same pinned buffer, all ranks copying at once: e2e is bound by it
        dst = torch.empty(min(n_bytes, 1 << 30), dtype=torch.uint8, device=dev)
        dst.copy_(host_text[:dst.numel()], non_blocking=True)
        torch.cuda.synchronize()
        cx.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            dst.copy_(host_text[:dst.numel()], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        tc = cx.allreduce(torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev), cx.dist.ReduceOp.MAX)
        e2e["bare_h2d_gbps_per_rank"] = 3 * dst.numel() / (float(tc[0]) * 1e-3) / 1e9
        e2e["h2d_gbps_per_rank"] = n_bytes / (e_ms * 1e-3) / 1e9
        del host_text, dst

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args, cx.kpr)

    if rank == 0:
        line = cx.base_line(value, ms_per_step, "weak")
        line["config"].update({
            "text_bytes_per_gpu": n_bytes, "reads_total": tot_reads,
            "l2": "inputs exceed L2: %.2f GB text + %.2f GB table per GPU vs 126 MB" % (n_bytes / 1e9, kset.table_bytes / 1e9),
            "hit_rate": h, "second_sector_rate": p2, "table_probe_rate": tot_table / max(tot_kmers, 1),
            "step": "K1 line index + K3 probe + K3b gather (+ all-reduce): a one-shot L1 pass over resident text"})
        line.update({
            "reads_per_s": tot_reads / (ms_per_step * 1e-3), "wall_ms_per_step": wall_ms,
            "value_rescan": tot_kmers / (ms_rescan * 1e-3), "ms_per_step_rescan": ms_rescan,
            "kernel_ms": {"index": index_ms, "probe": probe_ms, "gather": gather_ms},
            "multi_gpu_check": check, "node_reduce": node_reduce, "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": launches, "clocks": cx.sampler.summary()})
        emit(line)
    cx.finish()


# ---------------------------------------------------------------------------------------------
# config 3: one gzipped paired-end sample sharded over the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------
def make_c3_files(cx, tmp):
    """r1.fq.gz = reads [0, pairs), r2.fq.gz = reads [pairs, 2*pairs) of the generator, as multi-member gzip files
    (one member per --member-reads reads, deflate level --gz-level).  Every rank generates and compresses the members
    m with m % world == rank; rank 0 concatenates.  Untimed."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    a, torch = cx.args, cx.torch
    paths = [os.path.join(tmp, "r1.fq.gz"), os.path.join(tmp, "r2.fq.gz")]
    threads = len(cx.local_cpus) if cx.local_cpus else (os.cpu_count() or 1)
    members = []                                   # (file, member index, first read, n reads)
    for f in range(2):
        for m, lo in enumerate(range(0, a.pairs, a.member_reads)):
            members.append((f, m, f * a.pairs + lo, min(a.member_reads, a.pairs - lo)))
    mine = [x for i, x in enumerate(members) if i % cx.world == cx.rank]
    buf = torch.empty(a.member_reads * cx.rec, dtype=torch.uint8, device=cx.dev)

    def compress(job):
        f, m, host = job
        co = zlib.compressobj(a.gz_level, zlib.DEFLATED, 31)
        blob = co.compress(memoryview(host)) + co.flush()
        with open(os.path.join(tmp, "part_%d_%06d" % (f, m)), "wb") as fh:
            fh.write(blob)
        return len(blob)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        futs = []
        for f, m, first, n in mine:
            cx.eng.synth_reads_device(cx.params, buf.data_ptr(), n, first)
            host = buf[:n * cx.rec].cpu().numpy()
            futs.append(ex.submit(compress, (f, m, host)))
            while sum(1 for x in futs if not x.done()) >= 2 * threads:      # bound the host copies in flight
                time.sleep(0.01)
        for x in futs:
            x.result()
    del buf
    cx.barrier()
    if cx.rank == 0:
        for f in range(2):
            with open(paths[f], "wb") as out:
                for ff, m, _, _ in members:
                    if ff == f:
                        p = os.path.join(tmp, "part_%d_%06d" % (f, m))
                        with open(p, "rb") as src:
                            shutil.copyfileobj(src, out, 64 << 20)
                        os.remove(p)
    cx.barrier()
    return paths, time.perf_counter() - t0, len(members)


def run_c3(args):
    import torch
    cx = Ctx(args)
    eng, kset, counts, world, rank = cx.eng, cx.kset, cx.counts, cx.world, cx.rank
    n_reads = 2 * args.pairs
    # every rank must see the same files: rank 0 names the directory
    box = [scratch_dir(args, "ssb200_c3_") if rank == 0 else None]
    if world > 1:
        cx.dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    try:
        paths, t_files, n_members = make_c3_files(cx, tmp)
        gz_bytes = sum(os.path.getsize(p) for p in paths)
        text_bytes = n_reads * cx.rec

        # ---- value: this rank's shard resident in HBM (loaded once, untimed) ----------------------
        t0 = time.perf_counter()
        reads = eng.reads_from_files(paths, rank, world)
        t_load = time.perf_counter() - t0

        def step_resident():
            reads.drop_index()
            st = eng.count_device(kset, reads, counts.data_ptr())
            cx.allreduce(counts)
            return st

        ms_per_step, wall_ms, sts = cx.timed(step_resident, args.warmup, args.steps)
        st = sts[-1]
        probe_ms = sum(s.ms_probe for s in sts) / len(sts)
        tot_kmers, tot_hits, _, tot_reads, tot_table = cx.sum_stats(st)
        assert tot_reads == n_reads, "ranks scanned %d reads, the sample has %d" % (tot_reads, n_reads)
        assert cx.valid_sum(counts) == tot_hits
        value = tot_kmers / (ms_per_step * 1e-3)
        resident = counts.clone()
        shard_bytes = reads.n_bytes
        reads.free()
        roofline, h, p2 = cx.roofline(st, probe_ms)

        # ---- e2e: the two .fq.gz files -> ss_count_files(shard, n_shards) -> all-reduce -> host -----
        host_inflate = os.environ.get("SS_DGZ", "1") == "0"
        host_counts = torch.zeros(kset.n_records, dtype=torch.int32, pin_memory=True)

        def step_files():
            s = eng.count_files(kset, paths, rank, world, out_ptr=counts.data_ptr())[1]
            cx.allreduce(counts)
            host_counts.copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return s

        e_ms, e_wall, ses = cx.timed(step_files, min(args.warmup, 1), args.steps)
        se = ses[-1]
        assert torch.equal(counts, resident), "file pass and resident pass disagree"
        e_reads = cx.sum_stats(se)[3]
        assert e_reads == n_reads
        per_rank_text = torch.tensor([float(se.text_bytes)], dtype=torch.float64, device=cx.dev)
        mx = cx.allreduce(per_rank_text.clone(), cx.dist.ReduceOp.MAX)
        e2e = {"value": tot_kmers / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms, "wall_ms_per_step": e_wall,
               "reads_per_s": n_reads / (e_ms * 1e-3),
               "h2d_bytes_per_step": int(se.text_bytes) if host_inflate else gz_bytes // world,
               "d2h_bytes_per_step": 4 * kset.n_records,
               "text_gbps_per_rank": float(mx[0]) / (e_ms * 1e-3) / 1e9,
               "text_gbps_all_ranks": text_bytes / (e_ms * 1e-3) / 1e9,
               "gz_gbps_all_ranks": gz_bytes / (e_ms * 1e-3) / 1e9,
               "gpu_launches": sum(s.total_launches for s in ses),
               "api": ("ss_count_files(paths=[r1.fq.gz, r2.fq.gz], shard=rank, n_shards=world) (C ABI): host threads inflate "
                       "this rank's gzip members into pinned chunks, chunked H2D, K1+K3 per chunk, K3b, all-reduce, D2H; "
                       "h2d bytes are this rank's text") if host_inflate else
                      ("ss_count_files(paths=[r1.fq.gz, r2.fq.gz], shard=rank, n_shards=world) (C ABI): this rank's gzip members "
                       "are uploaded COMPRESSED (pread threads, pinned chunks) and inflated on the device (K7-K10), K1+K3 per "
                       "batch where it was inflated, K3b, all-reduce, D2H; h2d bytes are the compressed bytes per rank "
                       "(file bytes / ranks; the text they inflate to is text_gbps_per_rank x time)")}

        # ---- optional: the same file pass under other settings of the library's knobs ---------------
        variants = []
        for spec in [v for v in args.e2e_variants.split(";") if v]:
            sets = dict(kv.split("=", 1) for kv in spec.split(","))
            old_env = {k: os.environ.get(k) for k in sets}
            os.environ.update(sets)
            try:
                v_ms, v_wall, _ = cx.timed(step_files, 1, args.steps)
                assert torch.equal(counts, resident), "file pass under %s disagrees with the resident pass" % spec
            finally:
                for k, v in old_env.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            variants.append({"env": sets, "ms_per_step": v_ms, "wall_ms_per_step": v_wall, "value": tot_kmers / (v_ms * 1e-3)})
        if variants:
            e2e["variants"] = variants

        # ---- N > 1: rank 0 counts the whole sample alone and compares -----------------------------
        check = None
        if world > 1 and not args.no_check:
            if rank == 0:
                t0 = time.perf_counter()
                one = torch.zeros_like(counts)
                eng.count_files(kset, paths, 0, 1, out_ptr=one.data_ptr())
                same = bool(torch.equal(one, counts))
                check = {"all_reduced_equals_single_rank_count_of_the_whole_sample": same,
                         "records_compared": int(counts.numel()), "seconds": time.perf_counter() - t0}
                assert same, "all-reduced count vector differs from rank 0's own count of the whole sample"
            cx.barrier()

        if rank == 0:
            line = cx.base_line(value, ms_per_step, "strong")
            line["config"].update({
                "reads_total": n_reads, "pairs": args.pairs, "text_bytes": text_bytes, "gz_bytes": gz_bytes,
                "gz_members": n_members, "gz_level": args.gz_level, "files_written_s": t_files,
                "resident_load_s": t_load, "shard_text_bytes_rank0": shard_bytes,
                "l2": "inputs exceed L2: %.2f GB text per rank + %.2f GB table vs 126 MB" % (shard_bytes / 1e9, kset.table_bytes / 1e9),
                "hit_rate": h, "second_sector_rate": p2, "table_probe_rate": tot_table / max(tot_kmers, 1),
                "host_cores": os.cpu_count(), "cores_allowed_per_rank": cx.cpus_per_rank,
                "step": "resident: K1 + K3 + K3b over this rank's shard + all-reduce; e2e: from the .fq.gz files"})
            line.update({"reads_per_s": n_reads / (ms_per_step * 1e-3), "wall_ms_per_step": wall_ms,
                         "kernel_ms": {"index": sum(s.ms_index for s in sts) / len(sts), "probe": probe_ms,
                                       "gather": sum(s.ms_gather for s in sts) / len(sts)},
                         "multi_gpu_check": check, "roofline": roofline, "cpu_baseline": None, "e2e": e2e,
                         "gpu_launches": sum(s.total_launches for s in sts), "clocks": cx.sampler.summary()})
            emit(line)
    finally:
        cx.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
    cx.finish()


# ---------------------------------------------------------------------------------------------
# config 5: read-count sweep, reads generated on the device, sharded over the ranks
# ---------------------------------------------------------------------------------------------
def run_c5(args):
    import torch
    cx = Ctx(args)
    eng, kset, counts, world, rank = cx.eng, cx.kset, cx.counts, cx.world, cx.rank
    points = []
    free_b, _ = torch.cuda.mem_get_info()
    for m in [float(x) for x in args.sweep.split(",") if x]:
        total = int(m * 1e6)
        per = (total + world - 1) // world
        mine = max(0, min(per, total - rank * per))
        need = per * cx.rec + (1 << 30)
        if need > free_b:
            points.append({"reads": total, "skipped": "%.1f GB of text per GPU does not fit" % (per * cx.rec / 1e9)})
            continue
        text, reads = cx.device_text(mine, rank * per)

        def step():
            reads.drop_index()
            st = eng.count_device(kset, reads, counts.data_ptr())
            cx.allreduce(counts)
            return st

        steps = args.steps if total <= 100_000_000 else max(2, args.steps // 2)
        ms, wall, sts = cx.timed(step, min(args.warmup, 3), steps)
        st = sts[-1]
        tk, th, _, tr, _ = cx.sum_stats(st)
        vs = cx.valid_sum(counts)
        if rank == 0:
            sys.stderr.write("[c5] %d reads: %.3f ms/step, reads seen %d, hits %d, sum of counts %d\n" % (total, ms, tr, th, vs))
        assert tr == total, "ranks scanned %d reads, the point has %d" % (tr, total)
        assert vs == th, "sum of valid-record counts %d != hits %d at %d reads" % (vs, th, total)
        points.append({"reads": total, "reads_per_gpu": per, "kmers": tk, "ms_per_step": ms, "wall_ms_per_step": wall,
                       "kmers_per_s": tk / (ms * 1e-3), "reads_per_s": total / (ms * 1e-3),
                       "rank0_ms": {"index": sum(s.ms_index for s in sts) / len(sts),
                                    "probe": sum(s.ms_probe for s in sts) / len(sts),
                                    "gather": sum(s.ms_gather for s in sts) / len(sts)},
                       "launches_per_step": st.total_launches, "steps": steps})
        reads.free()
        del text
        torch.cuda.empty_cache()
    if rank == 0:
        done = [p for p in points if "kmers_per_s" in p]
        top = done[-1]
        line = cx.base_line(top["kmers_per_s"], top["ms_per_step"], "strong")
        line["config"].update({"reads_total": top["reads"], "step": "K1 + K3 + K3b + all-reduce over resident text, per sweep point; "
                               "`value` is the largest point that fits"})
        line.update({"sweep": points, "reads_per_s": top["reads_per_s"], "roofline": None, "cpu_baseline": None, "e2e": None,
                     "gpu_launches": sum(p["launches_per_step"] * p["steps"] for p in done), "clocks": cx.sampler.summary()})
        emit(line)
    cx.finish()


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    {"c2": run_c2, "c3": run_c3, "c5": run_c5}[args.config](args)


if __name__ == "__main__":
    main()
