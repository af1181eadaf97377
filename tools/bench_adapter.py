#!/usr/bin/env python
"""The reference's own function boundary at benchmark scale: `jellyfish_count(fq_path, db_dir)` of
library/identify.py:73-103 (the stock function from the sandbox baseline/_ref: Jellyfish subprocesses + Python
dump parse) against strainscan_b200.identify_shim.jellyfish_count on the same synthetic E. coli-scale
Tree_database/kmer.fa and FASTQ file, and a check that the two mappings are IDENTICAL (every key, every count).
Prints one JSON object.  Usage: python tools/bench_adapter.py [--reads 4000000] [--gz]"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--leaves", type=int, default=823)
    ap.add_argument("--gz", action="store_true", help="gzip the reads (the reference then runs its zcat pipe)")
    a = ap.parse_args()
    import numpy as np
    import torch
    from strainscan_b200 import Engine, identify_shim, synth

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.access(os.path.join(ref_dir, "library", "jellyfish-linux"), os.X_OK):
        sys.exit("baseline/_ref missing (python baseline/setup_ref.py where the reference is mounted)")
    eng = Engine(0)
    params = synth.default_params(n_leaves=a.leaves, seed=1)
    sizes = synth.node_sizes(params, seed=1)
    db_text, _ = eng.synth_db_host(params, sizes, want_nodes=False)
    tmp = tempfile.mkdtemp(prefix="ssb200_adp_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {"reads": a.reads, "records": int(sizes.sum()), "host_cores": os.cpu_count(), "gz": a.gz}
    try:
        db = os.path.join(tmp, "Tree_database")
        os.makedirs(db)
        db_text.tofile(os.path.join(db, "kmer.fa"))
        rec = eng.synth_read_record_bytes(params)
        buf = torch.empty(a.reads * rec, dtype=torch.uint8, device="cuda")
        eng.synth_reads_device(params, buf.data_ptr(), a.reads, 0)
        fq = os.path.join(tmp, "reads.fq")
        buf.cpu().numpy().tofile(fq)
        del buf
        if a.gz:
            os.system("gzip -1 %s" % fq)
            fq += ".gz"
        eng.close()

        # ---- this repository: cold (database load + table build + ingest + count) and warm (cached set and reads)
        t0 = time.perf_counter()
        cv = identify_shim.jellyfish_count((fq, ""), db)
        t1 = time.perf_counter()
        cv2 = identify_shim.jellyfish_count((fq, ""), db)
        t2 = time.perf_counter()
        assert np.array_equal(cv.counts, cv2.counts)
        out["b200_cold_s"], out["b200_warm_s"] = t1 - t0, t2 - t1
        out["b200_stats"] = {k: v for k, v in cv.stats.as_dict().items() if k in ("n_reads", "n_kmers", "n_hits", "ms_probe", "ms_gather")}

        # ---- the reference's stock function (cwd = sandbox: it writes temp_<uuid>.jf/.fa into cwd)
        sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), ref_dir, os.path.join(ref_dir, "library")]
        cwd = os.getcwd()
        os.chdir(tmp)
        from library import identify as ref_identify
        t0 = time.perf_counter()
        ref = ref_identify.jellyfish_count((fq, ""), db)
        out["reference_s"] = time.perf_counter() - t0
        os.chdir(cwd)

        # ---- parity: same keys (valid k-mers incl. zero counts), same counts
        keys = np.fromiter(ref.keys(), dtype=np.int64, count=len(ref))
        vals = np.fromiter(ref.values(), dtype=np.int64, count=len(ref))
        assert len(ref) == int(cv.valid_mask.sum()), (len(ref), int(cv.valid_mask.sum()))
        assert bool(cv.valid_mask[keys].all())
        assert np.array_equal(cv.counts[keys].astype(np.int64), vals)
        out["identical_mapping"] = True
        out["keys"] = int(len(ref))
        out["sum_counts"] = int(vals.sum())
        out["speedup_cold"] = out["reference_s"] / out["b200_cold_s"]
        out["speedup_warm"] = out["reference_s"] / out["b200_warm_s"]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
