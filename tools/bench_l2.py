#!/usr/bin/env python
"""L2-pass regimes of the probe kernel (one pass per identified cluster, Vote_Strain_L2_Lasso_new_sp.py:295-296,
over the reads already resident in HBM): cluster k-mer sets of 2e5 .. 1e7 records (both strands), from
L2-cache-resident tables probed directly (no filter) to HBM tables with a high hit rate, where the
random-sector roofline -- not the filter -- is what bounds the kernel.  Prints one JSON object.
Usage: python tools/bench_l2.py [--reads 10000000]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", type=int, default=-1, help="run only case number N (for ncu captures)")
    ap.add_argument("--verify", action="store_true",
                    help="also run every case with SS_BIN=0 (direct probing) and require the two dense vectors to be identical")
    a = ap.parse_args()
    import numpy as np
    import torch
    from strainscan_b200 import Engine, synth

    eng = Engine(0)
    out = {"reads": a.reads, "cases": []}
    rand = None
    cases = ((200_000, 0.0), (2_000_000, 0.0), (9_900_000, 0.0), (9_900_000, 0.7), (2_000_000, 0.7))
    for ci, (n_rec, off) in enumerate(cases):
        if a.only >= 0 and ci != a.only:
            continue
        # one cluster genome of 5 Mb; the set holds n_rec of its ~1e7 strand-specific 31-mers; a fraction
        # `off` of the reads comes from elsewhere (the other clusters of the sample)
        p = synth.default_params(n_leaves=1, genome_len=5_000_000, seed=3, sources=[(0, 0, 1.0)], p_offtarget=off)
        sizes = np.array([n_rec], dtype=np.uint32)
        db_text, _ = eng.synth_db_host(p, sizes, want_nodes=False)
        kset = eng.kmerset_from_text(db_text, p.k)
        rec = eng.synth_read_record_bytes(p)
        n_bytes = a.reads * rec
        cap = eng.reads_device_capacity(n_bytes)
        text = torch.empty(cap, dtype=torch.uint8, device="cuda")
        eng.synth_reads_device(p, text.data_ptr(), a.reads, 0)
        reads = eng.reads_from_device(text.data_ptr(), n_bytes, cap, keepalive=text)
        counts = torch.zeros(kset.n_records, dtype=torch.int32, device="cuda")
        for _ in range(3):
            st = eng.count_device(kset, reads, counts.data_ptr())
        ms = []
        for _ in range(a.steps):
            st = eng.count_device(kset, reads, counts.data_ptr())
            ms.append(st.ms_probe)
        assert int(counts.to(torch.int64)[torch.from_numpy(kset.valid).cuda()].sum()) == st.n_hits
        same = None
        if a.verify:
            prev = os.environ.get("SS_BIN")
            os.environ["SS_BIN"] = "0"
            direct = torch.zeros_like(counts)
            st0 = eng.count_device(kset, reads, direct.data_ptr())
            if prev is None:
                os.environ.pop("SS_BIN")
            else:
                os.environ["SS_BIN"] = prev
            same = bool(torch.equal(direct, counts)) and st0.n_hits == st.n_hits and st0.n_kmers == st.n_kmers
            assert same and st0.binned_rounds == 0, "binned and direct probing disagree"
            del direct
        t = sum(ms) / len(ms)
        h, p2 = st.n_hits / st.n_kmers, st.n_second_probe / st.n_kmers
        bpk = 33.0 + 32.0 * p2 + 8.0 * h
        if rand is None:
            rand = eng.random_gather_gbps(1 << 30, 1 << 28, iters=3)
        out["cases"].append({
            "records": int(kset.n_records), "table_mb": kset.table_bytes / 1e6, "offtarget": off,
            "filter": bool(st.n_table_probes), "hit_rate": h, "table_probe_rate": st.n_table_probes / st.n_kmers,
            "probe_ms": t, "kmers_per_s": st.n_kmers / (t * 1e-3), "algorithmic_gbps": st.n_kmers * bpk / (t * 1e-3) / 1e9,
            "hbm_sector_probes_per_s": (st.n_table_probes or st.n_kmers) / (t * 1e-3),
            "binned_rounds": st.binned_rounds, "bins": st.bins, "probe_launches": st.probe_launches,
            "identical_to_direct_probing": same})
        del reads, text, counts, kset
        torch.cuda.empty_cache()
    out["random_sector_gather_gbps_1GiB"] = rand
    print(json.dumps(out))


if __name__ == "__main__":
    main()
