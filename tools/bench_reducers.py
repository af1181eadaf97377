#!/usr/bin/env python
"""Consumer-side reducers at full scale, on the GPU box:

  L1  the `-b 1` coverage pass of identify_low_depth.identify_ranks (identify_low_depth.py:113-132): every node of an
      E. coli-scale search tree (1645 nodes, 2.58e7 list entries in Tree_database/kmers/<node>).
        reference : the stock match_node() per node (text parse + Python set / dict work), from baseline/_ref
        shim cold : identify_shim.node_coverage_all, first call (parses the lists once, uploads the CSR, one K4 launch)
        shim warm : the second call identify_ranks makes (identify_low_depth.py:119 and :124) -- one K4 launch
  L2  the per-strain reductions Pre_Scan asks for (identify_strains_L2_Enet_Pscan_new_sp.py:241-334): one cal_cov_all,
      one get_remainc and 15 get_candidate_arr calls over an N2 x S strain matrix.
        reference : the stock dense NumPy functions (from baseline/_ref)
        shim dense: l2_shim under the reference's signatures, dense arguments (what INTEGRATION.md section 3 installs)
        shim sparse: l2_shim on a StrainMatrix (CSC uploaded once into a persistent device handle; K5 per call)

Prints one JSON object.  The reference arm needs baseline/_ref (python baseline/setup_ref.py); without it only the
shim numbers are reported.   Usage: python tools/bench_reducers.py [--rows 4000000] [--strains 64] [--leaves 823]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "baseline", "_ref", "library")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--strains", type=int, default=64)
    ap.add_argument("--leaves", type=int, default=823)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--ref-nodes", type=int, default=200, help="nodes the stock match_node is timed on (scaled to all)")
    a = ap.parse_args()
    import numpy as np
    import scipy.sparse as sp
    import torch
    from strainscan_b200 import Engine, identify_shim, l2_shim, synth

    have_ref = os.path.isdir(REF_LIB)
    if have_ref:
        sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), os.path.dirname(REF_LIB), REF_LIB]
        import ss_compat
        ss_compat.apply()
    eng = Engine(0)
    out = {"have_reference": have_ref}

    # ---------------------------------------------------------------- L1: coverage of every tree node
    p = synth.default_params(n_leaves=a.leaves, seed=1)
    sizes = synth.node_sizes(p, seed=1)
    db_text, node_of = eng.synth_db_host(p, sizes, want_nodes=True)
    tmp = tempfile.mkdtemp(prefix="ssb200_red_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    db_dir = os.path.join(tmp, "Tree_database")
    os.makedirs(os.path.join(db_dir, "kmers"))
    db_text.tofile(os.path.join(db_dir, "kmer.fa"))
    ptr, order = synth.node_csr(node_of, sizes.size)
    for v in range(sizes.size):
        with open(os.path.join(db_dir, "kmers", str(v)), "w") as f:
            f.write(" ".join(map(str, order[int(ptr[v]):int(ptr[v + 1])].tolist())) + " \n")
    rec = eng.synth_read_record_bytes(p)
    buf = torch.empty(a.reads * rec, dtype=torch.uint8, device="cuda")
    eng.synth_reads_device(p, buf.data_ptr(), a.reads, 0)
    fq = os.path.join(tmp, "reads.fq")
    buf.cpu().numpy().tofile(fq)
    del buf
    mr = identify_shim.jellyfish_count((fq, ""), db_dir, engine=eng)
    node_ids = list(range(sizes.size))
    t0 = time.perf_counter()
    cov_cold = identify_shim.node_coverage_all(mr, db_dir, node_ids)
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    cov_warm = identify_shim.node_coverage_all(mr, db_dir, node_ids)
    t_warm = time.perf_counter() - t0
    assert cov_cold == cov_warm
    l1 = {"nodes": int(sizes.size), "list_entries": int(order.size), "shim_cold_s": t_cold, "shim_warm_s": t_warm,
          "nodes_covered": sum(1 for v in cov_warm.values() if v > 0)}
    if have_ref:
        from library import identify_low_depth as ild
        match_results = mr.to_dict()
        valid = set(match_results.keys())
        n_ref = min(a.ref_nodes, sizes.size)
        pick = np.linspace(0, sizes.size - 1, n_ref).astype(int).tolist()
        t0 = time.perf_counter()
        ref_cov = {}
        for v in pick:
            ln, prof = ild.match_node(match_results, db_dir, v, valid)
            ref_cov[v] = -1 if ln == 0 else len(prof) / ln
        t_ref = (time.perf_counter() - t0) * sizes.size / n_ref
        assert all(ref_cov[v] == cov_warm[v] for v in pick), "coverage differs from the stock match_node"
        l1.update({"reference_s_all_nodes": t_ref, "reference_timed_on_nodes": n_ref,
                   "speedup_cold": t_ref / t_cold, "speedup_warm": t_ref / t_warm})
    out["l1_node_coverage"] = l1

    # ---------------------------------------------------------------- L2: Pre_Scan's per-strain reductions
    rng = np.random.default_rng(0)
    n, S = a.rows, a.strains
    X = sp.random(n, S, density=0.08, format="csc", dtype=np.float32, random_state=1)
    X.data[:] = 1
    X = X.astype(np.int8)
    y = (rng.poisson(6, n) * (rng.random(n) < 0.7)).astype(np.int64)
    y[y == 1] = 0
    used = (rng.random(n) < 0.3).astype(np.int64)
    n_cand = 15

    def pre_scan_calls(cal_cov_all, get_remainc, get_candidate_arr, pX, pXt, npXt):
        cov = cal_cov_all(pX, y)
        rem = get_remainc(0, used, pXt, y, {})
        cands = [get_candidate_arr(npXt, y) for _ in range(n_cand)]
        return cov, rem, cands

    m = l2_shim.StrainMatrix(X)
    mask = (used == 0).astype(np.uint8)
    l2_shim.cal_cov_all(m, y, engine=eng)                              # upload + warm-up
    t0 = time.perf_counter()
    cov_s = l2_shim.cal_cov_all(m, y, engine=eng)
    rem_s = l2_shim.get_remainc(0, used, m.T, y, {}, engine=eng)
    cand_s = [l2_shim.get_candidate_arr(m.T, y, row_mask=mask, engine=eng) for _ in range(n_cand)]
    t_sparse = time.perf_counter() - t0
    l2 = {"rows": n, "strains": S, "nnz": int(X.nnz), "calls": "1 cal_cov_all + 1 get_remainc + %d get_candidate_arr" % n_cand,
          "shim_sparse_s": t_sparse}
    if n * S <= 600_000_000:
        pX = X.toarray()
        pXt = np.ascontiguousarray(pX.T).astype(np.int64)
        npXt = 2 * used + pXt
        npXt[npXt > 1] = 0
        t0 = time.perf_counter()
        cov_d, rem_d, cand_d = pre_scan_calls(l2_shim.cal_cov_all, l2_shim.get_remainc, l2_shim.get_candidate_arr, pX, pXt, npXt)
        l2["shim_dense_s"] = time.perf_counter() - t0
        assert cov_d == cov_s and rem_d == rem_s and cand_d == cand_s
        if have_ref:
            import identify_strains_L2_Enet_Pscan_new_sp as ref
            t0 = time.perf_counter()
            cov_r, rem_r, cand_r = pre_scan_calls(ref.cal_cov_all, ref.get_remainc, ref.get_candidate_arr, pX, pXt, npXt)
            t_ref = time.perf_counter() - t0
            assert cov_r == cov_s and all(rem_r[i] == rem_s[i] for i in rem_r) and [(c, int(k)) for c, k in cand_r] == cand_s
            l2.update({"reference_s": t_ref, "speedup_sparse": t_ref / t_sparse, "speedup_dense": t_ref / l2["shim_dense_s"]})
    out["l2_pre_scan_reductions"] = l2
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
