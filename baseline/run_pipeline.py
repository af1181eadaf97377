#!/usr/bin/env python
"""Run the reference's own StrainScan.py identification pipeline from the sandbox baseline/_ref/

    python baseline/run_pipeline.py --engine reference -- -i R.fq -d DB -o OUT [...StrainScan flags]
    python baseline/run_pipeline.py --engine b200      -- -i R.fq -d DB -o OUT [...]

--engine reference : everything stock (bundled jellyfish-linux subprocesses), only the image-compat
                     shims of baseline/shims/ are on sys.path.
--engine b200      : the SAME reference host code, with exactly the edits INTEGRATION.md documents
                     applied at import time: jellyfish_count() of identify / identify_low_mem /
                     identify_low_depth replaced by strainscan_b200.identify_shim.jellyfish_count, and
                     the jellyfish block of vote_strain_L2() (Vote_Strain_L2_Lasso_new_sp.py:348-403)
                     replaced by strainscan_b200.l2_shim.count_cluster.  Tree descent, Pre_Scan,
                     ElasticNet and the report writers are untouched.
--plasmid-db DIR   : plasmid_mode (-p 1 / -p 2) builds a database at run time by shelling out to
                     `python StrainScan_build.py -i <refs> -o <out>/DB_plasmid -n 500` (StrainScan.py:235),
                     whose tool chain (dashing, R, sibeliaz) is absent from this image.  With this option
                     that ONE os.system call is answered by copying the pre-made database DIR to the
                     -o directory of the command, for either engine alike; everything after it (the tree
                     search and the intra-cluster search against DB_plasmid, StrainScan.py:238-266) is the
                     reference's own code.
--engine b200-full : as b200, plus the SURVEY 8f mirrors in place of the reference's per-node and per-strain
                     reductions: match_node of identify / identify_low_mem / identify_low_depth (a5, a7) ->
                     identify_shim.match_node[_low_depth] (array gathers of the dense GPU vector), and
                     cal_cov_all / get_candidate_arr / get_remainc / optimize_dominat_y / get_avg_depth of
                     identify_strains_L2_Enet_Pscan_new_sp (a11, a12) -> l2_shim (ss_strain_reduce on the GPU,
                     order statistics on the sparse columns).  Decisions stay the reference's.
Both engines seed numpy / random identically (identify.py:214 draws unseeded Poisson samples).
Reports (final_report.txt, C*/StrainVote.report, strain_prob.txt) must come out byte-identical.
"""
import argparse
import importlib.util
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def _patch_l2_source(src):
    """INTEGRATION.md section 2 as a text edit of vote_strain_L2()."""
    a = src.index('\tkid_match=pickle.load(open(db_dir+"/all_kid.pkl","rb"))')
    end_marker = "\tpy_o=np.array(py_o)\n"
    b = src.index(end_marker, a) + len(end_marker)
    new = ("\tfrom strainscan_b200.l2_shim import count_cluster as __ss_count_cluster\n"
           "\tpy_o=__ss_count_cluster(input_fq,fq2,db_dir,ksize)\n")
    return src[:a] + new + src[b:]


def install_b200():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from strainscan_b200 import identify_shim
    import library                                                    # StrainScan.py:7 imports `from library import ...`
    from library import identify, identify_low_depth, identify_low_mem
    for mod in (identify, identify_low_mem, identify_low_depth):
        mod.jellyfish_count = identify_shim.jellyfish_count
    path = os.path.join(REF, "library", "Vote_Strain_L2_Lasso_new_sp.py")
    src = _patch_l2_source(open(path).read())
    mod = types.ModuleType("library.Vote_Strain_L2_Lasso_new_sp")
    mod.__file__ = path
    sys.modules["library.Vote_Strain_L2_Lasso_new_sp"] = mod
    library.Vote_Strain_L2_Lasso_new_sp = mod
    exec(compile(src, path, "exec"), mod.__dict__)


def install_b200_reducers():
    """The optional mirrors under the reference's own names and signatures."""
    import numpy as np
    from strainscan_b200 import identify_shim, l2_shim
    from library import identify, identify_low_depth, identify_low_mem
    identify.match_node = identify_shim.match_node
    identify_low_mem.match_node = identify_shim.match_node
    identify_low_depth.match_node = identify_shim.match_node_low_depth

    def make_adjust_profile(mod):
        """adjust_profile (identify.py:167-228) with its gather branch (:167-191) on the array mirror; the Poisson
        overlap correction (:192-228) stays the reference's own code."""
        ref_adjust = mod.adjust_profile

        def adjust_profile(node, results, valid_kmers, length, abundance, cov, match_results, cov_cutoff, db_dir, overlapping_info):
            delete_pos, taken = [], []
            for i in results:                  # identify.py:173-179; the stored values are one-shot map() iterators
                if i.identifier in overlapping_info and node.identifier in overlapping_info[i.identifier]:
                    lst = list(overlapping_info[i.identifier][node.identifier])
                    taken.append((i.identifier, lst))
                    delete_pos.extend(lst)
            res = identify_shim.adjust_profile_gather(match_results, db_dir, node.identifier, delete_pos)
            if res is None:                    # fewer than 1000 k-mers remain: hand the same iterators back
                for ident, lst in taken:
                    overlapping_info[ident][node.identifier] = iter(lst)
                return ref_adjust(node, results, valid_kmers, length, abundance, cov, match_results, cov_cutoff, db_dir, overlapping_info)
            length[node], k_profile = res
            cov[node] = len(k_profile) / length[node]
            abundance[node] = mod.piecewise(cov_cutoff, cov[node], node.data[0], k_profile)
            return 1 if length[node] < 3000 else 2

        return adjust_profile

    identify.adjust_profile = make_adjust_profile(identify)
    identify_low_mem.adjust_profile = make_adjust_profile(identify_low_mem)
    import identify_strains_L2_Enet_Pscan_new_sp as ids          # the module Vote_... imports (library/ on sys.path)

    # the mirrors take the reference's own orientations (cal_cov_all: rows x strains; get_candidate_arr / get_remainc:
    # strains x rows), so they go in directly -- INTEGRATION.md section 3 as written
    def get_candidate_arr(ix, iy):
        cand, check = l2_shim.get_candidate_arr(ix, iy)
        return cand, np.result_type(ix.dtype, np.asarray(iy).dtype).type(check)   # np.sum()'s scalar type is printed

    ids.cal_cov_all, ids.get_candidate_arr, ids.get_remainc = l2_shim.cal_cov_all, get_candidate_arr, l2_shim.get_remainc
    ids.optimize_dominat_y = l2_shim.optimize_dominat_y           # identify_strains...:136-175, same signature
    ids.get_avg_depth = l2_shim.get_avg_depth                     # identify_strains...:109-119


def install_plasmid_db(src_dir):
    """Answer `python StrainScan_build.py ... -o X ...` (StrainScan.py:235) with a copy of src_dir at X."""
    import shlex
    import shutil
    real_system = os.system

    def system(cmd):
        tok = shlex.split(cmd) if isinstance(cmd, str) else []
        if len(tok) > 2 and tok[0] == "python" and tok[1] == "StrainScan_build.py" and "-o" in tok:
            dst = tok[tok.index("-o") + 1]
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(src_dir, dst)
            print("[run_pipeline] StrainScan_build.py stand-in: %s -> %s" % (src_dir, dst))
            return 0
        return real_system(cmd)

    os.system = system


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", choices=["reference", "b200", "b200-full"], required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--plasmid-db", default=None)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    if not os.path.isdir(REF):
        sys.exit("baseline/_ref missing: run `python baseline/setup_ref.py` where the reference is mounted")
    if a.plasmid_db:
        a.plasmid_db = os.path.abspath(a.plasmid_db)
    # absolute paths before we chdir into the sandbox (the reference uses cwd-relative imports/temp files)
    for i, tok in enumerate(rest):
        if i > 0 and rest[i - 1] in ("-i", "-j", "-d", "-o", "-r") and not os.path.isabs(tok):
            rest[i] = os.path.abspath(tok)
    sys.path[:0] = [os.path.join(HERE, "shims"), REF, os.path.join(REF, "library")]
    os.chdir(REF)
    import ss_compat
    ss_compat.apply()
    import numpy as np
    np.random.seed(a.seed)
    random.seed(a.seed)
    if a.engine in ("b200", "b200-full"):
        install_b200()
    if a.engine == "b200-full":
        install_b200_reducers()
    if a.plasmid_db:
        install_plasmid_db(a.plasmid_db)
    spec = importlib.util.spec_from_file_location("StrainScan", os.path.join(REF, "StrainScan.py"))
    ss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ss)
    sys.argv = ["StrainScan.py"] + rest
    try:
        ss.main()
    except SystemExit as e:       # the reference calls exit() on several normal paths
        if e.code not in (None, 0):
            raise


if __name__ == "__main__":
    main()
