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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", choices=["reference", "b200"], required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    if not os.path.isdir(REF):
        sys.exit("baseline/_ref missing: run `python baseline/setup_ref.py` where the reference is mounted")
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
    if a.engine == "b200":
        install_b200()
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
