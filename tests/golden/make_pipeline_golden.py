#!/usr/bin/env python
"""Run the REFERENCE pipeline (baseline/_ref StrainScan.py + its bundled jellyfish-linux, stock code
path, only the image-compat shims of baseline/shims on sys.path) on the synthetic database and read
sets of tests/synth_db.py, and store its reports under tests/golden/pipeline/<case>/.

    python baseline/setup_ref.py && python tests/golden/make_pipeline_golden.py [case ...]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import synth_db  # noqa: E402

REPORTS = ("final_report.txt", "strain_prob.txt")


def collect_reports(out_dir):
    res = {}
    for dp, _, files in os.walk(out_dir):
        for f in files:
            if f in REPORTS or f == "StrainVote.report":
                p = os.path.join(dp, f)
                res[os.path.relpath(p, out_dir)] = open(p, "rb").read()
    return res


def run_case(engine, db_dir, name, work):
    """db_dir: one DB directory, or the {low_mem: dir} dict of synth_db.write_dbs()."""
    db = synth_db.SynthDB()
    if isinstance(db_dir, dict):
        db_dir = db_dir[synth_db.CASES[name].get("db", bool(synth_db.CASES[name].get("low_mem", False)))]
    args = synth_db.make_case_inputs(db, name, work)
    out = os.path.join(work, "out_%s_%s" % (engine, name))
    opts = []
    if "pmix" in synth_db.CASES[name]:
        pdir = os.path.join(work, "DB_plasmid_src")
        if not os.path.isdir(pdir):
            synth_db.write_plasmid_db(work)
        opts = ["--plasmid-db", pdir]
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_pipeline.py"), "--engine", engine] + opts + ["--"] + args + \
          ["-d", db_dir, "-o", out] + synth_db.CASES[name]["flags"]
    log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log.returncode != 0:
        raise RuntimeError("pipeline failed (%s, %s):\n%s" % (engine, name, log.stdout[-4000:]))
    return collect_reports(out), log.stdout


def main():
    names = sys.argv[1:] or sorted(synth_db.CASES)
    with tempfile.TemporaryDirectory() as work:
        db_dir = synth_db.write_dbs(work)
        for name in names:
            reports, log = run_case("reference", db_dir, name, work)
            dst = os.path.join(HERE, "pipeline", name)
            shutil.rmtree(dst, ignore_errors=True)
            for rel, data in reports.items():
                os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                open(os.path.join(dst, rel), "wb").write(data)
            print("== %s: %s" % (name, sorted(reports)))
            for rel in sorted(reports):
                print("--", rel)
                print(reports[rel].decode()[:1500])


if __name__ == "__main__":
    main()
