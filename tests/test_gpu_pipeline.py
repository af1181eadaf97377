"""End-to-end report parity: the reference's own StrainScan.py host code (baseline/_ref, unmodified
apart from the two documented call-site edits of INTEGRATION.md, applied at import time by
baseline/run_pipeline.py) driven by the CUDA counts must write final_report.txt,
C*/StrainVote.report and strain_prob.txt BYTE-IDENTICAL to the reference run with its bundled
jellyfish-linux (golden reports in tests/golden/pipeline/, made by tests/golden/make_pipeline_golden.py).

Covers single-end / paired-end + gz, two clusters, two strains in one cluster (ElasticNet path),
an all-singleton result, -l 2 -b 1 (low depth + probability report), -e 1 (extraRegion_mode), a
memory-efficient database (Memory_DB -> identify_low_mem, canonical-strand k-mer set), paired
blocked-gzip inputs (inflated on the device) and plasmid_mode -p 1 / -p 2 (the run-time StrainScan_build.py
call is answered with a pre-made DB_plasmid, baseline/run_pipeline.py --plasmid-db) and a reconstructed
node whose k-mer list overlaps an identified cluster's (adjust_profile, identify.py:167-191).
A second test repeats every case with the per-node and per-strain reductions (match_node, cal_cov_all,
get_candidate_arr, get_remainc) replaced by the GPU-backed mirrors as well (--engine b200-full).
Needs baseline/_ref (git-ignored, travels to the GPU box with gpurun); skipped when it is absent."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from tests import synth_db  # noqa: E402
import make_pipeline_golden as mpg  # noqa: E402

pytestmark = pytest.mark.gpu
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "library"))


@pytest.fixture(scope="module")
def db_dir(tmp_path_factory):
    return synth_db.write_dbs(str(tmp_path_factory.mktemp("db")))


def _golden(name):
    base = os.path.join(ROOT, "tests", "golden", "pipeline", name)
    return mpg.collect_reports(base)


def _same_up_to_float_ulps(a, b, rel=1e-9):
    """Token-wise equality; numeric tokens may differ in the last ulps (goldens were written on another
    CPU: sklearn's coordinate descent / BLAS round differently across SIMD widths)."""
    ta, tb = a.decode().replace("/", "\t").split(), b.decode().replace("/", "\t").split()
    if len(ta) != len(tb):
        return False
    for x, y in zip(ta, tb):
        if x == y:
            continue
        try:
            fx, fy = float(x), float(y)
        except ValueError:
            return False
        if abs(fx - fy) > rel * max(abs(fx), abs(fy)):
            return False
    return True


LIVE = HAVE_REF and os.access(os.path.join(ROOT, "baseline", "_ref", "library", "jellyfish-linux"), os.X_OK)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not present (python baseline/setup_ref.py)")
@pytest.mark.parametrize("name", sorted(synth_db.CASES))
def test_reports_identical_to_reference(name, db_dir, tmp_path):
    gold = _golden(name)
    assert "final_report.txt" in gold, "golden reports missing for %s" % name
    got, log = mpg.run_case("b200", db_dir, name, str(tmp_path))
    assert sorted(got) == sorted(gold), log[-2000:]
    # (1) against the committed goldens (reference run in the build container): every integer and
    #     string identical, floats to 1e-9 (different host CPU)
    for rel in gold:
        assert _same_up_to_float_ulps(got[rel], gold[rel]), "%s/%s differs\n--- reference\n%s\n--- b200\n%s" % (
            name, rel, gold[rel].decode(), got[rel].decode())
    # (2) against the reference pipeline run live on THIS machine: byte-identical
    if LIVE:
        ref, rlog = mpg.run_case("reference", db_dir, name, str(tmp_path))
        assert sorted(ref) == sorted(got), rlog[-2000:]
        for rel in ref:
            assert got[rel] == ref[rel], "%s/%s not byte-identical\n--- reference (live)\n%s\n--- b200\n%s" % (
                name, rel, ref[rel].decode(), got[rel].decode())


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not present (python baseline/setup_ref.py)")
@pytest.mark.parametrize("name", sorted(synth_db.CASES))
def test_reports_identical_with_gpu_reducers(name, db_dir, tmp_path):
    """SURVEY 8a rows a5-a7, a11, a12 inside the pipeline: the reference's decisions consume the mirrors'
    numbers and must write the same reports."""
    gold = _golden(name)
    got, log = mpg.run_case("b200-full", db_dir, name, str(tmp_path))
    assert sorted(got) == sorted(gold), log[-2000:]
    for rel in gold:
        assert _same_up_to_float_ulps(got[rel], gold[rel]), "%s/%s differs\n--- reference\n%s\n--- b200-full\n%s" % (
            name, rel, gold[rel].decode(), got[rel].decode())
