"""CPU check that the committed pipeline goldens are reproducible from the reference sandbox
(skipped where baseline/_ref is absent, e.g. a fresh clone)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from tests import synth_db  # noqa: E402
import make_pipeline_golden as mpg  # noqa: E402

HAVE_REF = os.access(os.path.join(ROOT, "baseline", "_ref", "library", "jellyfish-linux"), os.X_OK)


def test_goldens_cover_every_case():
    for name in synth_db.CASES:
        g = mpg.collect_reports(os.path.join(ROOT, "tests", "golden", "pipeline", name))
        assert "final_report.txt" in g, name
    assert "strain_prob.txt" in mpg.collect_reports(os.path.join(ROOT, "tests", "golden", "pipeline", "low_depth_prob"))


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not present")
def test_reference_reproduces_golden(tmp_path):
    db_dir = synth_db.SynthDB().write(str(tmp_path / "DB"))
    got, log = mpg.run_case("reference", db_dir, "same_cluster_two_strains", str(tmp_path))
    assert got == mpg.collect_reports(os.path.join(ROOT, "tests", "golden", "pipeline", "same_cluster_two_strains")), log[-2000:]
