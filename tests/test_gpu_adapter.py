"""The drop-in boundary itself: the reference's stock `jellyfish_count(fq_path, db_dir)` (library/identify.py:73-103,
run from the sandbox baseline/_ref with its bundled Jellyfish and its own Python dump parse) and
strainscan_b200.identify_shim.jellyfish_count must return the SAME mapping -- every key (valid k-mers including
zero counts, "last duplicate wins"), every count -- on a synthetic search-tree database in the on-disk layout and
plain / gzip'ed / paired read files.  Needs baseline/_ref (travels to the GPU box); skipped when absent."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
HAVE_REF = os.access(os.path.join(REF, "library", "jellyfish-linux"), os.X_OK)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not present (python baseline/setup_ref.py)")
def test_jellyfish_count_mapping_identical_to_the_stock_function(tmp_path, monkeypatch):
    import torch
    from strainscan_b200 import Engine, identify_shim, synth
    eng = Engine(0)
    params = synth.default_params(n_leaves=60, genome_len=1_000_000, seed=5)
    sizes = synth.node_sizes(params, seed=5)
    db_text, _ = eng.synth_db_host(params, sizes, want_nodes=False)
    db = tmp_path / "Tree_database"
    db.mkdir()
    # a few adapter-level oddities on top of the generator's records: duplicate, lowercase, N, short
    text = db_text.tobytes()
    first = text.split(b"\n")[1]
    text += b">1\n" + first + b"\n>1\n" + first.lower() + b"\n>1\n" + first[:-1] + b"N\n>1\n" + first[:-3] + b"\n"
    (db / "kmer.fa").write_bytes(text)
    n_reads = 300_000
    rec = eng.synth_read_record_bytes(params)
    buf = torch.empty(n_reads * rec, dtype=torch.uint8, device="cuda")
    eng.synth_reads_device(params, buf.data_ptr(), n_reads, 0)
    fq = buf.cpu().numpy().tobytes()
    eng.close()
    half = (n_reads // 2) * rec
    (tmp_path / "r.fq").write_bytes(fq)
    (tmp_path / "r1.fq.gz").write_bytes(gzip.compress(fq[:half], 1))
    (tmp_path / "r2.fq.gz").write_bytes(gzip.compress(fq[half:], 1))

    sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), REF, os.path.join(REF, "library")]
    monkeypatch.chdir(tmp_path)                      # the reference writes temp_<uuid>.jf/.fa into cwd
    from library import identify as ref_identify
    for fq_path in ((str(tmp_path / "r.fq"), ""), (str(tmp_path / "r1.fq.gz"), str(tmp_path / "r2.fq.gz"))):
        ref = ref_identify.jellyfish_count(fq_path, str(db))
        identify_shim.drop_caches()
        cv = identify_shim.jellyfish_count(fq_path, str(db))
        keys = np.fromiter(ref.keys(), dtype=np.int64, count=len(ref))
        vals = np.fromiter(ref.values(), dtype=np.int64, count=len(ref))
        assert len(ref) == len(cv) == int(cv.valid_mask.sum())
        assert bool(cv.valid_mask[keys].all())
        assert np.array_equal(cv.counts[keys].astype(np.int64), vals)
        assert set(ref.keys()) == set(cv.valid_kmers())
        assert vals.sum() > 10_000
    identify_shim.drop_caches()
