"""The oracle (oracle/) against the golden vectors produced by the reference's own engine
(jellyfish-linux 2.3.0, tests/golden/make_golden.py), and against the live binary when
oracle/_ref/ is present.  CPU only."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import adapters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JF = os.path.join(ROOT, "oracle", "_ref", "jellyfish-linux")


def _dense_to_dump(case, d):
    """Rebuild {KMER: count} for the records the adapter view can see."""
    lines = case["fasta"].split("\n")
    dump = {}
    for i in range(len(d.cnt)):
        if d.in_set[i]:
            dump[lines[2 * i + 1].rstrip().upper()] = int(d.cnt[i])
    return dump


def test_golden_present(golden_cases):
    names = {c["name"] for c in golden_cases}
    for need in ("strand", "nbreak_lower", "if_oddities", "short_reads", "crlf", "qual_at", "pe_files",
                 "pe_zcat", "wrapped_fastq", "fasta_reads", "l2_k21", "k32", "k11", "polyT_k32", "lowcomplex",
                 "medium_random"):
        assert need in names


def test_c_oracle_matches_jellyfish_golden(golden_cases):
    for case in golden_cases:
        d = adapters.count_dense(case["fasta"].encode(), case["k"], [r.encode() for r in case["reads"]])
        got = _dense_to_dump(case, d)
        gold = case["dump"]
        # every record string of length k is a dumped key with the same count
        for kmer, c in got.items():
            assert gold[kmer] == c, (case["name"], kmer)
        # dumped keys the adapter view cannot see are windows of over-long records only
        extra = set(gold) - set(got)
        if case["name"] != "if_oddities":
            assert not extra, case["name"]
        assert d.n_distinct == len(gold), case["name"]


def test_py_oracle_matches_jellyfish_golden(golden_cases):
    for case in golden_cases:
        if case["name"] == "medium_random":
            continue  # covered by the C oracle; keep the pure-Python loop tiny
        dump = adapters.py_count_tiny(case["fasta"], case["k"], case["reads"])
        assert dump == case["dump"], case["name"]


def test_py_oracle_medium(golden_cases):
    case = [c for c in golden_cases if c["name"] == "medium_random"][0]
    dump = adapters.py_count_tiny(case["fasta"], case["k"], case["reads"])
    assert dump == case["dump"]


def test_l1_adapter_views_agree(golden_cases):
    """identify.py:90-101 built from the golden dump == the dense oracle view."""
    for case in golden_cases:
        if case["name"] == "if_oddities":
            continue  # reference raises KeyError there (windows of an over-long record are dumped)
        d = adapters.count_dense(case["fasta"].encode(), case["k"], [r.encode() for r in case["reads"]])
        assert adapters.l1_match_results(d) == adapters.l1_match_results_from_dump(case["fasta"], case["dump"])


def test_l1_last_duplicate_wins(golden_cases):
    case = [c for c in golden_cases if c["name"] == "if_oddities"][0]
    d = adapters.count_dense(case["fasta"].encode(), case["k"], [r.encode() for r in case["reads"]])
    lines = case["fasta"].split("\n")
    recs = [lines[2 * i + 1] for i in range(len(d.cnt))]
    # records 0 and 12 are the same string, 1 and 11, 2 and 6 (lowercase twin)
    assert recs[0] == recs[12] and recs[1] == recs[11] and recs[2].lower() == recs[6]
    assert d.is_last[12] and not d.is_last[0]
    assert d.is_last[11] and not d.is_last[1]
    assert d.is_last[6] and not d.is_last[2]
    assert d.cnt[0] == d.cnt[12] and d.cnt[2] == d.cnt[6]
    # N record, short record, long record are not dumped keys
    assert not d.in_set[7] and not d.in_set[8] and not d.in_set[9]
    # lowercase record is a key after folding but not as a raw string (L2 adapter view)
    assert d.in_set[6] and not d.raw_upper[6]
    # absent poly-G: seeded, zero count, still valid
    assert d.in_set[10] and d.cnt[10] == 0


def test_l2_adapter_views_agree(golden_cases):
    for case in golden_cases:
        if not case["fasta"].startswith(">1\n") or case["fasta"].count(">1\n") > 1:
            continue
        d = adapters.count_dense(case["fasta"].encode(), case["k"], [r.encode() for r in case["reads"]])
        a = adapters.l2_py_o(d)
        b = adapters.l2_py_o_from_dump(case["fasta"], case["dump"])
        assert np.array_equal(a, b), case["name"]
        if case["name"] == "l2_k21":
            assert (a == 1).sum() == 0 and (a >= 2).sum() > 0
            assert (np.array(list(case["dump"].values())) == 1).sum() > 0   # remove_1 did something


def test_window_count():
    reads = b"@a\nACGTACGTAC\n+\nIIIIIIIIII\n@b\nACGNACGTACGT\n+\nIIIIIIIIIIII\n"
    assert adapters.count_windows([reads], 4) == 7 + 0 + 5


@pytest.mark.skipif(not os.path.exists(JF), reason="oracle/_ref/jellyfish-linux not built here")
def test_c_oracle_matches_live_jellyfish():
    rng = np.random.default_rng(99)
    G = "".join("ACGT"[i] for i in rng.integers(0, 4, 50000))
    pos = rng.choice(50000 - 31, 8000, replace=False)
    fa = "".join(">1\n%s\n" % G[p:p + 31] for p in pos)
    reads = []
    for i in range(3000):
        s = int(rng.integers(0, 50000 - 150))
        r = list(G[s:s + 150])
        for j in np.nonzero(rng.random(150) < 0.01)[0]:
            r[j] = "ACGTN"[int(rng.integers(0, 5))]
        reads.append("@r%d\n%s\n+\n%s\n" % (i, "".join(r), "I" * 150))
    fq = "".join(reads)
    with tempfile.TemporaryDirectory() as td:
        open(td + "/k.fa", "w").write(fa)
        open(td + "/r.fq", "w").write(fq)
        subprocess.check_call([JF, "count", "-m", "31", "-s", "100M", "-t", "8", "--if", td + "/k.fa",
                               "-o", td + "/o.jf", td + "/r.fq"])
        txt = subprocess.check_output([JF, "dump", "-c", td + "/o.jf"]).decode()
    gold = {a: int(b) for a, b in (l.split(" ") for l in txt.splitlines())}
    d = adapters.count_dense(fa.encode(), 31, [fq.encode()])
    assert d.n_distinct == len(gold)
    lines = fa.split("\n")
    for i in range(len(d.cnt)):
        assert d.in_set[i] and gold[lines[2 * i + 1]] == int(d.cnt[i])
