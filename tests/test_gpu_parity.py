"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors of
the reference's own engine.  Integer work: every comparison is bit-exact.  Run with -m gpu."""
import gzip
import os

import numpy as np
import pytest

from oracle import adapters
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from strainscan_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _gold_dense(case):
    return adapters.count_dense(case["fasta"].encode(), case["k"], [r.encode() for r in case["reads"]])


def _check_against_dump(case, got, flags):
    """Direct check against the jellyfish dump, not via the oracle."""
    lines = case["fasta"].split("\n")
    for i in range(len(got)):
        s = lines[2 * i + 1].rstrip().upper()
        if flags[i] & 1:
            assert case["dump"][s] == int(got[i]), (case["name"], i)
        else:
            assert got[i] == 0


def test_golden_resident(eng, golden_cases):
    for case in golden_cases:
        ks = eng.kmerset_from_text(case["fasta"], case["k"])
        reads = eng.reads_from_host([r.encode() for r in case["reads"]])
        got, st = eng.count(ks, reads)
        _check_against_dump(case, got, ks.flags)
        d = _gold_dense(case)
        assert np.array_equal(got.astype(np.uint64), d.cnt), case["name"]
        assert np.array_equal(ks.flags & 1, d.in_set), case["name"]
        assert np.array_equal((ks.flags >> 1) & 1, d.is_last), case["name"]
        assert np.array_equal((ks.flags >> 2) & 1, d.raw_upper), case["name"]
        assert np.array_equal(ks.header_ids, d.header_id), case["name"]
        assert st.n_kmers == adapters.count_windows([r.encode() for r in case["reads"]], case["k"]), case["name"]
        if case["name"] != "if_oddities":
            assert ks.n_distinct == d.n_distinct, case["name"]


def test_golden_streaming_host(eng, golden_cases):
    for case in golden_cases:
        ks = eng.kmerset_from_text(case["fasta"], case["k"])
        got, st = eng.count_host(ks, [r.encode() for r in case["reads"]])
        _check_against_dump(case, got, ks.flags)


def test_golden_files_plain_and_gz(eng, golden_cases, tmp_path):
    for case in golden_cases:
        fa = tmp_path / ("%s.fa" % case["name"])
        fa.write_bytes(case["fasta"].encode())
        plain, gz = [], []
        for i, r in enumerate(case["reads"]):
            p = tmp_path / ("%s_%d.fq" % (case["name"], i))
            p.write_bytes(r.encode())
            plain.append(str(p))
            g = tmp_path / ("%s_%d.fq.gz" % (case["name"], i))
            with gzip.open(g, "wb") as f:
                f.write(r.encode())
            gz.append(str(g))
        ks = eng.kmerset_from_fasta(str(fa), case["k"])
        for paths in (plain, gz):
            got, _ = eng.count_files(ks, paths)
            _check_against_dump(case, got, ks.flags)
            reads = eng.reads_from_files(paths)
            got2, _ = eng.count(ks, reads)
            assert np.array_equal(got, got2)


def test_unsupported_inputs_fail_loudly(eng, golden_cases):
    from strainscan_b200 import StrainScanB200Error
    case = [c for c in golden_cases if c["name"] == "strand"][0]
    ks = eng.kmerset_from_text(case["fasta"], case["k"])
    bad_inputs = [
        b"hello world\n",                                           # not a sequence file
        b"@r1\nACGT\n+\nIIII\n@r2\nACGT\nACGT\n+\nIIIIIIII\n",       # 4-line head, wrapped record later: framing breaks
        b"@r1\nACGT\nACGT\n+\nIIII\n",                              # wrapped, quality shorter than the sequence
    ]
    for text in bad_inputs:
        with pytest.raises(StrainScanB200Error) as ei:
            eng.count(ks, eng.reads_from_host([text]))
        assert ei.value.code == 4, text   # SS_ERR_FORMAT
        with pytest.raises(StrainScanB200Error) as ei:
            eng.count_host(ks, [text])
        assert ei.value.code == 4, text
    with pytest.raises(StrainScanB200Error):
        eng.kmerset_from_text(">1\nACGT\n", 33)
    with pytest.raises(StrainScanB200Error):
        eng.kmerset_from_fasta("/nonexistent/kmer.fa", 31)


@pytest.mark.parametrize("k,filt", [(11, None), (21, None), (31, None), (32, None), (31, "16"), (21, "4"), (32, "6")])
def test_random_vs_oracle(eng, k, filt, monkeypatch):
    """filt: force the L2-resident prefilter on with that many bits per key (4 = many false
    positives, which must all be rejected by the exact table probe)."""
    if filt:
        monkeypatch.setenv("SS_FILTER", "1")
        monkeypatch.setenv("SS_FILTER_BITS", filt)
    else:
        monkeypatch.setenv("SS_FILTER", "0")          # the path without a filter: every k-mer probes the table
    rng = np.random.default_rng(1000 + k)
    G = util.rand_genome(rng, 200_000)
    fa = util.make_db(rng, G, k, 20_000, both_strands=True, lower_frac=0.02, junk=40)
    fq1 = util.make_reads(rng, G, 6000, 150, var_len=False, lower_frac=0.05, offtarget=0.3)
    fq2 = util.make_reads(rng, G, 3000, 100, var_len=True, crlf=(k == 21))
    ks = eng.kmerset_from_text(fa, k)
    d = adapters.count_dense(fa, k, [fq1, fq2])
    for got, st in (eng.count(ks, eng.reads_from_host([fq1, fq2])), eng.count_host(ks, [fq1, fq2])):
        assert np.array_equal(got.astype(np.uint64), d.cnt)
        assert st.n_kmers == adapters.count_windows([fq1, fq2], k)
        assert st.n_reads == 9000
        if filt:
            assert st.n_hits <= st.n_table_probes < st.n_kmers    # the filter rejected most misses
        else:
            assert st.n_table_probes == 0
    assert np.array_equal(ks.flags & 1, d.in_set)
    assert np.array_equal((ks.flags >> 1) & 1, d.is_last)
    assert d.cnt.sum() > 1000


def test_golden_with_filter_forced(eng, golden_cases, monkeypatch):
    monkeypatch.setenv("SS_FILTER", "1")
    for case in golden_cases:
        ks = eng.kmerset_from_text(case["fasta"], case["k"])
        got, st = eng.count(ks, eng.reads_from_host([r.encode() for r in case["reads"]]))
        _check_against_dump(case, got, ks.flags)
        assert st.n_kmers == adapters.count_windows([r.encode() for r in case["reads"]], case["k"]), case["name"]


def test_poly_t_32mer(eng):
    """k = 32 all-T packs to 0xFFFF...F, the table's empty marker: it lives in a side slot."""
    fa = b">1\n" + b"T" * 32 + b"\n>1\n" + b"A" * 32 + b"\n>1\n" + b"T" * 31 + b"G\n"
    fq = b"@a\n" + b"T" * 40 + b"\n+\n" + b"I" * 40 + b"\n@b\n" + b"A" * 33 + b"T" * 31 + b"G\n+\n" + b"I" * 65 + b"\n"
    ks = eng.kmerset_from_text(fa, 32)
    got, _ = eng.count(ks, eng.reads_from_host(fq))
    d = adapters.count_dense(fa, 32, [fq])
    assert np.array_equal(got.astype(np.uint64), d.cnt) and got[0] == 9 and got[1] == 2 and got[2] == 1
    assert ks.n_distinct == 3


def test_empty_and_degenerate(eng):
    ks = eng.kmerset_from_text(b">1\nACGTACGTACGTACGTACGTACGTACGTACG\n", 31)
    for fq in (b"", b"\n\n", b"@r\n\n+\n\n", b"@r\nACGT\n+\nIIII"):
        got, st = eng.count(ks, eng.reads_from_host(fq))
        assert got.tolist() == [0] and st.n_kmers == 0
        got, st = eng.count_host(ks, [fq])
        assert got.tolist() == [0]
    empty = eng.kmerset_from_text(b"", 31)
    assert empty.n_records == 0
    got, st = eng.count(empty, eng.reads_from_host(b"@r\n" + b"ACGT" * 9 + b"\n+\n" + b"I" * 36 + b"\n"))
    assert got.size == 0 and st.n_kmers == 6
    # no trailing newline on the last record, and on the first of two files
    fq = b"@r\nACGTACGTACGTACGTACGTACGTACGTACGA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII"
    got, _ = eng.count(ks, eng.reads_from_host([fq, fq]))
    assert got.tolist() == [2]


def test_linearity_and_sharding(eng, tmp_path):
    """counts(A + B) == counts(A) + counts(B); record-aligned shards of one file sum to the whole."""
    rng = np.random.default_rng(5)
    G = util.rand_genome(rng, 100_000)
    fa = util.make_db(rng, G, 31, 10_000)
    fq = util.make_reads(rng, G, 8000, 150, var_len=True)
    p = tmp_path / "r.fq"
    p.write_bytes(fq)
    ks = eng.kmerset_from_text(fa, 31)
    whole, st = eng.count_files(ks, [str(p)])
    for n in (2, 3, 8):
        acc = np.zeros_like(whole)
        kmers = 0
        for s in range(n):
            part, sp = eng.count_files(ks, [str(p)], shard=s, n_shards=n)
            acc += part
            kmers += sp.n_kmers
        assert np.array_equal(acc, whole) and kmers == st.n_kmers
    d = adapters.count_dense(fa, 31, [fq])
    assert np.array_equal(whole.astype(np.uint64), d.cnt)


def test_l2_finalize_and_reducers(eng):
    import torch
    rng = np.random.default_rng(77)
    k = 31
    G = util.rand_genome(rng, 60_000)
    fa = util.make_db(rng, G, k, 6000, both_strands=True, header=None, lower_frac=0.01)
    fq = util.make_reads(rng, G, 4000, 150)
    ks = eng.kmerset_from_text(fa, k)
    reads = eng.reads_from_host(fq)
    dev = torch.zeros(ks.n_records, dtype=torch.int32, device="cuda:0")
    eng.count_device(ks, reads, dev.data_ptr())
    d = adapters.count_dense(fa, k, [fq])
    assert np.array_equal(dev.cpu().numpy().astype(np.uint64), d.cnt)
    # L2 adapter (remove_1, kid order)
    py_o = eng.l2_finalize(ks, dev.data_ptr())
    assert np.array_equal(py_o, adapters.l2_py_o(d))
    assert (py_o == 1).sum() == 0 and (py_o > 1).sum() > 100
    # per-node reducer vs match_node
    mr = adapters.l1_match_results(d)
    valid = set(mr)
    n_nodes = 9
    lists = [np.unique(rng.integers(0, ks.n_records, int(rng.integers(1, 3000)))) for _ in range(n_nodes)]
    lists[3] = np.zeros(0, dtype=np.int64)
    ptr = np.zeros(n_nodes + 1, dtype=np.uint64)
    ptr[1:] = np.cumsum([len(x) for x in lists])
    length, covered, total = eng.node_reduce(ks, dev.data_ptr(), ptr, np.concatenate(lists))
    for v in range(n_nodes):
        vk = valid & set(int(x) for x in lists[v])
        prof = [mr[x] for x in vk if mr[x] > 0]
        assert length[v] == len(vk) and covered[v] == len(prof) and total[v] == sum(prof)
    # per-strain reducer vs stat_cov / get_remainc / get_candidate_arr
    n_rows, n_str = ks.n_records, 7
    X = (rng.random((n_rows, n_str)) < 0.3).astype(np.int8)
    used = (rng.random(n_rows) < 0.4).astype(np.uint8)
    col_ptr = np.zeros(n_str + 1, dtype=np.uint64)
    rows = []
    for j in range(n_str):
        r = np.nonzero(X[:, j])[0]
        rows.append(r)
        col_ptr[j + 1] = col_ptr[j] + len(r)
    rows = np.concatenate(rows)
    tot, cov, _ = eng.strain_reduce(col_ptr, rows, py_o)
    ref = adapters.stat_cov_all(X, py_o)
    assert [(int(c), int(t)) for c, t in zip(cov, tot)] == ref
    assert [int(c) for c in cov] == adapters.candidate_counts(X, py_o)
    tot2, cov2, _ = eng.strain_reduce(col_ptr, rows, py_o, row_mask=1 - used)
    assert [(int(c), int(t)) for c, t in zip(cov2, tot2)] == adapters.remain_cov(used, X, py_o)


def test_synthetic_workload_sample_vs_oracle(eng):
    """The bench's generators: a small instance end to end against the oracle, plus the
    size-independent invariants used at full size (sum of counts == hits; halves add up)."""
    import torch
    from strainscan_b200 import synth
    p = synth.default_params(n_leaves=16, genome_len=200_000, seed=11)
    sizes = synth.node_sizes(p, lo=200, hi=2000, seed=3)
    text, node_of = eng.synth_db_host(p, sizes)
    ks = eng.kmerset_from_text(text, p.k)
    assert ks.n_records == int(sizes.sum())
    n_reads = 20_000
    rec = eng.synth_read_record_bytes(p)
    cap = eng.reads_device_capacity(n_reads * rec)
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
    eng.synth_reads_device(p, buf.data_ptr(), n_reads, 0)
    reads = eng.reads_from_device(buf.data_ptr(), n_reads * rec, cap, keepalive=buf)
    got, st = eng.count(ks, reads)
    fq = buf[:n_reads * rec].cpu().numpy().tobytes()
    d = adapters.count_dense(text.tobytes(), p.k, [fq])
    assert np.array_equal(got.astype(np.uint64), d.cnt)
    assert st.n_reads == n_reads and int(got[ks.valid].sum()) == st.n_hits and st.n_hits > 1000
    # the same reads generated in two halves
    half = n_reads // 2
    acc = np.zeros_like(got)
    for first in (0, half):
        b2 = torch.empty(eng.reads_device_capacity(half * rec), dtype=torch.uint8, device="cuda:0")
        eng.synth_reads_device(p, b2.data_ptr(), half, first)
        r2 = eng.reads_from_device(b2.data_ptr(), half * rec, b2.numel(), keepalive=b2)
        acc += eng.count(ks, r2)[0]
    assert np.array_equal(acc, got)
    # hits land on the nodes of the source leaves' root paths only
    hit_nodes = set(np.unique(node_of[got > 0]).tolist())
    allowed = set()
    for i in range(p.n_sources):
        v = p.n_leaves - 1 + p.source_leaf[i]
        while True:
            allowed.add(v)
            if v == 0:
                break
            v = (v - 1) // 2
    assert hit_nodes and hit_nodes <= allowed


def test_file_ingest_many_chunks_and_segments(tmp_path, monkeypatch):
    """Read files through the producer-thread ingest with tiny chunks (256 KiB) and tiny read-cache
    segments (1 MiB): paired plain + gz inputs, the resident cache (ss_reads_from_files, several
    segments -> several probe launches) and the streaming driver (ss_count_files), whole and sharded."""
    from strainscan_b200 import Engine
    monkeypatch.setenv("SS_CHUNK_BYTES", str(256 << 10))
    monkeypatch.setenv("SS_SEG_BYTES", str(1 << 20))
    monkeypatch.setenv("SS_INGEST_THREADS", "4")
    e = Engine(0)
    try:
        rng = np.random.default_rng(77)
        G = util.rand_genome(rng, 150_000)
        fa = util.make_db(rng, G, 31, 15_000)
        fq1 = util.make_reads(rng, G, 12_000, 150, var_len=True)         # ~4 MB
        fq2 = util.make_reads(rng, G, 9_000, 150, var_len=False, lower_frac=0.1)
        p1, p2 = tmp_path / "r1.fq.gz", tmp_path / "r2.fq"
        p1.write_bytes(gzip.compress(fq1, 6))
        p2.write_bytes(fq2)
        ks = e.kmerset_from_text(fa, 31)
        d = adapters.count_dense(fa, 31, [fq1, fq2])
        reads = e.reads_from_files([str(p1), str(p2)])
        got, st = e.count(ks, reads)
        assert np.array_equal(got.astype(np.uint64), d.cnt)
        assert st.n_reads == 21_000 and st.probe_launches >= 4          # one launch per cache segment
        # the same cache of several segments in binned mode (rounds per segment, the sample taken from the first one)
        monkeypatch.setenv("SS_FILTER", "1")
        ksf = e.kmerset_from_text(fa, 31)
        monkeypatch.delenv("SS_FILTER")
        for k, v in (("SS_BIN", "2"), ("SS_BIN_SLICE_MB", "0.1"), ("SS_BIN_POOL_MB", "64"), ("SS_BIN_ROUND_TILES", "400")):
            monkeypatch.setenv(k, v)
        gotb, stb = e.count(ksf, reads)
        assert np.array_equal(gotb, got) and stb.binned_rounds >= 8 and stb.n_reads == 21_000 and stb.n_kmers == st.n_kmers
        for k in ("SS_BIN", "SS_BIN_SLICE_MB", "SS_BIN_POOL_MB", "SS_BIN_ROUND_TILES"):
            monkeypatch.delenv(k)
        got2, st2 = e.count_files(ks, [str(p1), str(p2)])
        assert np.array_equal(got2, got) and st2.n_kmers == st.n_kmers and st2.probe_launches >= 16
        for n in (2, 5):
            acc = np.zeros_like(got)
            acc_c = np.zeros_like(got)
            for s in range(n):
                acc += e.count_files(ks, [str(p1), str(p2)], shard=s, n_shards=n)[0]
                acc_c += e.count(ks, e.reads_from_files([str(p1), str(p2)], shard=s, n_shards=n))[0]
            assert np.array_equal(acc, got) and np.array_equal(acc_c, got)
        # errors surface from the producer threads
        bad = tmp_path / "bad.fq.gz"
        bad.write_bytes(gzip.compress(fq1)[:100_000])
        with pytest.raises(Exception, match="inflate failed"):
            e.reads_from_files([str(bad)])
        with pytest.raises(Exception, match="inflate failed"):
            e.count_files(ks, [str(p2), str(bad)])
        assert np.array_equal(e.count_files(ks, [str(p1), str(p2)])[0], got)   # the context stays usable
    finally:
        e.close()


def test_bgzf_device_inflate(tmp_path, monkeypatch):
    """Blocked-gzip read files: the members are inflated by ss_gunzip_kernel on the device (compressed
    bytes cross PCIe), both into the resident read cache and in the streaming driver; counts must equal
    the oracle's on the plain text, whole and sharded, and equal the host-inflate path (SS_BGZF_GPU=0)."""
    from strainscan_b200 import Engine
    monkeypatch.setenv("SS_CHUNK_BYTES", str(1 << 20))
    monkeypatch.setenv("SS_BGZF_OUT_CAP", str(1 << 20))
    monkeypatch.setenv("SS_SEG_BYTES", str(2 << 20))
    rng = np.random.default_rng(99)
    G = util.rand_genome(rng, 150_000)
    fa = util.make_db(rng, G, 31, 15_000)
    fq1 = util.make_reads(rng, G, 12_000, 150, var_len=True)
    fq2 = util.make_reads(rng, G, 9_000, 150, lower_frac=0.1)
    p1, p2, p3 = tmp_path / "r1.fq.gz", tmp_path / "r2.fq.gz", tmp_path / "r3.fq.gz"
    p1.write_bytes(util.bgzf_compress(fq1, level=6))
    p2.write_bytes(util.bgzf_compress(fq2[:-1], block=7000, level=1, eof_marker=False, empty_every=5))
    p3.write_bytes(gzip.compress(fq2))                         # an ordinary gzip stream beside a BGZF file
    d = adapters.count_dense(fa, 31, [fq1, fq2])
    e = Engine(0)
    try:
        ks = e.kmerset_from_text(fa, 31)
        got, st = e.count(ks, e.reads_from_files([str(p1), str(p2)]))
        assert np.array_equal(got.astype(np.uint64), d.cnt) and st.n_reads == 21_000
        got2, st2 = e.count_files(ks, [str(p1), str(p2)])
        assert np.array_equal(got2, got) and st2.n_kmers == st.n_kmers
        assert st2.total_launches > 2 * st2.probe_launches      # index + probe per chunk, gather ... plus the inflate launches
        got3, _ = e.count_files(ks, [str(p1), str(p3)])
        assert np.array_equal(got3, got)
        for n in (2, 3):
            acc = np.zeros_like(got)
            for s in range(n):
                acc += e.count(ks, e.reads_from_files([str(p1), str(p2)], shard=s, n_shards=n))[0]
            assert np.array_equal(acc, got)
        # a block whose ISIZE trailer disagrees with its deflate stream is reported by the device (CRC32 is not
        # verified: the reference's `zcat | jellyfish` pipe also counts whatever zcat emitted before its warning)
        bad = bytearray(p1.read_bytes())
        off = 0
        for _ in range(5):
            off += int.from_bytes(bad[off + 16:off + 18], "little") + 1
        bad[off - 4] ^= 0x01
        pb = tmp_path / "bad.fq.gz"
        pb.write_bytes(bytes(bad))
        with pytest.raises(Exception, match="inflate failed"):
            e.count_files(ks, [str(pb)])
        with pytest.raises(Exception, match="inflate failed"):
            e.reads_from_files([str(pb)])
        assert np.array_equal(e.count_files(ks, [str(p1), str(p2)])[0], got)     # the context stays usable
    finally:
        e.close()
    monkeypatch.setenv("SS_BGZF_GPU", "0")
    e = Engine(0)
    try:
        ks = e.kmerset_from_text(fa, 31)
        assert np.array_equal(e.count_files(ks, [str(p1), str(p2)])[0], got)
    finally:
        e.close()


def test_binary_database_cache(eng, golden_cases, tmp_path):
    """the parsed form of a k-mer FASTA cached in binary next to the text.  A set built from the cache
    must be indistinguishable from one parsed from the text (flags, header ids, counts); a stale or foreign cache
    is ignored and rewritten; the FASTA stays the source of truth."""
    rng = np.random.default_rng(8)
    G = util.rand_genome(rng, 100_000)
    cases = [(util.make_db(rng, G, 31, 12_000, lower_frac=0.03, junk=40), 31),
             (util.make_db(rng, G, 21, 5_000, header=None), 21)]            # header ids 1..n (all_kmer.fasta form)
    cases += [(c["fasta"].encode(), c["k"]) for c in golden_cases if c["name"] in ("if_oddities", "polyT_k32", "l2_k21")]
    fq = util.make_reads(rng, G, 4000, 150)
    for i, (fa, k) in enumerate(cases):
        p, cp = tmp_path / ("db%d.fa" % i), str(tmp_path / ("db%d.cache" % i))
        p.write_bytes(fa)
        ks0 = eng.kmerset_from_fasta(str(p), k)
        ks1, hit1 = eng.kmerset_from_fasta_cached(str(p), k, cp)
        ks2, hit2 = eng.kmerset_from_fasta_cached(str(p), k, cp)
        assert (hit1, hit2) == (False, True) and os.path.getsize(cp) < len(fa)
        reads = eng.reads_from_host([fq])
        want = eng.count(ks0, reads)[0]
        for ks in (ks1, ks2):
            assert np.array_equal(ks.flags, ks0.flags) and np.array_equal(ks.header_ids, ks0.header_ids)
            assert ks.n_distinct == ks0.n_distinct and ks.n_records == ks0.n_records
            assert np.array_equal(eng.count(ks, reads)[0], want)
        # another k, a changed FASTA, a truncated cache: not used
        assert eng.kmerset_from_fasta_cached(str(p), k - 2, cp)[1] is False
        assert eng.kmerset_from_fasta_cached(str(p), k, cp)[1] is False          # the k-2 call rewrote it
        p.write_bytes(fa + b">1\n" + G[:k] + b"\n")
        ks3, hit3 = eng.kmerset_from_fasta_cached(str(p), k, cp)
        assert hit3 is False and ks3.n_records == ks0.n_records + 1
        blob = open(cp, "rb").read()
        open(cp, "wb").write(blob[:len(blob) // 2])
        assert eng.kmerset_from_fasta_cached(str(p), k, cp)[1] is False
        assert eng.kmerset_from_fasta_cached(str(p), k, cp)[1] is True


@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 25, 30])
def test_small_and_odd_k(eng, k):
    """Every k the one-word key supports, against the oracle: short k stresses the window masks (nearly every
    position is a valid window, lines shorter than a 32-byte run) and table collisions (4^k < records)."""
    rng = np.random.default_rng(500 + k)
    G = util.rand_genome(rng, 30_000)
    fa = util.make_db(rng, G, k, min(3000, len(G) - k - 1), both_strands=True, junk=12)
    fq = util.make_reads(rng, G, 1500, 80, var_len=True, p_n=0.02, lower_frac=0.1)
    fq += b"@tiny\nA\n+\nI\n@nn\nNNNN\n+\nIIII\n"
    ks = eng.kmerset_from_text(fa, k)
    d = adapters.count_dense(fa, k, [fq])
    got, st = eng.count(ks, eng.reads_from_host([fq]))
    assert np.array_equal(got.astype(np.uint64), d.cnt)
    assert st.n_kmers == adapters.count_windows([fq], k)
    got2, _ = eng.count_host(ks, [fq])
    assert np.array_equal(got2, got)


@pytest.mark.parametrize("k,slice_mb,pool_mb,round_tiles,bin_filter", [
    (31, "0.02", "64", "300", "1"), (21, "0.3", "256", "1500", "0"), (32, "0.02", "0.0625", "300", "1"),
    (31, "0.05", "2", "97", "0"), (32, "0.1", "64", "700", "0")])
def test_binned_mode_vs_oracle(eng, k, slice_mb, pool_mb, round_tiles, bin_filter, monkeypatch):
    """Binned probing (large table + many table probes: survivors of the filter are grouped by table range, then
    probed range after range) forced on at test size: many bins, several rounds per segment, a pool so small
    that bins overflow (the round is then redone with direct probes), with and without the filter in front of
    the bins, against the oracle and the direct mode."""
    monkeypatch.setenv("SS_FILTER", "1")
    rng = np.random.default_rng(4242 + k)
    G = util.rand_genome(rng, 300_000)
    fa = util.make_db(rng, G, k, 60_000, both_strands=True, lower_frac=0.02, junk=40)
    fq1 = util.make_reads(rng, G, 12000, 150, var_len=False, lower_frac=0.05, offtarget=0.1)
    fq2 = util.make_reads(rng, G, 4000, 100, var_len=True)
    poly = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"T" * 150, b"I" * 150) for i in range(300))   # one bin gets everything
    fa += b">1\n" + b"T" * k + b"\n"
    ks = eng.kmerset_from_text(fa, k)
    d = adapters.count_dense(fa, k, [fq1, poly, fq2])
    reads = eng.reads_from_host([fq1, poly, fq2])
    monkeypatch.setenv("SS_BIN", "0")
    direct, st0 = eng.count(ks, reads)
    assert st0.binned_rounds == 0 and np.array_equal(direct.astype(np.uint64), d.cnt)
    monkeypatch.setenv("SS_BIN", "2")
    monkeypatch.setenv("SS_BIN_SLICE_MB", slice_mb)
    monkeypatch.setenv("SS_BIN_POOL_MB", pool_mb)
    monkeypatch.setenv("SS_BIN_ROUND_TILES", round_tiles)
    monkeypatch.setenv("SS_BIN_FILTER", bin_filter)
    got, st = eng.count(ks, reads)
    assert np.array_equal(got.astype(np.uint64), d.cnt)
    assert (st.n_kmers, st.n_reads, st.n_hits) == (st0.n_kmers, st0.n_reads, st0.n_hits)
    if float(pool_mb) >= 64:                   # no overflow: every round after the sample was binned
        assert st.binned_rounds >= 2 and st.bins >= 2
        if bin_filter == "1":
            assert st.n_table_probes == st0.n_table_probes
    # idempotent: a second pass over the same cache gives the same vector
    again, _ = eng.count(ks, reads)
    assert np.array_equal(again, got)


def test_long_reads_vs_oracle(eng, tmp_path):
    """Sequence and quality lines far longer than a 992-byte unit (long-read FASTQ: 3 kb .. 120 kb per line), through
    the resident, the streaming-host and the file paths, with and without the filter."""
    rng = np.random.default_rng(777)
    G = util.rand_genome(rng, 400_000)
    fa = util.make_db(rng, G, 31, 30_000, both_strands=True)
    fq = b"".join(util.make_reads(rng, G, n, L, p_sub=0.02, p_n=0.001) for n, L in ((40, 3000), (6, 40_000), (2, 120_000), (30, 150)))
    d = adapters.count_dense(fa, 31, [fq])
    path = tmp_path / "long.fq"
    path.write_bytes(fq)
    for filt in (None, "1"):
        if filt:
            os.environ["SS_FILTER"] = filt
        try:
            ks = eng.kmerset_from_text(fa, 31)
        finally:
            os.environ.pop("SS_FILTER", None)
        for how, (got, st) in (("resident", eng.count(ks, eng.reads_from_host([fq]))), ("host", eng.count_host(ks, [fq])),
                               ("files", eng.count_files(ks, [str(path)]))):
            assert np.array_equal(got.astype(np.uint64), d.cnt), how
            assert st.n_kmers == adapters.count_windows([fq], 31), (how, st.n_kmers)
            assert st.n_reads == 78, (how, st.n_reads)
    assert d.cnt.sum() > 10_000


def test_line_index_many_blocks_and_rebuild(eng):
    """K1 (one-pass line index: per-unit newline counts + decoupled look-back scan over 512-unit blocks): text of
    several thousand units with ragged line lengths, first pass (index built inside it), a second pass (index kept) and
    a pass after ss_reads_drop_index give the oracle's vector and the same read / k-mer totals."""
    rng = np.random.default_rng(77)
    G = util.rand_genome(rng, 200_000)
    fa = util.make_db(rng, G, 31, 30_000, both_strands=True)
    fq = util.make_reads(rng, G, 30_000, 150, var_len=True)           # ~9 MB: ~9500 units, 19 index blocks
    ks = eng.kmerset_from_text(fa, 31)
    reads = eng.reads_from_host([fq])
    o = adapters.count_dense(fa, 31, [fq])
    got1, st1 = eng.count(ks, reads)
    got2, st2 = eng.count(ks, reads)
    reads.drop_index()
    got3, st3 = eng.count(ks, reads)
    for got, st in ((got1, st1), (got2, st2), (got3, st3)):
        assert np.array_equal(got.astype(np.uint64), o.cnt)
        assert st.n_reads == 30_000 and st.n_kmers == adapters.count_windows([fq], 31)
    assert st1.ms_index > 0 and st2.ms_index == 0 and st3.ms_index > 0


def test_host_generator_equals_device_generator(eng):
    """tools/libss_synth_host.so (what bench.py --impl reference feeds the reference engine) writes the same bytes as the
    device generators the GPU arm uses."""
    import torch
    from strainscan_b200 import synth
    from tools import synth_host
    p = synth.default_params(n_leaves=37, seed=5)
    sizes = synth.node_sizes(p, seed=5)
    db_d, node_d = eng.synth_db_host(p, sizes, want_nodes=True)
    db_h, node_h = synth_host.db(p, sizes, want_nodes=True, threads=3)
    assert np.array_equal(db_d, db_h) and np.array_equal(node_d, node_h)
    n, first = 20_000, 123_456_789
    rec = eng.synth_read_record_bytes(p)
    assert rec == synth_host.read_record_bytes(p)
    buf = torch.empty(n * rec, dtype=torch.uint8, device="cuda")
    eng.synth_reads_device(p, buf.data_ptr(), n, first)
    assert np.array_equal(buf.cpu().numpy(), synth_host.reads(p, n, first, threads=3))


@pytest.mark.parametrize("lanes", ["0", "2ring", "2", "4", "3free"])
@pytest.mark.parametrize("shape", ["single", "members", "level9_small_batches"])
def test_device_inflate_of_ordinary_gzip(eng, tmp_path, monkeypatch, shape, lanes):
    """ss_dgz.cu: ordinary (non-blocked) gzip read files inflated on the device -- block starts found per piece, marker
    symbols, exact stitching, window chain -- through ss_count_files and ss_reads_from_files, against the oracle.  Small
    pieces / batches force many pieces, several batches and records carried from batch to batch; sharded runs split a
    multi-member file at member starts."""
    rng = np.random.default_rng({"single": 1, "members": 2, "level9_small_batches": 3}[shape])
    G = util.rand_genome(rng, 150_000)
    fa = util.make_db(rng, G, 31, 20_000, both_strands=True)
    fq1 = util.make_reads(rng, G, 20_000, 150, var_len=True)                 # ~6 MB each
    fq2 = util.make_reads(rng, G, 18_000, 120, var_len=True, crlf=False)
    monkeypatch.setenv("SS_DGZ_MIN_BYTES", "0")
    monkeypatch.setenv("SS_DGZ_PIECE_BYTES", "16384")
    monkeypatch.setenv("SS_DGZ_SYM_PER_BYTE", "32")
    monkeypatch.setenv("SS_DGZ_BATCH_MB", "8")
    # K8: one decoder per one-warp CTA (0); 2 per warp in lockstep with the shared-memory ring of recent symbols (what
    # ships) or without it; 4 per warp with short rounds; 3 per warp, each at its own pace (ss_dgz2.cuh)
    monkeypatch.setenv("SS_DGZ_LANES", lanes.rstrip("freing"))
    monkeypatch.setenv("SS_DGZ_LOCKSTEP", "0" if lanes.endswith("free") else "1")
    monkeypatch.setenv("SS_DGZ_RING", "1" if lanes.endswith("ring") else "0")
    if lanes == "4":
        monkeypatch.setenv("SS_DGZ_ROUND", "5")          # many trips through the state machine
    p1, p2 = str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fq.gz")
    if shape == "single":
        open(p1, "wb").write(gzip.compress(fq1, 6))
        open(p2, "wb").write(gzip.compress(fq2[:-1], 1))                       # no final newline
    elif shape == "members":
        step = 700_001                                                          # members cut records apart
        open(p1, "wb").write(b"".join(gzip.compress(fq1[i:i + step], 6) for i in range(0, len(fq1), step)))
        open(p2, "wb").write(b"".join(gzip.compress(fq2[i:i + step], 1) for i in range(0, len(fq2), step)) + b"\0\0trailing garbage")
    else:
        monkeypatch.setenv("SS_DGZ_MAX_PIECES", "16")
        open(p1, "wb").write(gzip.compress(fq1 + b"\n\n", 9))                  # blank tail lines
        open(p2, "wb").write(gzip.compress(fq2, 9))
    ks = eng.kmerset_from_text(fa, 31)
    o = adapters.count_dense(fa, 31, [fq1, fq2])
    got, st = eng.count_files(ks, [p1, p2])
    assert np.array_equal(got.astype(np.uint64), o.cnt) and st.n_reads == 38_000
    reads = eng.reads_from_files([p1, p2])
    got2, st2 = eng.count(ks, reads)
    assert np.array_equal(got2, got) and st2.n_reads == 38_000
    # batches decoded while the rest of the file is still being uploaded (small chunks, one reader, hardly any slack
    # behind a batch: decoders run into bytes that are not there yet, report it, and the batch is decoded again)
    from strainscan_b200 import Engine
    monkeypatch.setenv("SS_CHUNK_BYTES", str(256 << 10))
    monkeypatch.setenv("SS_INGEST_THREADS", "1")
    monkeypatch.setenv("SS_DGZ_GATE_SLACK", "4096")
    monkeypatch.setenv("SS_DGZ_MAX_PIECES", "8")
    eng2 = Engine(0)
    ks2 = eng2.kmerset_from_text(fa, 31)
    got_s, st_s = eng2.count_files(ks2, [p1, p2])
    assert np.array_equal(got_s, got) and st_s.n_reads == 38_000
    ks2.free()
    eng2.close()
    for name in ("SS_CHUNK_BYTES", "SS_INGEST_THREADS", "SS_DGZ_GATE_SLACK"):
        monkeypatch.delenv(name)
    if shape != "level9_small_batches":
        monkeypatch.delenv("SS_DGZ_MAX_PIECES")
    # the host decoders give the same vector (SS_DGZ=0)
    monkeypatch.setenv("SS_DGZ", "0")
    got3, _ = eng.count_files(ks, [p1, p2])
    assert np.array_equal(got3, got)
    monkeypatch.setenv("SS_DGZ", "1")
    if shape == "members":                                                      # shards: every member decoded by exactly one
        monkeypatch.setenv("SS_GZ_SPLIT", "2")
        monkeypatch.setenv("SS_GZ_PART_BYTES", str(256 << 10))
        for n_shards in (2, 5):
            acc = np.zeros_like(got, dtype=np.uint64)
            n_reads = 0
            for s in range(n_shards):
                part, stp = eng.count_files(ks, [p1, p2], s, n_shards)
                acc += part
                n_reads += stp.n_reads
            assert np.array_equal(acc, o.cnt) and n_reads == 38_000, n_shards
            # mixed: one shard on the device path, the others on the host threads -- the same partition of the reads
            monkeypatch.setenv("SS_DGZ", "0")
            acc2 = np.zeros_like(acc)
            for s in range(n_shards):
                if s == 1:
                    monkeypatch.setenv("SS_DGZ", "1")
                acc2 += eng.count_files(ks, [p1, p2], s, n_shards)[0]
                monkeypatch.setenv("SS_DGZ", "0")
            monkeypatch.setenv("SS_DGZ", "1")
            assert np.array_equal(acc2, o.cnt), n_shards
