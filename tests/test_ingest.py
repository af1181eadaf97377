"""Read-file ingest on the host (no GPU): the library's own gzip inflate (strainscan_b200/csrc/ss_inflate.cuh,
the same code the device BGZF kernel runs), record-aligned chunking, producer threads and sharding,
checked against Python's zlib/gzip -- the stand-in for the `zcat a b |` of library/identify.py:82 and
library/Vote_Strain_L2_Lasso_new_sp.py:359,367.  Calls go through the C ABI (ss_ingest_files_host)."""
import ctypes as C
import gzip
import io
import os
import zlib

import numpy as np
import pytest

from strainscan_b200 import _lib
from tests import util

CHUNK = 256 << 10      # smallest legal chunk: forces many chunks on ~MB inputs


def ingest(paths, shard=0, n_shards=1, chunk=CHUNK, threads=3):
    lib = _lib.load()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    total = sum(os.path.getsize(p) for p in paths if os.path.exists(p)) * 40 + (1 << 20)
    out = C.create_string_buffer(total)
    n, nch = C.c_size_t(), C.c_uint32()
    rc = lib.ss_ingest_files_host(arr, len(paths), shard, n_shards, chunk, threads, out, total, C.byref(n), C.byref(nch))
    if rc:
        raise _lib.StrainScanB200Error(rc, lib.ss_last_error().decode())
    return out.raw[:n.value], nch.value


def records(text):
    lines = text.split(b"\n")
    assert lines[-1] == b"", "chunks must end with a newline"
    lines = lines[:-1]
    assert len(lines) % 4 == 0
    return sorted(tuple(lines[i:i + 4]) for i in range(0, len(lines), 4))


@pytest.fixture(scope="module")
def fastq():
    rng = np.random.default_rng(7)
    g = util.rand_genome(rng, 50_000)
    return util.make_reads(rng, g, 6000, read_len=150, var_len=True)      # ~2 MB, '@'/'+' qualities included


def _variants(fq):
    v = {}
    for lvl in (1, 6, 9):
        v["level%d" % lvl] = gzip.compress(fq, lvl)
    v["stored"] = gzip.compress(fq, 0)
    v["multi_member"] = b"".join(gzip.compress(fq[i:i + 70_001], 6) for i in range(0, len(fq), 70_001))
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
    v["fixed_huffman"] = co.compress(fq) + co.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_HUFFMAN_ONLY)
    v["huffman_only"] = co.compress(fq) + co.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 1)
    v["memlevel1"] = co.compress(fq) + co.flush()              # a new dynamic block every ~100 symbols
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(fq), 9973):
        parts += [co.compress(fq[i:i + 9973]), co.flush(zlib.Z_SYNC_FLUSH)]
    v["sync_flush"] = b"".join(parts) + co.flush()
    b = io.BytesIO()
    with gzip.GzipFile(filename="reads_R1.fastq", mode="wb", fileobj=b, mtime=1) as gz:
        gz.write(fq)
    v["fname_header"] = b.getvalue()
    v["trailing_garbage"] = gzip.compress(fq) + b"\0\0\0not a member"
    return v


def test_gz_stream_matches_zlib_bytewise(fastq, tmp_path):
    """One gz file, one producer: chunks arrive in stream order, so the bytes must equal zlib's output."""
    for name, data in _variants(fastq).items():
        p = str(tmp_path / ("%s.fq.gz" % name))
        open(p, "wb").write(data)
        got, n_chunks = ingest([p], threads=1)
        assert got == fastq, name
        assert n_chunks >= len(fastq) // CHUNK, name


def test_gz_by_magic_and_plain_by_content(fastq, tmp_path):
    p1, p2 = str(tmp_path / "a.fq"), str(tmp_path / "b.dat")
    open(p1, "wb").write(fastq)
    open(p2, "wb").write(gzip.compress(fastq))            # gzip without the .gz suffix: detected by its magic
    assert records(ingest([p1])[0]) == records(fastq)
    assert records(ingest([p2])[0]) == records(fastq)


def test_paired_files_mixed_and_tail_normalisation(fastq, tmp_path):
    half = fastq.index(b"\n@", len(fastq) // 2) + 1
    while not fastq[half:].split(b"\n", 3)[2].startswith(b"+"):     # a real record start ('@' may open a quality line)
        half = fastq.index(b"\n@", half) + 1
    a, b = fastq[:half], fastq[half:]
    p1, p2 = str(tmp_path / "r1.fq.gz"), str(tmp_path / "r2.fq")
    open(p1, "wb").write(gzip.compress(a[:-1]))                       # no final newline
    open(p2, "wb").write(b + b"\n\n  \n")                             # trailing blank lines
    got, _ = ingest([p1, p2], threads=4)
    assert records(got) == records(fastq)


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_shards_partition_the_reads(fastq, tmp_path, gz, n_shards):
    p = str(tmp_path / ("s.fq.gz" if gz else "s.fq"))
    open(p, "wb").write(gzip.compress(fastq) if gz else fastq)
    parts = [ingest([p], s, n_shards)[0] for s in range(n_shards)]
    assert sum(len(x) for x in parts) == len(fastq)
    assert records(b"".join(parts)) == records(fastq)
    assert sum(1 for x in parts if x) >= 2          # the work really is spread


def test_large_chunk_single_delivery(fastq, tmp_path):
    p = str(tmp_path / "one.fq.gz")
    open(p, "wb").write(gzip.compress(fastq))
    got, n = ingest([p], chunk=32 << 20)
    assert got == fastq and n == 1


def test_empty_inputs(tmp_path):
    p1, p2, p3 = str(tmp_path / "e.fq"), str(tmp_path / "e.fq.gz"), str(tmp_path / "blank.fq")
    open(p1, "wb").close()
    open(p2, "wb").write(gzip.compress(b""))
    open(p3, "wb").write(b"\n\n")
    assert ingest([p1, p2, p3]) == (b"", 0)


def test_errors_are_loud(fastq, tmp_path):
    gz = gzip.compress(fastq)
    bad = {
        "truncated.fq.gz": gz[:len(gz) // 2],
        "no_trailer.fq.gz": gz[:-8],
        "not_gzip.fq.gz": fastq,
        "junk.fq": b"hello\n",
        "bad_wrapped.fq": b"@r1\nACGT\nACGT\n+\nIIII\n",
    }
    corrupt = bytearray(gz)
    corrupt[len(gz) // 3] ^= 0x55
    for name, data in bad.items():
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        with pytest.raises(_lib.StrainScanB200Error) as e:
            ingest([p])
        assert e.value.code in (_lib.SS_ERR_IO, _lib.SS_ERR_FORMAT), name
    p = str(tmp_path / "corrupt.fq.gz")
    open(p, "wb").write(bytes(corrupt))
    try:                                       # a flipped byte either fails to decode or decodes to a wrong length
        got, _ = ingest([p])
        assert got != fastq
    except _lib.StrainScanB200Error as e:
        assert e.code in (_lib.SS_ERR_IO, _lib.SS_ERR_FORMAT)
    with pytest.raises(_lib.StrainScanB200Error):
        ingest([str(tmp_path / "missing.fq")])


def test_random_deflate_streams_roundtrip(tmp_path):
    """Non-FASTQ payloads exercise long codes, long matches and stored blocks: wrap each payload as the
    sequence line of a single record so the chunker accepts it."""
    rng = np.random.default_rng(11)
    payloads = [
        bytes(rng.integers(65, 91, 300_000, dtype=np.uint8)),                                # incompressible letters
        b"ACGT" * 100_000,                                                                    # distance-4 matches, length 258
        b"A" * 400_000,                                                                       # run-length
        bytes(rng.choice(np.arange(65, 91, dtype=np.uint8), 400_000,
                         p=np.array([2.0 ** -i for i in range(1, 26)] + [2.0 ** -25]))),      # skewed: 15-bit codes
    ]
    for i, pl in enumerate(payloads):
        fq = b"@r\n" + pl + b"\n+\n" + pl + b"\n"
        for lvl in (1, 6, 9):
            p = str(tmp_path / ("p%d_%d.fq.gz" % (i, lvl)))
            open(p, "wb").write(gzip.compress(fq, lvl))
            got, _ = ingest([p], chunk=4 << 20, threads=1)
            assert got == fq, (i, lvl)


# ---- BGZF: the producer path that feeds the device inflate kernel (ss_gunzip.cu) ----------------------
# ss_ingest_files_host runs the same batching (member chain walk, host-decoded boundary members, record
# aligned batches) and executes the per-member inflate the kernel would run with the same source.
BGZF_CHUNK = 1 << 20


@pytest.mark.parametrize("block,level,empties", [(0xFF00, 6, 0), (10_000, 1, 0), (65536, 9, 0), (3000, 6, 3), (200, 6, 0)])
def test_bgzf_batches_roundtrip(fastq, tmp_path, monkeypatch, block, level, empties):
    import gzip as _gz
    data = util.bgzf_compress(fastq, block=block, level=level, empty_every=empties)
    assert _gz.decompress(data) == fastq                       # the writer itself is a valid multi-member gzip
    p = str(tmp_path / "b.fq.gz")
    open(p, "wb").write(data)
    monkeypatch.setenv("SS_BGZF_OUT_CAP", str(1 << 20))        # ~600 KB of device text per batch: several batches
    got, n_chunks = ingest([p], chunk=BGZF_CHUNK, threads=1)
    assert got == fastq
    assert n_chunks >= 3
    monkeypatch.setenv("SS_BGZF_GPU", "0")                     # same file through the host gzip stream
    assert ingest([p], chunk=BGZF_CHUNK, threads=1)[0] == fastq


def test_bgzf_sharded_paired_and_tails(fastq, tmp_path, monkeypatch):
    monkeypatch.setenv("SS_BGZF_OUT_CAP", str(1 << 20))
    half = len(fastq) // 2
    half = fastq.index(b"\n@", half) + 1
    while not fastq[half:].split(b"\n", 3)[2].startswith(b"+"):
        half = fastq.index(b"\n@", half) + 1
    p1, p2 = str(tmp_path / "r1.fq.gz"), str(tmp_path / "r2.fq.gz")
    open(p1, "wb").write(util.bgzf_compress(fastq[:half - 1], eof_marker=False))      # no final newline, no EOF marker
    open(p2, "wb").write(util.bgzf_compress(fastq[half:] + b"\n \n"))                  # blank tail lines
    assert records(ingest([p1, p2], chunk=BGZF_CHUNK)[0]) == records(fastq)
    for n in (2, 3):
        parts = [ingest([p1, p2], s, n, chunk=BGZF_CHUNK)[0] for s in range(n)]
        assert records(b"".join(parts)) == records(fastq)
        assert all(parts)


def test_bgzf_small_and_broken(tmp_path, monkeypatch):
    tiny = b"@r\nACGTACGT\n+\nIIIIIIII\n"
    p = str(tmp_path / "tiny.fq.gz")
    open(p, "wb").write(util.bgzf_compress(tiny))
    assert ingest([p], chunk=BGZF_CHUNK) == (tiny, 1)
    open(p, "wb").write(util.bgzf_compress(b""))
    assert ingest([p], chunk=BGZF_CHUNK) == (b"", 0)
    rng = np.random.default_rng(3)
    fq = util.make_reads(rng, util.rand_genome(rng, 20_000), 3000, 150)
    good = util.bgzf_compress(fq, block=20_000)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x10                                 # inside some member's deflate data
    open(p, "wb").write(bytes(bad))
    try:
        got, _ = ingest([p], chunk=BGZF_CHUNK)
        assert got != fq
    except _lib.StrainScanB200Error as e:
        assert e.code in (_lib.SS_ERR_IO, _lib.SS_ERR_FORMAT)
    open(p, "wb").write(good[:len(good) // 2])                 # chain cut inside a member
    with pytest.raises(_lib.StrainScanB200Error):
        ingest([p], chunk=BGZF_CHUNK)
    open(p, "wb").write(util.bgzf_compress(b">r1\nACGT\n"))            # FASTA inside BGZF: rewritten on the host
    assert ingest([p], chunk=BGZF_CHUNK)[0] == b"@r1\nACGT\n+\nIIII\n"


@pytest.mark.parametrize("block", [0xFF00, 5000, 300])
def test_bgzf_file_parts_agree_on_record_boundaries(tmp_path, monkeypatch, block):
    """A BGZF file is cut into parts at member starts (one producer each); neighbouring parts split the
    boundary member's text at the same record start, so every read is delivered exactly once."""
    rng = np.random.default_rng(21)
    fq = util.make_reads(rng, util.rand_genome(rng, 30_000), 16_000, 150, var_len=True)     # ~5 MB
    monkeypatch.setenv("SS_BGZF_PART_BYTES", str(256 << 10))
    monkeypatch.setenv("SS_BGZF_OUT_CAP", str(1 << 20))
    p = str(tmp_path / "parts.fq.gz")
    open(p, "wb").write(util.bgzf_compress(fq, block=block, level=1))
    for threads in (1, 4, 7):
        got, n_chunks = ingest([p], chunk=BGZF_CHUNK, threads=threads)
        assert len(got) == len(fq) and records(got) == records(fq), (block, threads)
    for n in (2, 3):
        parts = [ingest([p], s, n, chunk=BGZF_CHUNK, threads=4)[0] for s in range(n)]
        assert records(b"".join(parts)) == records(fq)
    # one very long record spanning many members: no record start inside the boundary members
    long_fq = b"@r0\n" + b"ACGT" * 100_000 + b"\n+\n" + b"I" * 400_000 + b"\n" + fq[:200_000]
    long_fq = long_fq[:long_fq.rindex(b"\n@") + 1]
    while not long_fq.endswith(b"\n") or len(long_fq.split(b"\n")) % 4 != 1:
        long_fq = long_fq[:long_fq.rindex(b"\n@") + 1]
    open(p, "wb").write(util.bgzf_compress(long_fq, block=0xFF00))
    try:
        got, _ = ingest([p], chunk=BGZF_CHUNK, threads=4)
        assert records(got) == records(long_fq)
    except _lib.StrainScanB200Error as e:          # a loud, documented limit (like the 64 KiB rule of the plain chunker)
        assert e.code == _lib.SS_ERR_FORMAT and "record boundary" in str(e)


def test_randomized_inputs_chunks_threads_shards(tmp_path, monkeypatch):
    """Seeded sweep over file encodings (plain / gzip / BGZF), record lengths, CRLF, file tails, chunk
    sizes, producer counts, BGZF part and batch sizes and shard counts: the delivered records must be the
    file's records, every byte accounted for (tail blank lines trimmed, last line closed)."""
    rng = np.random.default_rng(2026)
    G = util.rand_genome(rng, 40_000)
    for it in range(36):
        fq = util.make_reads(rng, G, int(rng.integers(1, 5000)), int(rng.integers(20, 300)), var_len=rng.random() < 0.7,
                             crlf=rng.random() < 0.2)
        tail = [b"", b"\n", b"\n\n", b" \n", b"\r\n\r\n"][int(rng.integers(0, 5))]
        data = (fq[:-1] if rng.random() < 0.3 else fq) + tail
        kind = ("plain", "gz", "bgzf")[it % 3]
        p = str(tmp_path / ("f%d.%s" % (it, "fq" if kind == "plain" else "fq.gz")))
        if kind == "plain":
            blob = data
        elif kind == "gz":
            blob = gzip.compress(data, int(rng.integers(1, 10)))
        else:
            blob = util.bgzf_compress(data, block=int(rng.integers(100, 65000)), level=int(rng.integers(1, 10)),
                                      eof_marker=rng.random() < 0.5, empty_every=int(rng.integers(0, 4)))
        open(p, "wb").write(blob)
        monkeypatch.setenv("SS_BGZF_OUT_CAP", str(int(rng.integers(1 << 20, 4 << 20))))
        monkeypatch.setenv("SS_BGZF_PART_BYTES", str(int(rng.integers(256 << 10, 2 << 20))))
        chunk = (256 << 10, 1 << 20, 3 << 20)[int(rng.integers(0, 3))]
        n_sh = int(rng.integers(1, 5))
        got = b"".join(ingest([p], s, n_sh, chunk=chunk, threads=int(rng.integers(1, 9)))[0] for s in range(n_sh))
        exp = data.rstrip(b"\r\n \t") + b"\n"
        assert len(got) == len(exp), (it, kind)
        assert records(got) == records(exp), (it, kind)


# ---- parallel decoding of ordinary (single-stream) gzip: ss_pgz.cuh ----------------------------------------
@pytest.mark.parametrize("threads,span", [(2, 64 << 10), (5, 100_000), (8, 256 << 10)])
def test_parallel_gzip_equals_zlib(tmp_path, monkeypatch, threads, span):
    """Pieces of one deflate stream decoded concurrently (block starts found by search, unknown 32 KiB windows
    carried as 16-bit marker symbols, pieces stitched only where the previous one ended exactly on the found
    start): the bytes must equal zlib's, for every stream shape."""
    rng = np.random.default_rng(31)
    fq = util.make_reads(rng, util.rand_genome(rng, 60_000), 16_000, 150, var_len=True)     # ~5 MB
    monkeypatch.setenv("SS_PGZ_THREADS", str(threads))
    monkeypatch.setenv("SS_PGZ_SPAN", str(span))
    for name, data in _variants(fq).items():
        p = str(tmp_path / ("%s.fq.gz" % name))
        open(p, "wb").write(data)
        got, _ = ingest([p], threads=1, chunk=1 << 20)
        assert got == fq, name
    # sharded: chunks of the (deterministic) stream are dealt round-robin
    p = str(tmp_path / "level6.fq.gz")
    parts = [ingest([p], s, 3, chunk=512 << 10)[0] for s in range(3)]
    assert records(b"".join(parts)) == records(fq)


def test_parallel_gzip_errors(tmp_path, monkeypatch):
    rng = np.random.default_rng(32)
    fq = util.make_reads(rng, util.rand_genome(rng, 60_000), 12_000, 150)
    gz = gzip.compress(fq, 6)
    monkeypatch.setenv("SS_PGZ_THREADS", "4")
    monkeypatch.setenv("SS_PGZ_SPAN", str(64 << 10))
    p = str(tmp_path / "x.fq.gz")
    for bad in (gz[:len(gz) // 2], gz[:-8], gz[:-3]):
        open(p, "wb").write(bad)
        with pytest.raises(_lib.StrainScanB200Error, match="inflate failed"):
            ingest([p])
    flipped = bytearray(gz)
    for i in range(7):
        flipped[len(gz) * (i + 1) // 9] ^= 0xA5
    open(p, "wb").write(bytes(flipped))
    try:
        got, _ = ingest([p])
        assert got != fq
    except _lib.StrainScanB200Error as e:
        assert e.code in (_lib.SS_ERR_IO, _lib.SS_ERR_FORMAT)
    # sequential and parallel modes agree
    open(p, "wb").write(gz)
    a = ingest([p], threads=1)[0]
    monkeypatch.setenv("SS_PGZ_THREADS", "1")
    assert ingest([p], threads=1)[0] == a == fq


def test_fasta_and_wrapped_fastq_are_rewritten(tmp_path):
    """Dialects Jellyfish accepts beyond 4-line FASTQ (csrc/ss_fastx.h): the sequence characters arrive in the
    same order on one line per record, whatever the file encoding."""
    rng = np.random.default_rng(41)
    seqs = [util.rand_genome(rng, int(n)) for n in rng.integers(1, 400, 300)]
    fasta = b"".join(b">s%d some text\n" % i + b"\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + b"\n" for i, s in enumerate(seqs))
    wrapped = b"".join(b"@s%d\n" % i + b"\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n+\n" +
                       b"\n".join((b"@" * len(s))[j:j + 50] for j in range(0, len(s), 50)) + b"\n" for i, s in enumerate(seqs))
    want_fa = b"".join(b"@s%d some text\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs))
    want_fq = b"".join(b"@s%d\n%s\n+\n%s\n" % (i, s, b"@" * len(s)) for i, s in enumerate(seqs))
    for name, blob, want in (("a.fa", fasta, want_fa), ("a.fa.gz", gzip.compress(fasta), want_fa),
                             ("w.fq", b"\n\n" + wrapped, want_fq), ("w.fq.gz", util.bgzf_compress(wrapped), want_fq),
                             ("crlf.fa", fasta.replace(b"\n", b"\r\n"), None)):
        p = str(tmp_path / name)
        open(p, "wb").write(blob)
        got, _ = ingest([p], chunk=BGZF_CHUNK)
        if want is not None:
            assert got == want, name
        else:                                   # CR characters stay inside the joined sequence (they break windows there too)
            assert got.count(b"\n") == 4 * len(seqs) and got.startswith(b"@s0 some text\r\n")
        parts = [ingest([p], s, 2, chunk=256 << 10)[0] for s in range(2)]
        assert sorted(b"".join(parts).split(b"\n")) == sorted(got.split(b"\n"))


def test_long_records_cut_with_a_widening_window(tmp_path):
    """Long-read FASTQ: records of 6 kb .. 240 kb.  The chunker looks for a record start in the last 64 KiB of a chunk
    first and widens the window 4x at a time; only a record longer than the chunk itself is refused (loudly)."""
    rng = np.random.default_rng(5)
    G = util.rand_genome(rng, 300_000)
    fq = b"".join(util.make_reads(rng, G, n, L) for n, L in ((20, 3000), (5, 40_000), (2, 120_000), (20, 150)))
    p = tmp_path / "long.fq"
    p.write_bytes(fq)
    pg = tmp_path / "long.fq.gz"
    pg.write_bytes(gzip.compress(fq, 4))
    for path in (p, pg):
        for chunk in (1 << 19, 1 << 20):
            for threads in (1, 4):
                text, _ = ingest([str(path)], chunk=chunk, threads=threads)
                assert sorted(records(text)) == sorted(records(fq)), (path.name, chunk, threads)
    with pytest.raises(_lib.StrainScanB200Error) as ei:
        ingest([str(p)], chunk=1 << 18, threads=2)
    assert ei.value.code == 4 and "record boundary" in str(ei.value)


def _members(fq, sizes, level=6):
    out, i = [], 0
    for n in sizes:
        out.append(gzip.compress(fq[i:i + n], level))
        i += n
    assert i >= len(fq)
    return b"".join(out)


@pytest.mark.parametrize("aligned", [True, False])
@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_multi_member_gzip_is_split_at_member_starts(tmp_path, monkeypatch, aligned, n_shards):
    """A file of many gzip members (what `cat lane*.fq.gz` or a block-wise compressor writes) is cut into fixed parts and
    every member is decoded by exactly one shard; members need not hold whole records (a record that straddles a part
    boundary goes to the earlier part, as for BGZF).  SS_GZ_SPLIT=2 forces the mode on this small file."""
    rng = np.random.default_rng(31 + n_shards)
    g = util.rand_genome(rng, 60_000)
    fq = util.make_reads(rng, g, 24_000, read_len=150, var_len=True)          # ~7 MB
    if aligned:                                                                 # members = whole records
        starts = [m.start() for m in __import__("re").finditer(rb"(?m)^@read", fq)]
        cuts = [starts[i] for i in range(0, len(starts), 700)] + [len(fq)]
        sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
    else:                                                                       # arbitrary byte boundaries, some tiny members
        sizes, left = [], len(fq)
        while left > 0:
            n = int(rng.choice([17, 300, 40_000, 130_001, 260_000]))
            sizes.append(min(n, left)); left -= sizes[-1]
    p = str(tmp_path / "mm.fq.gz")
    open(p, "wb").write(_members(fq, sizes, level=int(rng.choice([1, 6]))))
    monkeypatch.setenv("SS_GZ_SPLIT", "2")
    monkeypatch.setenv("SS_GZ_PART_BYTES", str(256 << 10))                      # ~12 parts
    for threads in (1, 4):                                                      # the cut does not depend on the thread count
        parts = [ingest([p], s, n_shards, threads=threads)[0] for s in range(n_shards)]
        assert sum(len(x) for x in parts) == len(fq)
        assert records(b"".join(parts)) == records(fq)
        if n_shards in (2, 3):
            assert sum(1 for x in parts if x) == n_shards
        if threads == 1:
            first = parts
        else:
            assert [records(x) for x in parts] == [records(x) for x in first]
    monkeypatch.setenv("SS_GZ_SPLIT", "0")                                      # the whole-stream decoder gives the same reads
    assert records(b"".join(ingest([p], s, n_shards)[0] for s in range(n_shards))) == records(fq)


def test_multi_member_gzip_odd_shapes(tmp_path, monkeypatch):
    """Empty members, trailing garbage, a missing final newline, a lone huge member among small ones, and a header-like
    byte pattern inside the text of a stored block."""
    rng = np.random.default_rng(5)
    g = util.rand_genome(rng, 30_000)
    fq = util.make_reads(rng, g, 9000, read_len=150)
    monkeypatch.setenv("SS_GZ_SPLIT", "2")
    monkeypatch.setenv("SS_GZ_PART_BYTES", str(128 << 10))
    n = len(fq)
    third = fq.index(b"\n@read", n // 3) + 1
    blobs = {
        "empties": gzip.compress(fq[:third]) + gzip.compress(b"") + gzip.compress(b"") + gzip.compress(fq[third:]),
        "garbage": gzip.compress(fq[:third]) + gzip.compress(fq[third:]) + b"\0\0\0\0not a member at all",
        "no_final_newline": gzip.compress(fq[:third]) + gzip.compress(fq[third:-1]),
        "one_big_member": gzip.compress(fq[:2000]) + gzip.compress(fq[2000:n - 3000], 1) + gzip.compress(fq[n - 3000:]),
        # stored blocks keep the text verbatim: a gzip magic inside it must not be taken for a member
        "magic_in_stored": gzip.compress(fq[:third], 0) + gzip.compress(fq[third:], 0),
    }
    want = records(fq)
    for name, blob in blobs.items():
        p = str(tmp_path / (name + ".fq.gz"))
        open(p, "wb").write(blob)
        for n_shards in (1, 3):
            got = b"".join(ingest([p], s, n_shards)[0] for s in range(n_shards))
            assert records(got) == want, name
    bad = gzip.compress(fq[:third]) + b"junk junk junk junk junk" + gzip.compress(fq[third:])
    p = str(tmp_path / "junk_between.fq.gz")
    open(p, "wb").write(bad)
    with pytest.raises(_lib.StrainScanB200Error):
        for s in range(3):
            ingest([p], s, 3)
    # Header look-alikes inside the data.  (1) the gzip magic + a fixed-Huffman "block" of text bytes, in a quality line
    # of a STORED member: the trial decode of the member finder rejects it, the file splits as usual.  (2) a complete,
    # valid empty gzip member spelled out inside stored text: no finder can tell it from a real member, so the split
    # mode must fail loudly (and SS_GZ_SPLIT=0 must still read the file like zcat).
    lines = fq.split(b"\n")
    q = len(lines) // 2 // 4 * 4 + 3
    fake = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + b"\x4b\x4c\x4a" * 40
    look = lines[:]
    look[q] = (fake + look[q])[:len(look[q])] if len(look[q]) > 60 else look[q]
    look[q - 2] = b"N" * len(look[q])
    text1 = b"\n".join(look)
    half = text1.index(b"\n@read", len(text1) // 5) + 1
    p = str(tmp_path / "lookalike.fq.gz")
    open(p, "wb").write(gzip.compress(text1[:half], 0) + gzip.compress(text1[half:], 0))
    for n_shards in (1, 4):
        assert records(b"".join(ingest([p], s, n_shards)[0] for s in range(n_shards))) == records(text1)
    empty_member = gzip.compress(b"", mtime=0)
    evil = lines[:]
    evil[q] = empty_member + b"I" * 200
    evil[q - 2] = b"A" * len(evil[q])
    text2 = b"\n".join(evil)
    p = str(tmp_path / "evil.fq.gz")
    open(p, "wb").write(gzip.compress(text2[:half], 0) + gzip.compress(text2[half:], 0))
    with pytest.raises(_lib.StrainScanB200Error) as ei:
        for s in range(4):
            ingest([p], s, 4)
    assert "SS_GZ_SPLIT=0" in str(ei.value)
    monkeypatch.setenv("SS_GZ_SPLIT", "0")
    assert records(ingest([p])[0]) == records(text2)
