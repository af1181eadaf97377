"""The device inflate of ORDINARY gzip streams (strainscan_b200/csrc/ss_dgz.cuh: block starts found per piece, marker-symbol
decode with an unknown window, exact stitching, window chain, resolve) run on the CPU with the same source
(ss_dgz_inflate_host) against Python's zlib/gzip -- the stand-in for the `zcat` of library/identify.py:82.  The GPU form
of the same pipeline is checked in tests/test_gpu_parity.py."""
import ctypes as C
import gzip
import io
import zlib

import numpy as np
import pytest

from strainscan_b200 import _lib
from tests import util


def dgz(blob, first=0, stop=None, max_pieces=64, piece=8192, sym_per_byte=64, cap=None):
    lib = _lib.load()
    cap = cap or (len(blob) * 40 + (1 << 20))
    out = C.create_string_buffer(cap)
    n, stopped = C.c_size_t(), C.c_size_t()
    stats = (C.c_uint64 * 4)()
    rc = lib.ss_dgz_inflate_host(blob, len(blob), first, len(blob) if stop is None else stop, max_pieces, piece, sym_per_byte,
                                 out, cap, C.byref(n), C.byref(stopped), stats)
    if rc:
        raise _lib.StrainScanB200Error(rc, lib.ss_last_error().decode())
    return out.raw[:n.value], stopped.value, list(stats)


@pytest.fixture(autouse=True, params=["one_decoder_per_warp", "lanes", "lanes_ring_short_rounds"])
def decoder(request, monkeypatch):
    """K8 exists twice: dgz_decode_piece (ss_dgz.cuh) and the decoder of ss_dgz2.cuh that several lanes of a warp run
    (16-bit tables, rounds, a ring of recent symbols); every test runs on both, the second one also with the ring and
    rounds of 3 iterations."""
    monkeypatch.setenv("SS_DGZ_LANES", "0" if request.param == "one_decoder_per_warp" else "4")
    monkeypatch.setenv("SS_DGZ_RING", "0")
    if request.param == "lanes_ring_short_rounds":       # + the ring of recent symbols the match copies read from
        monkeypatch.setenv("SS_DGZ_ROUND", "3")
        monkeypatch.setenv("SS_DGZ_RING", "1")
    return request.param


@pytest.fixture(scope="module")
def fastq():
    rng = np.random.default_rng(11)
    g = util.rand_genome(rng, 80_000)
    return util.make_reads(rng, g, 9000, read_len=150, var_len=True)      # ~3 MB


def test_single_stream_levels_and_piece_sizes(fastq):
    for level in (1, 6, 9):
        blob = gzip.compress(fastq, level)
        for piece, max_pieces in ((4096, 1000), (12544, 37), (16384, 7), (65536, 3), (1 << 20, 2)):      # (any multiple of 256 may be planned)
            text, stopped, st = dgz(blob, max_pieces=max_pieces, piece=piece)
            assert text == fastq, (level, piece)
            assert stopped == len(blob) and st[3] == 1
        # most pieces really start in the middle of the stream (found block starts), and their text went through markers
        _, _, st = dgz(blob, max_pieces=1000, piece=16384)
        assert st[1] > 5 and st[1] <= st[0] + st[2]


def test_stream_shapes(fastq):
    shapes = {}
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
    shapes["fixed_huffman"] = co.compress(fastq) + co.flush()               # no dynamic header to find: one piece runs through
    shapes["stored"] = gzip.compress(fastq, 0)
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 1)
    shapes["memlevel1"] = co.compress(fastq) + co.flush()                   # a new block every ~100 symbols
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(fastq), 9973):
        parts += [co.compress(fastq[i:i + 9973]), co.flush(zlib.Z_SYNC_FLUSH)]
    shapes["sync_flush"] = b"".join(parts) + co.flush()
    for name, strategy in (("huffman_only", zlib.Z_HUFFMAN_ONLY), ("rle", zlib.Z_RLE), ("filtered", zlib.Z_FILTERED)):
        co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strategy)            # no distance code at all / a single distance code
        shapes[name] = co.compress(fastq) + co.flush()
    b = io.BytesIO()
    with gzip.GzipFile(filename="reads_R1.fastq", mode="wb", fileobj=b, mtime=1) as gz:
        gz.write(fastq)
    shapes["fname_header"] = b.getvalue()
    shapes["trailing_garbage"] = gzip.compress(fastq) + b"\0\0\0not a member"
    shapes["multi_member"] = b"".join(gzip.compress(fastq[i:i + 70_001], 6) for i in range(0, len(fastq), 70_001))
    shapes["empty_members"] = gzip.compress(b"") + gzip.compress(fastq[:50_000]) + gzip.compress(b"") + gzip.compress(fastq[50_000:])
    low = b"A" * 300_000 + fastq[:10_000] + b"\n".join([b"ACGT" * 30] * 4000)      # very high compression ratio
    for name, blob in shapes.items():
        text, _, st = dgz(blob, max_pieces=50, piece=8192, sym_per_byte=512 if name in ("fixed_huffman", "stored") else 64)
        assert text == fastq, name
    text, _, st = dgz(gzip.compress(low, 9), max_pieces=16, piece=4096, sym_per_byte=256)
    assert text == low
    # a block that inflates to more than a piece's whole buffer is refused loudly, never mis-decoded
    with pytest.raises(_lib.StrainScanB200Error):
        dgz(gzip.compress(b"A" * 3_000_000, 9), max_pieces=4, piece=4096, sym_per_byte=2)


def test_member_ranges(fastq):
    """Member-split files: a decoder owns the members that start in [first, stop) and says where it stopped."""
    chunks = [fastq[i:i + 200_003] for i in range(0, len(fastq), 200_003)]
    blobs = [gzip.compress(c, 6) for c in chunks]
    starts = np.cumsum([0] + [len(b) for b in blobs]).tolist()
    blob = b"".join(blobs)
    text, stopped, st = dgz(blob, first=starts[2], stop=starts[5], piece=4096, max_pieces=33)
    assert text == b"".join(chunks[2:5]) and stopped == starts[5] and st[3] == 3
    text, stopped, _ = dgz(blob, first=starts[1], stop=starts[1] + 1, piece=4096, max_pieces=9)
    assert text == chunks[1] and stopped == starts[2]
    text, stopped, _ = dgz(blob, first=starts[-2], piece=4096)
    assert text == chunks[-1] and stopped == len(blob)


def test_corrupt_streams_fail(fastq):
    blob = bytearray(gzip.compress(fastq, 6))
    for at in (len(blob) // 3, len(blob) // 2, len(blob) - 20):
        bad = bytearray(blob)
        bad[at] ^= 0x5A
        try:
            text, _, _ = dgz(bytes(bad), piece=8192)
        except _lib.StrainScanB200Error:
            continue
        # a flipped bit that still decodes (literal changed) must at least not be silently "repaired"
        assert text != fastq
    with pytest.raises(_lib.StrainScanB200Error):
        dgz(bytes(blob[:len(blob) // 2]), piece=8192)                        # truncated
    with pytest.raises(_lib.StrainScanB200Error):
        dgz(b"hello, not gzip at all" * 100)


def test_random_binary_streams_roundtrip():
    rng = np.random.default_rng(3)
    for trial in range(12):
        n = int(rng.integers(1, 400_000))
        kind = trial % 3
        if kind == 0:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()                      # incompressible
        elif kind == 1:
            data = bytes(rng.integers(0, 4, n, dtype=np.uint8) + 65)                      # 2 bits per byte
        else:
            words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(50)]
            data = b" ".join(words[int(i)] for i in rng.integers(0, 50, n // 6 + 1))      # long matches
        level = int(rng.choice([1, 4, 6, 9]))
        blob = gzip.compress(data, level)
        text, _, _ = dgz(blob, piece=int(rng.choice([4096, 8192, 32768])), max_pieces=int(rng.integers(2, 40)), sym_per_byte=64)
        assert text == data, (trial, level)


def test_both_decoders_tell_the_same_story(fastq, monkeypatch):
    """Pieces found / used, batches and members are properties of the stream and the piece size, not of the decoder."""
    blobs = [gzip.compress(fastq, 1), gzip.compress(fastq, 9), b"".join(gzip.compress(fastq[i:i + 300_007], 4) for i in range(0, len(fastq), 300_007))]
    for blob in blobs:
        seen = []
        for lanes in ("0", "4"):
            monkeypatch.setenv("SS_DGZ_LANES", lanes)
            text, stopped, st = dgz(blob, max_pieces=40, piece=8192, sym_per_byte=16)
            assert text == fastq
            seen.append((stopped, st))
        assert seen[0] == seen[1]


def test_match_distances_around_the_ring():
    """Matches whose source lies exactly at, just inside and just outside the 1024-symbol ring of the lane decoder,
    run-length matches (distance 1-3, overlapping copies) and copies that wrap around the ring's end."""
    rng = np.random.default_rng(5)
    parts = []
    for period in (1, 2, 3, 4, 5, 1020, 1021, 1022, 1023, 1024, 1025, 1026, 1027, 2048, 5000, 31000):
        block = bytes(rng.integers(33, 127, period, dtype=np.uint8))
        reps = max(3, 2600 // period)
        parts.append(block * reps)                                   # every repetition is a match at distance `period`
        parts.append(bytes(rng.integers(33, 127, int(rng.integers(1, 40)), dtype=np.uint8)))     # shifts the ring phase
    data = b"".join(parts) * 3
    for level in (1, 6, 9):
        blob = gzip.compress(data, level)
        for piece, spb, max_pieces in ((4096, 256, 16), (65536, 64, 8)):
            text, _, _ = dgz(blob, piece=piece, max_pieces=max_pieces, sym_per_byte=spb)
            assert text == data, (level, piece)


def test_plan_of_pieces_and_buffers(monkeypatch):
    """ss_dgz_make_plan: whole waves of decoders, a wave's text inside the batch buffer, symbol room that follows the
    input's compression ratio (a piece out of room ends its batch), environment overrides."""
    lib = _lib.load()
    for name in ("SS_DGZ_LANES", "SS_DGZ_WARPS", "SS_DGZ_BATCH_MB", "SS_DGZ_PIECE_BYTES", "SS_DGZ_MAX_PIECES", "SS_DGZ_SYM_PER_BYTE", "SS_DGZ_RING"):
        monkeypatch.delenv(name, raising=False)

    def plan(n, ratio, n_sm=148):
        out = (C.c_uint64 * 6)()
        assert lib.ss_dgz_plan_host(n, n_sm, ratio, out) == 0
        return dict(zip(("piece", "max_pieces", "expand", "scratch", "device_bytes", "wave"), [int(x) for x in out]))

    for lanes in ("0", "2", "4"):
        monkeypatch.setenv("SS_DGZ_LANES", lanes)
        for n, ratio in ((9_300_000_000, 1.7), (1_160_000_000, 1.7), (3_700_000_000, 4.5), (40_000_000, 3.0), (2_000_000_000, 9.0)):
            p = plan(n, ratio)
            assert p["wave"] == 148 * {"0": 24, "2": 48, "4": 92}[lanes]
            assert 16384 <= p["piece"] <= 128 << 10 and p["piece"] % 256 == 0
            assert p["max_pieces"] >= p["wave"]                                    # a whole wave per batch ...
            assert p["wave"] * p["piece"] * 1.3 * max(3.5, ratio) <= p["scratch"] * 1.001 or p["piece"] == 16384   # ... whose text fits
            assert p["expand"] >= max(5, int(2.5 * ratio))                         # room for well-compressed inputs
            n_pieces = -(-n // p["piece"])
            waves = -(-n_pieces // p["wave"])
            if n_pieces > p["wave"]:                                               # the last wave is (nearly) full
                assert n_pieces > (waves - 1) * p["wave"] + 0.9 * p["wave"], (lanes, n, ratio, p)
            assert p["device_bytes"] < 60 << 30
    monkeypatch.setenv("SS_DGZ_LANES", "2")
    # an eighth of a config-3 file: two full waves of ~80 KiB pieces, not one full and a quarter
    p = plan(1_160_000_000, 1.7)
    assert 2 * p["wave"] * p["piece"] >= 1_160_000_000 > 2 * p["wave"] * (p["piece"] - 256)
    monkeypatch.setenv("SS_DGZ_PIECE_BYTES", "20480")
    monkeypatch.setenv("SS_DGZ_MAX_PIECES", "17")
    monkeypatch.setenv("SS_DGZ_SYM_PER_BYTE", "33")
    monkeypatch.setenv("SS_DGZ_BATCH_MB", "64")
    p = plan(10 ** 9, 2.0)
    assert (p["piece"], p["max_pieces"], p["expand"], p["scratch"]) == (20480, 17, 33, 64 << 20)


def test_decode_tables_agree_on_random_codes():
    """The 16-bit decode tables of the lane decoder against the 32-bit tables of the host decoder on random prefix codes
    (shapes zlib never writes: 15-bit codes under many prefixes, two-symbol codes, single-code and empty distance codes)."""
    lib = _lib.load()
    n = C.c_uint64()
    for seed in (1, 2, 3):
        rc = lib.ss_dgz_tables_selftest_host(seed, 150, C.byref(n))
        assert rc == 0, lib.ss_last_error().decode()
        assert n.value > 150 * 32768                                  # most codes are valid and were compared in full
