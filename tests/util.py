"""Seeded random inputs shared by the CPU and GPU tests (NumPy only)."""
import numpy as np

_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp(b):
    return b.translate(_COMP)[::-1]


def rand_genome(rng, n):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes()


def make_db(rng, genome, k, n_pos, both_strands=True, header="1", lower_frac=0.0, junk=0):
    """k-mer FASTA text: n_pos positions of `genome`, optionally with the reverse complement as a
    separate record (Build_tree.py:101-109), shuffled; `junk` malformed records mixed in."""
    pos = rng.choice(len(genome) - k, n_pos, replace=False)
    recs = []
    for p in pos:
        s = genome[p:p + k]
        recs.append(s)
        if both_strands:
            recs.append(revcomp(s))
    for j in range(junk):
        s = bytearray(genome[j:j + k])
        kind = j % 4
        if kind == 0:
            s[j % k] = ord("N")
        elif kind == 1:
            s = s[:k - 3]
        elif kind == 2:
            s = s + b"AC"
        else:
            s = bytearray(recs[j % len(recs)])      # exact duplicate of an earlier record
        recs.append(bytes(s))
    order = rng.permutation(len(recs))
    recs = [recs[i] for i in order]
    if lower_frac > 0:
        for i in np.nonzero(rng.random(len(recs)) < lower_frac)[0]:
            recs[i] = recs[i].lower()
    out = []
    for i, r in enumerate(recs):
        out.append(b">" + (header.encode() if header else str(i + 1).encode()) + b"\n" + r + b"\n")
    return b"".join(out)


def make_reads(rng, genome, n, read_len=150, p_sub=0.01, p_n=0.002, var_len=False, lower_frac=0.0,
               crlf=False, qual_noise=True, offtarget=0.0):
    """4-line FASTQ text.  Quality strings use the full Phred+33 range so lines may open with '@'/'+'."""
    nl = b"\r\n" if crlf else b"\n"
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(n):
        L = int(rng.integers(1, read_len * 2)) if var_len else read_len
        L = min(L, len(genome) - 1)
        if rng.random() < offtarget:
            r = bytearray(acgt[rng.integers(0, 4, L)].tobytes())
        else:
            s = int(rng.integers(0, len(genome) - L))
            r = bytearray(genome[s:s + L])
        for j in np.nonzero(rng.random(L) < p_sub)[0]:
            r[j] = b"ACGT"[int(rng.integers(0, 4))]
        for j in np.nonzero(rng.random(L) < p_n)[0]:
            r[j] = b"NRYKM"[int(rng.integers(0, 5))]
        r = bytes(r)
        if rng.random() < 0.5:
            r = revcomp(r)
        if rng.random() < lower_frac:
            r = r.lower()
        if qual_noise:
            q = (rng.integers(0, 42, L) + 33).astype(np.uint8).tobytes()
        else:
            q = b"I" * L
        out.append(b"@read%d some comment" % i + nl + r + nl + b"+" + nl + q + nl)
    return b"".join(out)


def bgzf_compress(data, block=0xFF00, level=6, eof_marker=True, empty_every=0):
    """Blocked gzip (BGZF, SAM spec 4.1): every member holds <= 64 KiB of text and its own total size in a
    'BC' extra subfield, so members inflate independently.  Written from the spec (bgzip is not in the image)."""
    import struct
    import zlib

    def member(chunk):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(raw) + 8
        assert bsize <= 65536
        return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                + raw + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    out = []
    for n, i in enumerate(range(0, len(data), block)):
        out.append(member(data[i:i + block]))
        if empty_every and n % empty_every == empty_every - 1:
            out.append(member(b""))
    if eof_marker:
        out.append(member(b""))
    return b"".join(out)
