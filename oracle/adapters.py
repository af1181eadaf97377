"""oracle/adapters.py -- TEST INFRASTRUCTURE ONLY (never imported by strainscan_b200/).

CPU restatement of what StrainScan's Python does on either side of the Jellyfish
subprocess on the identification hot path, so that tests can diff the CUDA path's
dense vectors against the reference's dict/list results.  Each function cites the
reference lines it follows.  Integer work only; results must match bit for bit.

  count_dense()            liboracle.so (kmer_count_oracle.c): jellyfish count --if / dump -c
  py_count_tiny()          pure-Python dict restatement of the same, for tiny cases
  l1_match_results()       identify.py:90-101  -> {ordinal: count}, valid_kmers = keys
  l2_py_o()                Vote_Strain_L2_Lasso_new_sp.py:312-322,386-403 (remove_1)
  match_node()             identify.py:106-127 (+ identify_low_depth.py:86-102 guard)
  adjust_profile_gather()  identify.py:167-191 (gather part only)
  stat_cov_all()           identify_strains_L2_Enet_Pscan_new_sp.py:33-49
  remain_cov()             identify_strains_L2_Enet_Pscan_new_sp.py:94-108 (get_remainc)
  candidate_counts()       identify_strains_L2_Enet_Pscan_new_sp.py:121-134 (get_candidate_arr)

Parity pin: tests/golden/*.json were produced by the real jellyfish-linux binary
(tests/golden/make_golden.py); tests/test_oracle.py checks count_dense() and
py_count_tiny() against all of them.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile kmer_count_oracle.c -> oracle/liboracle.so with gcc (see oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "kmer_count_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-Wall", "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_fasta_records.restype = ctypes.c_uint64
        L.orc_fasta_records.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.orc_count.restype = ctypes.c_int
        L.orc_count.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t),
                                ctypes.c_int, ctypes.c_uint64,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.orc_count_windows.restype = ctypes.c_int
        L.orc_count_windows.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t),
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
        _LIB = L
    return _LIB


class DenseCounts:
    """Per-record view of one `jellyfish count --if F ; dump -c` run (A.3 rules 1-5)."""

    def __init__(self, cnt, in_set, is_last, raw_upper, header_id, n_distinct):
        self.cnt, self.in_set, self.is_last = cnt, in_set, is_last
        self.raw_upper, self.header_id, self.n_distinct = raw_upper, header_id, n_distinct


def count_dense(fasta_bytes, k, read_files):
    """read_files: list of bytes objects (decompressed file contents, argv order)."""
    L = _lib()
    n = int(L.orc_fasta_records(fasta_bytes, len(fasta_bytes)))
    cnt = np.zeros(n, dtype=np.uint64)
    in_set = np.zeros(n, dtype=np.uint8)
    is_last = np.zeros(n, dtype=np.uint8)
    raw_upper = np.zeros(n, dtype=np.uint8)
    hid = np.zeros(n, dtype=np.uint64)
    nd = ctypes.c_uint64(0)
    nf = len(read_files)
    bufs = (ctypes.c_char_p * max(nf, 1))(*read_files)
    lens = (ctypes.c_size_t * max(nf, 1))(*[len(b) for b in read_files])
    rc = L.orc_count(fasta_bytes, len(fasta_bytes), int(k), bufs, lens, nf, n,
                     cnt.ctypes.data, in_set.ctypes.data, is_last.ctypes.data,
                     raw_upper.ctypes.data, hid.ctypes.data, ctypes.byref(nd))
    if rc != 0:
        raise RuntimeError("oracle orc_count failed rc=%d" % rc)
    return DenseCounts(cnt, in_set, is_last, raw_upper, hid, int(nd.value))


def count_windows(read_files, k):
    L = _lib()
    nf = len(read_files)
    bufs = (ctypes.c_char_p * max(nf, 1))(*read_files)
    lens = (ctypes.c_size_t * max(nf, 1))(*[len(b) for b in read_files])
    out = ctypes.c_uint64(0)
    rc = L.orc_count_windows(bufs, lens, nf, int(k), ctypes.byref(out))
    if rc != 0:
        raise RuntimeError("oracle orc_count_windows failed rc=%d" % rc)
    return int(out.value)


# ---------------------------------------------------------------------------------------------
# pure-Python restatement (tiny cases): returns the dump as {KMER: count}
# ---------------------------------------------------------------------------------------------
def _py_sequences(text):
    """Jellyfish-style record walk over one decoded file; yields joined sequence strings."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    i = 0
    while i < len(lines) and lines[i].strip() == "":
        i += 1
    if i == len(lines):
        return
    typ = lines[i][0]
    while i < len(lines):
        if lines[i].strip() == "" and all(x.strip() == "" for x in lines[i:]):
            return
        assert lines[i][0] == typ, "malformed"
        i += 1
        seq = []
        if typ == ">":
            while i < len(lines) and not lines[i].startswith(">"):
                seq.append(lines[i]); i += 1
            yield "".join(seq)
        else:
            while i < len(lines) and not lines[i].startswith("+"):
                seq.append(lines[i]); i += 1
            s = "".join(seq)
            i += 1
            q = 0
            while q < len(s) and i < len(lines):
                q += len(lines[i]); i += 1
            assert q == len(s), "malformed"
            yield s


def _py_windows(seq, k):
    seq = seq.upper()
    run = 0
    for p, ch in enumerate(seq):
        run = run + 1 if ch in "ACGT" else 0
        if run >= k:
            yield seq[p - k + 1:p + 1]


def py_count_tiny(fasta_text, k, read_texts):
    dump = {}
    for rec in _py_sequences(fasta_text):
        for w in _py_windows(rec, k):
            dump.setdefault(w, 0)
    for t in read_texts:
        for s in _py_sequences(t):
            for w in _py_windows(s, k):
                if w in dump:
                    dump[w] += 1
    return dump


# ---------------------------------------------------------------------------------------------
# adapters (what the reference's Python builds from the dump)
# ---------------------------------------------------------------------------------------------
def l1_match_results_from_dump(fasta_text, dump):
    """identify.py:90-101 verbatim in spirit: dict by record ordinal, .upper(), last wins."""
    lines = fasta_text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    idx = {}
    for i in range(len(lines) // 2):
        idx[lines[2 * i + 1].rstrip().upper()] = i
    return {idx[kmer]: c for kmer, c in dump.items()}


def l1_match_results(d):
    """Same mapping from the dense oracle view: keys = valid ordinals (identify.py:410)."""
    valid = (d.in_set & d.is_last).astype(bool)
    idx = np.nonzero(valid)[0]
    return {int(i): int(d.cnt[i]) for i in idx}


def l2_py_o(d):
    """Vote_Strain_L2_Lasso_new_sp.py:386-403: rows in kid order, remove_1."""
    c = np.where(d.raw_upper.astype(bool), d.cnt, 0).astype(np.int64)
    c[c == 1] = 0
    if d.header_id.size and np.all(d.header_id > 0):
        order = np.argsort(d.header_id, kind="stable")     # sorted(kid_match.items(), key=kid)
        c = c[order]
    return c


def l2_py_o_from_dump(fasta_text, dump):
    lines = fasta_text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    kid = {}
    for i in range(len(lines) // 2):
        kid[lines[2 * i + 1].rstrip()] = int(lines[2 * i][1:])
    res = sorted(kid.items(), key=lambda t: t[1])
    py = []
    for r in res:
        if r[0] not in dump:
            py.append(0)
        elif dump[r[0]] == 1:
            py.append(0)
        else:
            py.append(dump[r[0]])
    return np.array(py, dtype=np.int64)


def del_outlier(profile):
    """identify.py:106-112."""
    cutoff = 100 * np.median(profile)
    out = list(profile)
    for v in profile:
        if v >= cutoff:
            out.remove(v)
    return out


def match_node(match_results, node_ordinals, valid_kmers, min_valid=0):
    """identify.py:115-127; min_valid=1000 gives identify_low_depth.py:86-102.
    Returns (length, sorted k_profile) -- the reference's list order is set order; consumers use
    only len/mean/median of it, so tests compare it sorted."""
    d = set(int(x) for x in node_ordinals)
    valid = valid_kmers & d
    if min_valid and len(valid) < min_valid:
        return 0, []
    prof = [match_results[k] for k in valid if match_results[k] > 0]
    if prof:
        prof = del_outlier(prof)
    return len(valid), sorted(prof)


def adjust_profile_gather(match_results, node_ordinals, delete_positions, valid_kmers):
    """identify.py:167-191 gather part: node list minus overlap positions; None if < 1000 remain."""
    d = [int(x) for x in node_ordinals]
    delete = set(d[p] for p in delete_positions)
    ds = set(d)
    if len(ds) - len(delete) < 1000:
        return None
    valid = valid_kmers & (ds - delete)
    prof = [match_results[k] for k in valid if match_results[k] > 0]
    if prof:
        prof = del_outlier(prof)
    return len(valid), sorted(prof)


def stat_cov_all(X_dense, y):
    """identify_strains_L2_Enet_Pscan_new_sp.py:33-49 per column: (valid_kmr, total_kmr)."""
    out = []
    for j in range(X_dense.shape[1]):
        ix = X_dense[:, j].astype(np.int64)
        total = int(np.count_nonzero(ix))
        ic = ix * y
        ic[ic == 1] = 0
        out.append((int(np.count_nonzero(ic)), total))
    return out


def remain_cov(used_kmer, X_dense, y):
    """get_remainc (identify_strains...:94-108) integer parts: per strain (check, all_k) over rows
    not yet used.  used_kmer: 0/1 per row; X_dense rows x strains."""
    out = []
    for j in range(X_dense.shape[1]):
        npx = 2 * used_kmer.astype(np.int64) + X_dense[:, j].astype(np.int64)
        npx[npx > 1] = 0
        all_k = int(npx.sum())
        t = npx * y
        t[t == 1] = 0
        t[t > 1] = 1
        out.append((int(t.sum()), all_k))
    return out


def candidate_counts(X_dense, y):
    """get_candidate_arr (identify_strains...:121-134): per strain #rows with X=1 and y>1."""
    out = []
    for j in range(X_dense.shape[1]):
        t = X_dense[:, j].astype(np.int64) * y
        t[t == 1] = 0
        t[t > 1] = 1
        out.append(int(t.sum()))
    return out


def _percentile_nearest(a, q):
    """np.percentile(a, q, interpolation='nearest') as the reference calls it (the keyword is `method` in NumPy >= 1.22)."""
    return np.percentile(a, q, method="nearest")


def optimize_dominat_y(X_dense, y):
    """identify_strains_L2_Enet_Pscan_new_sp.py:136-175, dense as the reference: per strain the sum of the y values of
    its rows that lie inside the 5th..95th percentile (nearest) of its non-zero values; the first maximum wins."""
    res = []
    for c in range(X_dense.shape[1]):
        ix = X_dense[:, c].astype(np.int64)
        da = ix * y
        da_noz = da[da != 0]
        if np.sum(da_noz) == 0 or len(da_noz) < 1:
            res.append(0)
        else:
            f25 = _percentile_nearest(da_noz, 5)
            f75 = _percentile_nearest(da_noz, 95)
            tem = np.copy(y)
            tem[tem < f25] = 0
            tem[tem > f75] = 0
            res.append(int(np.dot(ix, tem)))
    res = np.array(res)
    return int(np.where(res == np.max(res))[0][0]), res


def get_avg_depth(dominat, X_dense, y):
    """identify_strains...:109-119: mean of the strain's y values (1 -> 0, zeros dropped) inside their 25th..75th
    percentile (nearest)."""
    doarr = X_dense[:, dominat].astype(np.int64) * y
    doarr[doarr == 1] = 0
    noz = doarr[doarr != 0]
    f25 = _percentile_nearest(noz, 25)
    f75 = _percentile_nearest(noz, 75)
    noz = noz.copy()
    noz[noz < f25] = 0
    noz[noz > f75] = 0
    return np.mean(noz[noz != 0])


def unique_cluster_rows(om_dense, all_cls):
    """detect_strains (identify_strains...:183-197): ln[row] = 1 when the row's k-mer belongs to exactly one of the
    identified clusters `all_cls` (1-based cluster ids = columns + 1 of overlap_matrix.npz), else 0."""
    cols = [int(a) - 1 for a in all_cls]
    ln = np.sum(om_dense[:, cols].astype(np.int64), axis=1)
    ln[ln > 1] = 0
    return ln
