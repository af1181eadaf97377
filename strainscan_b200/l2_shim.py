"""Drop-in mirrors of the reference's L2 count block and per-strain hit statistics.

  count_cluster(input_fq, fq2, db_dir, ksize)   library/Vote_Strain_L2_Lasso_new_sp.py:348-403
        -> py_o, the int64 vector vote_strain_L2() hands to detect_strains(): rows in kid order
           (kid - 1 = record ordinal of C<id>/all_kmer.fasta), absent -> 0, count == 1 -> 0
           (remove_1, Vote_...:312-322).
  stat_cov / cal_cov_all(ix, iy)                library/identify_strains_L2_Enet_Pscan_new_sp.py:33-49
  get_candidate_arr(ix, iy)                     ...:121-134
  get_remainc(dominat, used_kmer, pXt_tem, py, strain_remainc)        ...:94-108
        integer reductions on the GPU over the CSC form of all_strains_re.npz (no X.A densification,
        identify_strains...:200-201); the ratios are formed on the host exactly as the reference does.
  optimize_dominat_y(ix, iy)                    ...:136-175
  get_avg_depth(dominat, pX, py)                ...:109-119
  unique_cluster_rows(omatrix, all_cls)         ...:183-197  (`ln`, so that py_u = py * ln)
        order statistics ('nearest' percentiles) over one strain's rows: host NumPy on the CSC columns,
        O(nnz) instead of O(rows x strains).

The multi-GPU form sums the raw dense vectors first and applies remove_1 afterwards
(strainscan_b200/dist.py) -- never before the sum.
"""
import os

import numpy as np

from . import identify_shim
from ._lib import SS_REC_RAW_UPPER


def count_cluster(input_fq, fq2, db_dir, ksize, engine=None, return_stats=False):
    """db_dir = <DB>/Kmer_Sets_L2/Kmer_Sets/C<id> (Vote_...:287).  fq2 == '' for single-end."""
    eng = engine or identify_shim.default_engine()
    kset = identify_shim.cached_kmerset(eng, os.path.join(db_dir, "all_kmer.fasta"), int(ksize))
    reads = identify_shim.cached_reads(eng, (input_fq, fq2))
    counts, st = eng.count(kset, reads)
    py_o = remove_1(counts, kset)
    return (py_o, st) if return_stats else py_o


def remove_1(counts, kset):
    """Vote_...:312-322 + :386-388 on the dense vector: raw string not a dumped key -> 0,
    count == 1 -> 0, rows sorted by kid."""
    c = np.where((kset.flags & SS_REC_RAW_UPPER) != 0, counts, 0).astype(np.int64)
    c[c == 1] = 0
    hid = kset.header_ids
    n = c.size
    if n and not np.array_equal(hid, np.arange(1, n + 1, dtype=np.uint64)):
        if hid.min() >= 1 and np.unique(hid).size == n:
            c = c[np.argsort(hid, kind="stable")]
    return c


class StrainMatrix:
    """CSC view of the 0/1 strain matrix (all_strains_re.npz, Recls_withR_new.py:110-112)."""

    def __init__(self, X):
        import scipy.sparse as sp
        X = sp.csc_matrix(X)
        X.eliminate_zeros()
        X.sort_indices()
        self.n_rows, self.n_strains = X.shape
        self.col_ptr = X.indptr.astype(np.uint64)
        self.rows = X.indices.astype(np.uint32)

    @classmethod
    def load(cls, npz_path):
        import scipy.sparse as sp
        return cls(sp.load_npz(npz_path))


def strain_hits(X, y, row_mask=None, engine=None):
    """Per strain: (total, covered, sum) = (#rows X=1, #rows X=1 and y>1, sum of those y); rows limited
    to row_mask != 0 when given."""
    eng = engine or identify_shim.default_engine()
    if not isinstance(X, StrainMatrix):
        X = StrainMatrix(X)
    return eng.strain_reduce(X.col_ptr, X.rows, np.asarray(y, dtype=np.int64), row_mask)


def stat_cov_all(X, y, engine=None):
    """[[cov, valid_kmr, total_kmr], ...] per strain = stat_cov (identify_strains...:33-43)."""
    total, covered, _ = strain_hits(X, y, engine=engine)
    out = []
    for t, c in zip(total.tolist(), covered.tolist()):
        out.append([0 if t == 0 else float(c / t), c, t])
    return out


def cal_cov_all(X, y, engine=None):
    """identify_strains...:44-49."""
    return [r[0] for r in stat_cov_all(X, y, engine=engine)]


def get_candidate_arr(X, y, engine=None):
    """identify_strains...:121-134: (argmax strain, its count) of #rows with X=1 and y>1; ties go to
    the lowest index, as Python's stable sort with reverse=True keeps first-seen order."""
    _, covered, _ = strain_hits(X, y, engine=engine)
    best = int(np.argmax(covered))
    return best, int(covered[best])


def get_remainc(dominat, used_kmer, X, py, strain_remainc, engine=None):
    """identify_strains...:94-108: per strain i != dominat, check/all_k over rows not yet used."""
    mask = (np.asarray(used_kmer) == 0).astype(np.uint8)
    total, covered, _ = strain_hits(X, py, row_mask=mask, engine=engine)
    for i in range(total.size):
        if i == dominat:
            continue
        strain_remainc[i] = 0 if total[i] == 0 else covered[i] / total[i]
    return strain_remainc


def _column_values(X, y, j):
    """y restricted to the rows of strain j (the non-zero part of X[:, j] * y needs nothing else: X is 0/1)."""
    if not isinstance(X, StrainMatrix):
        X = StrainMatrix(X)
    rows = X.rows[int(X.col_ptr[j]):int(X.col_ptr[j + 1])]
    return np.asarray(y)[rows]


def optimize_dominat_y(X, y):
    """identify_strains...:136-175: the strain whose rows carry the largest sum of y inside the 5th..95th
    percentile ('nearest') of its non-zero values; the first maximum wins (np.where(...)[0][0])."""
    if not isinstance(X, StrainMatrix):
        X = StrainMatrix(X)
    y = np.asarray(y)
    res = np.zeros(X.n_strains, dtype=np.result_type(y.dtype, np.int64))
    for j in range(X.n_strains):
        v = _column_values(X, y, j)
        nz = v[v != 0]
        if nz.size < 1 or nz.sum() == 0:
            continue
        lo, hi = np.percentile(nz, 5, method="nearest"), np.percentile(nz, 95, method="nearest")
        res[j] = v[(v >= lo) & (v <= hi)].sum()
    return int(np.where(res == np.max(res))[0][0])


def get_avg_depth(dominat, X, y):
    """identify_strains...:109-119: mean of the dominant strain's y values (1 -> 0, zeros dropped) inside their
    25th..75th percentile ('nearest')."""
    v = _column_values(X, y, int(dominat)).copy()
    v[v == 1] = 0
    nz = v[v != 0]
    lo, hi = np.percentile(nz, 25, method="nearest"), np.percentile(nz, 75, method="nearest")
    return np.mean(nz[(nz >= lo) & (nz <= hi)])


def unique_cluster_rows(omatrix, all_cls):
    """detect_strains' `ln` (identify_strains...:183-197): 1 for the rows whose k-mer belongs to exactly one of the
    identified clusters (overlap_matrix.npz column = cluster id - 1), else 0; sparse, no `.A`."""
    import scipy.sparse as sp
    om = sp.load_npz(omatrix) if isinstance(omatrix, (str, os.PathLike)) else sp.csr_matrix(omatrix)
    cols = [int(a) - 1 for a in all_cls]
    ln = np.asarray(om.tocsc()[:, cols].astype(np.int64).sum(axis=1)).ravel()
    ln[ln > 1] = 0
    return ln
