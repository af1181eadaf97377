"""Drop-in mirrors of the reference's L2 count block and per-strain hit statistics.

  count_cluster(input_fq, fq2, db_dir, ksize)   library/Vote_Strain_L2_Lasso_new_sp.py:348-403
        -> py_o, the int64 vector vote_strain_L2() hands to detect_strains(): rows in kid order
           (kid - 1 = record ordinal of C<id>/all_kmer.fasta), absent -> 0, count == 1 -> 0
           (remove_1, Vote_...:312-322).
  stat_cov / cal_cov_all(ix, iy)                library/identify_strains_L2_Enet_Pscan_new_sp.py:33-49
  get_candidate_arr(ix, iy)                     ...:121-134   (ix: strains x rows, as the reference passes npXt)
  get_remainc(dominat, used_kmer, pXt_tem, py, strain_remainc)        ...:94-108   (pXt_tem: strains x rows)
        integer reductions on the GPU over the CSC form of all_strains_re.npz (K5; the matrix is uploaded once
        into a persistent device handle); the ratios are formed on the host exactly as the reference does.
        Orientations are the reference's; a matrix whose row count does not match len(y) raises ValueError.
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


def kid_row_order(header_ids):
    """The ONE rule for the row order of py_o, shared with ss_l2_finalize (csrc/ss_api.cu) and dist.py: the reference
    orders rows by kid (all_kid.pkl, Vote_...:386-388), and the builder writes kid as the FASTA header, a 1..n counter
    (Build_kmer_sets_...:397-399).  So: rows follow the header ids when they are an exact permutation of 1..n; any
    other header layout (all ">1" as in kmer.fa, gaps, duplicates) keeps FASTA record order.  Returns the gather
    index (py_o = c[index]) or None for the identity."""
    hid = np.asarray(header_ids)
    n = hid.size
    if n == 0 or np.array_equal(hid, np.arange(1, n + 1, dtype=hid.dtype)):
        return None
    if hid.min() < 1 or hid.max() > n:
        return None
    seen = np.zeros(n, dtype=bool)
    seen[(hid - 1).astype(np.int64)] = True
    if not seen.all():
        return None
    return np.argsort(hid, kind="stable")


def remove_1(counts, kset):
    """Vote_...:312-322 + :386-388 on the dense vector: raw string not a dumped key -> 0,
    count == 1 -> 0, rows in kid order (kid_row_order)."""
    c = np.where((kset.flags & SS_REC_RAW_UPPER) != 0, counts, 0).astype(np.int64)
    c[c == 1] = 0
    order = kid_row_order(kset.header_ids)
    return c if order is None else c[order]


class StrainMatrix:
    """CSC view of the 0/1 strain matrix (all_strains_re.npz, Recls_withR_new.py:110-112), rows x strains.
    `.T` is the strains x rows orientation the reference's get_candidate_arr / get_remainc take (a view, no copy).
    With `engine` (or at first use on the GPU) the CSC arrays are uploaded once into a persistent device handle
    (ss_strain_matrix_create), so that the ~45 reductions Pre_Scan asks for per cluster move only y and the mask."""

    def __init__(self, X, engine=None, _csc=None):
        if _csc is not None:
            self.n_rows, self.n_strains, self.col_ptr, self.rows = _csc
        elif isinstance(X, np.ndarray):
            # the dense rows x strains array detect_strains builds (pX = X.A): the rows of every strain, read in place
            # (scipy's dense -> CSC conversion is several times slower)
            if X.ndim != 2:
                raise ValueError("StrainMatrix: a 2-D matrix is expected, got shape %r" % (X.shape,))
            self.n_rows, self.n_strains, self.col_ptr, self.rows = _csc_of_rows_by_strains(X)
        else:
            import scipy.sparse as sp
            X = sp.csc_matrix(X)
            X.eliminate_zeros()
            X.sort_indices()
            self.n_rows, self.n_strains = X.shape
            self.col_ptr = X.indptr.astype(np.uint64)
            self.rows = X.indices.astype(np.uint32)
        self.shape = (self.n_rows, self.n_strains)
        self._dev = None
        self._eng = engine

    @classmethod
    def load(cls, npz_path, engine=None):
        import scipy.sparse as sp
        return cls(sp.load_npz(npz_path), engine)

    @property
    def T(self):
        return _StrainsByRows(self)

    def device(self, engine):
        """The persistent device copy (created at first use, freed with the object)."""
        if self._dev is None or self._eng is not engine:
            self.free()
            self._eng = engine
            self._dev = engine.strain_matrix_create(self.col_ptr, self.rows, self.n_rows)
        return self._dev

    def free(self):
        if self._dev is not None:
            self._eng.strain_matrix_free(self._dev)
            self._dev = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _nonzero_lists(X, axis):
    """(ptr, idx): the index lists of the non-zeros of a dense 2-D array -- axis 0: one list per X[j] holding the
    positions i with X[j, i] != 0; axis 1: one list per column holding the rows, ascending.  C- and F-contiguous arrays
    of bool / integer / float32 / float64 go through the threaded host helper ss_dense_nonzero_lists (NumPy's
    flatnonzero, list by list on one thread, was most of the dense mirrors' time); anything else through NumPy."""
    n_lists = X.shape[axis]
    if X.dtype.kind in "biuf" and X.dtype.itemsize in (1, 2, 4, 8) and not (X.dtype.kind == "f" and X.dtype.itemsize < 4) \
            and X.dtype.isnative and (X.flags.c_contiguous or X.flags.f_contiguous) and X.size:
        from . import _lib
        lib = _lib.load()
        if X.flags.c_contiguous:
            A, ax = X, axis
        else:                                  # the same memory read as the C-contiguous transpose
            A, ax = X.T, 1 - axis
        if A.shape[1 - ax] < 2 ** 32:
            ptr = np.zeros(n_lists + 1, dtype=np.uint64)
            args = (A.ctypes.data, A.shape[0], A.shape[1], A.dtype.itemsize, int(A.dtype.kind == "f"), ax, 0)
            _lib.check(lib.ss_dense_nonzero_lists(*args, ptr.ctypes.data, None))
            np.cumsum(ptr, out=ptr)
            idx = np.empty(int(ptr[-1]), dtype=np.uint32)
            _lib.check(lib.ss_dense_nonzero_lists(*args, ptr.ctypes.data, idx.ctypes.data if idx.size else None))
            return ptr, idx
    lists = [np.flatnonzero(X[j] if axis == 0 else X[:, j]).astype(np.uint32) for j in range(n_lists)]
    ptr = np.zeros(n_lists + 1, dtype=np.uint64)
    if n_lists:
        ptr[1:] = np.cumsum([p.size for p in lists])
    idx = np.concatenate(lists) if lists else np.zeros(0, dtype=np.uint32)
    return ptr, idx


def _csc_of_strains_by_rows(Xt):
    """(n_rows, n_strains, col_ptr, rows) of a dense STRAINS x ROWS array: the rows of every strain."""
    S, R = Xt.shape
    col_ptr, rows = _nonzero_lists(Xt, 0)
    return R, S, col_ptr, rows


def _csc_of_rows_by_strains(X):
    """(n_rows, n_strains, col_ptr, rows) of a dense ROWS x STRAINS array (no transposed copy is made)."""
    R, S = X.shape
    col_ptr, rows = _nonzero_lists(X, 1)
    return R, S, col_ptr, rows


class _StrainsByRows:
    """StrainMatrix seen as strains x rows (what the reference calls pXt / npXt)."""

    def __init__(self, m):
        self.m = m
        self.shape = (m.n_strains, m.n_rows)

    @property
    def T(self):
        return self.m


def _rows_by_strains(X, y_len, what):
    """X in the rows x strains orientation (cal_cov_all's ix) as a StrainMatrix; raises when its row count is not len(y)."""
    m = X if isinstance(X, StrainMatrix) else None
    if m is None:
        if isinstance(X, _StrainsByRows):
            raise ValueError("%s: got a strains x rows matrix where rows x strains is expected" % what)
        m = StrainMatrix(X)
    if m.n_rows != y_len:
        raise ValueError("%s: matrix has %d rows but y has %d entries (wrong orientation?)" % (what, m.n_rows, y_len))
    return m


def _strains_by_rows(X, y_len, what):
    """X in the strains x rows orientation (get_candidate_arr's ix, get_remainc's pXt_tem) as a rows x strains
    StrainMatrix; raises when its column count is not len(y)."""
    if isinstance(X, _StrainsByRows):
        m = X.m
    elif isinstance(X, StrainMatrix):
        raise ValueError("%s: got a rows x strains StrainMatrix where strains x rows is expected (pass X.T)" % what)
    else:
        if getattr(X, "ndim", 2) != 2 or X.shape[1] != y_len:
            raise ValueError("%s: matrix is %s but y has %d entries: expected strains x rows" % (what, getattr(X, "shape", None), y_len))
        if isinstance(X, np.ndarray):
            m = StrainMatrix(None, _csc=_csc_of_strains_by_rows(X))     # strains x rows, row by row
        else:
            import scipy.sparse as sp
            m = StrainMatrix(sp.csr_matrix(X).T)      # CSR of strains x rows = CSC of rows x strains
    if m.n_rows != y_len:
        raise ValueError("%s: matrix has %d rows but y has %d entries" % (what, m.n_rows, y_len))
    return m


def strain_hits(X, y, row_mask=None, engine=None):
    """Per strain: (total, covered, sum) = (#rows X=1, #rows X=1 and y>1, sum of those y); rows limited
    to row_mask != 0 when given.  X: rows x strains (StrainMatrix, scipy sparse or dense)."""
    y = np.ascontiguousarray(y, dtype=np.int64)
    m = _rows_by_strains(X, y.size, "strain_hits")
    if row_mask is not None and np.asarray(row_mask).size != y.size:
        raise ValueError("strain_hits: row_mask has %d entries, y has %d" % (np.asarray(row_mask).size, y.size))
    eng = engine or identify_shim.default_engine()
    return eng.strain_matrix_reduce(m.device(eng), y, row_mask)


def stat_cov_all(X, y, engine=None):
    """[[cov, valid_kmr, total_kmr], ...] per strain = stat_cov (identify_strains...:33-43)."""
    total, covered, _ = strain_hits(X, y, engine=engine)
    out = []
    for t, c in zip(total.tolist(), covered.tolist()):
        out.append([0 if t == 0 else float(c / t), c, t])
    return out


def cal_cov_all(ix, iy, engine=None):
    """identify_strains...:44-49; ix: rows x strains, as the reference passes pX."""
    return [r[0] for r in stat_cov_all(ix, iy, engine=engine)]


def get_candidate_arr(ix, iy, row_mask=None, engine=None):
    """identify_strains...:121-134: (argmax strain, its count) of #rows with ix=1 and iy>1; ties go to the lowest
    index, as Python's stable sort with reverse=True keeps first-seen order.  ix: STRAINS x ROWS, as the reference
    passes npXt (identify_strains...:332) -- dense, scipy sparse, or `StrainMatrix.T` (then `row_mask` stands in
    for the masking npXt carries: rows with row_mask == 0 are skipped)."""
    iy = np.asarray(iy)
    m = _strains_by_rows(ix, iy.size, "get_candidate_arr")
    _, covered, _ = strain_hits(m, iy, row_mask=row_mask, engine=engine)
    best = int(np.argmax(covered))
    return best, int(covered[best])


def get_remainc(dominat, used_kmer, pXt_tem, py, strain_remainc, engine=None):
    """identify_strains...:94-108: per strain i != dominat, check/all_k over the rows not yet used.
    pXt_tem: STRAINS x ROWS, as the reference passes it (identify_strains...:316); used_kmer: 0/1 per row."""
    py = np.asarray(py)
    m = _strains_by_rows(pXt_tem, py.size, "get_remainc")
    used = np.asarray(used_kmer)
    if used.size != py.size:
        raise ValueError("get_remainc: used_kmer has %d entries, py has %d" % (used.size, py.size))
    mask = (used == 0).astype(np.uint8)
    total, covered, _ = strain_hits(m, py, row_mask=mask, engine=engine)
    for i in range(total.size):
        if i == dominat:
            continue
        strain_remainc[i] = 0 if total[i] == 0 else covered[i] / total[i]
    return strain_remainc


def _column_values(X, y, j):
    """y restricted to the rows of strain j (the non-zero part of X[:, j] * y needs nothing else: X is 0/1)."""
    if not isinstance(X, StrainMatrix):
        X = StrainMatrix(X)
    if X.n_rows != np.asarray(y).size:
        raise ValueError("matrix has %d rows but y has %d entries (expected rows x strains)" % (X.n_rows, np.asarray(y).size))
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
