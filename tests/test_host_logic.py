"""Host-side mirrors of the reference interface (identify_shim / l2_shim) against the oracle's
restatement of the reference's dict/list code.  CPU only: count vectors come from the oracle here;
the GPU tests check that the CUDA path produces the same vectors."""
import os
import types

import numpy as np
import pytest

from oracle import adapters
from strainscan_b200 import identify_shim, l2_shim
from tests import util


@pytest.fixture(scope="module")
def l1_case():
    rng = np.random.default_rng(42)
    G = util.rand_genome(rng, 80_000)
    fa = util.make_db(rng, G, 31, 9000, both_strands=True, lower_frac=0.02, junk=24)
    fq = util.make_reads(rng, G, 5000, 150)
    d = adapters.count_dense(fa, 31, [fq])
    valid = (d.in_set & d.is_last).astype(bool)
    cv = identify_shim.CountVector(d.cnt.astype(np.uint32), valid)
    return d, cv, adapters.l1_match_results(d)


def test_count_vector_is_the_reference_dict(l1_case):
    d, cv, ref = l1_case
    assert len(cv) == len(ref)
    assert cv.to_dict() == ref
    assert set(cv.keys()) == set(ref.keys())
    some = list(ref)[:50]
    assert all(cv[k] == ref[k] for k in some)
    invalid = int(np.nonzero(~cv.valid_mask)[0][0])
    with pytest.raises(KeyError):
        cv[invalid]
    with pytest.raises(KeyError):
        cv[10 ** 9]
    assert invalid not in cv and some[0] in cv


def test_valid_kmers_intersection_matches_set_semantics(l1_case):
    d, cv, ref = l1_case
    vk = cv.valid_kmers()
    ref_valid = set(ref.keys())
    rng = np.random.default_rng(1)
    dset = set(int(x) for x in rng.integers(0, len(d.cnt), 3000))
    assert (vk & dset) == (ref_valid & dset)
    assert (dset & vk) == (ref_valid & dset)
    assert len(vk) == len(ref_valid)


def test_match_node_equals_reference(l1_case, tmp_path):
    d, cv, ref = l1_case
    ref_valid = set(ref.keys())
    rng = np.random.default_rng(2)
    os.makedirs(tmp_path / "kmers")
    # hot k-mer to trigger the 100x-median outlier trim
    cv.counts[int(np.nonzero(cv.valid_mask & (cv.counts > 0))[0][0])] = 100000
    ref = {k: int(cv.counts[k]) for k in ref}
    for node, n in enumerate([1, 50, 1500, 4000, 0]):
        ords = rng.integers(0, len(d.cnt), n)
        if node == 3:
            ords = np.concatenate([ords, ords[:100], np.nonzero(cv.counts == 100000)[0]])   # duplicates + hot k-mer
        with open(tmp_path / "kmers" / str(node), "w") as f:
            f.write("".join("%d " % x for x in ords))
        if n == 0:
            open(tmp_path / "kmers" / str(node), "w").close()
            with pytest.raises(IndexError):
                identify_shim.match_node(cv, str(tmp_path), node, cv.valid_kmers())
            assert identify_shim.match_node_low_depth(cv, str(tmp_path), node) == (0, [])
            continue
        length, prof = identify_shim.match_node(cv, str(tmp_path), node, cv.valid_kmers())
        rl, rp = adapters.match_node(ref, ords, ref_valid)
        assert length == rl and sorted(prof) == rp
        ll, lp = identify_shim.match_node_low_depth(cv, str(tmp_path), node)
        rl2, rp2 = adapters.match_node(ref, ords, ref_valid, min_valid=1000)
        assert ll == rl2 and sorted(lp) == rp2
    ptr, ordinals = identify_shim.load_node_csr(str(tmp_path), [0, 1, 2, 3, 4])
    assert ptr[-1] == ordinals.size and ptr[5] == ptr[4]


def test_adjust_profile_gather_and_node_coverage(l1_case, tmp_path):
    """identify.py:167-191 (gather part) and identify_low_depth.py:113-132 (coverage of every node)."""
    d, cv, ref = l1_case
    ref_valid = set(ref.keys())
    rng = np.random.default_rng(9)
    os.makedirs(tmp_path / "kmers")
    lists = {0: rng.integers(0, len(d.cnt), 5000), 1: rng.integers(0, len(d.cnt), 1200), 2: rng.integers(0, len(d.cnt), 30)}
    for node, ords in lists.items():
        with open(tmp_path / "kmers" / str(node), "w") as f:
            f.write("".join("%d " % x for x in ords))
    open(tmp_path / "kmers" / "3", "w").close()
    for node, n_del in ((0, 0), (0, 700), (0, 4500), (1, 100), (1, 400)):
        pos = rng.choice(len(lists[node]), n_del, replace=False).tolist()
        got = identify_shim.adjust_profile_gather(cv, str(tmp_path), node, pos, cv.valid_kmers())
        want = adapters.adjust_profile_gather(ref, lists[node], pos, ref_valid)
        if want is None:
            assert got is None
        else:
            assert got[0] == want[0] and sorted(got[1]) == want[1]
    cov = identify_shim.node_coverage_all(cv, str(tmp_path), [0, 1, 2, 3])
    for node in (0, 1, 2):
        length, prof = adapters.match_node(ref, lists[node], ref_valid, min_valid=1000)
        assert cov[node] == (-1 if length == 0 else len(prof) / length)
    assert cov[3] == -1 and cov[2] == -1


def test_del_outlier_equals_reference():
    rng = np.random.default_rng(0)
    for _ in range(20):
        prof = rng.integers(1, 5, 200).tolist() + [1000, 500, 399, 400]
        assert sorted(identify_shim.del_outlier(prof).tolist()) == sorted(adapters.del_outlier(prof))


def test_remove_1_and_kid_order(golden_cases):
    case = [c for c in golden_cases if c["name"] == "l2_k21"][0]
    # shuffle the records so that header kids are a non-identity permutation
    lines = case["fasta"].strip().split("\n")
    recs = [(lines[i], lines[i + 1]) for i in range(0, len(lines), 2)]
    rng = np.random.default_rng(4)
    recs = [recs[i] for i in rng.permutation(len(recs))]
    fa = "".join("%s\n%s\n" % r for r in recs)
    d = adapters.count_dense(fa.encode(), case["k"], [r.encode() for r in case["reads"]])
    kset = types.SimpleNamespace(flags=(d.in_set | (d.is_last << 1) | (d.raw_upper << 2)).astype(np.uint8),
                                 header_ids=d.header_id)
    py_o = l2_shim.remove_1(d.cnt.astype(np.uint32), kset)
    assert np.array_equal(py_o, adapters.l2_py_o_from_dump(fa, case["dump"]))
    assert np.array_equal(py_o, adapters.l2_py_o_from_dump(case["fasta"], case["dump"]))


def test_read_paths_follow_reference_convention():
    assert identify_shim.read_paths(("a.fq", "")) == ["a.fq"]
    assert identify_shim.read_paths(("a.fq.gz", "b.fq.gz")) == ["a.fq.gz", "b.fq.gz"]
    assert identify_shim.read_paths("a.fq b.fq") == ["a.fq", "b.fq"]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_dominant_strain_and_depth_mirrors_equal_the_dense_reference_code(seed):
    """optimize_dominat_y / get_avg_depth / the unique-cluster row mask (identify_strains...:109-119, 136-175,
    183-197) on the sparse columns against the dense restatement, incl. empty strains, all-zero y and ties."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n, S = 4000, 9
    X = (rng.random((n, S)) < rng.uniform(0.02, 0.5, S)).astype(np.int8)
    X[:, 4] = 0                                           # a strain without rows
    X[:, 7] = X[:, 2]                                     # a tie: the first maximum wins
    y = rng.poisson(6, n).astype(np.int64) * (rng.random(n) < 0.7)
    y[y == 1] = 0                                         # remove_1 ran before (Vote_...:386-403)
    if seed == 3:
        y[:] = 0
        y[:50] = 7
    dom, res = adapters.optimize_dominat_y(X, y)
    assert l2_shim.optimize_dominat_y(sp.csr_matrix(X), y) == dom
    assert l2_shim.optimize_dominat_y(l2_shim.StrainMatrix(X), y) == dom
    if res[dom] > 0:
        a, b = adapters.get_avg_depth(dom, X, y), l2_shim.get_avg_depth(dom, l2_shim.StrainMatrix(X), y)
        assert a == b and type(a) is type(b)
    om = (rng.random((n, 5)) < 0.4).astype(np.int8)
    for cls in ([1], [2, 5], [1, 3, 4]):
        assert np.array_equal(l2_shim.unique_cluster_rows(sp.csr_matrix(om), cls), adapters.unique_cluster_rows(om, cls))


_REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "library")


@pytest.mark.skipif(not os.path.isdir(_REF_LIB), reason="baseline/_ref not present (python baseline/setup_ref.py)")
def test_per_strain_restatements_pinned_against_the_reference_module():
    """The oracle's restatements of identify_strains_L2_Enet_Pscan_new_sp.py (and therefore the mirrors checked
    against them) against the reference's own functions, imported from the sandbox, on random dense inputs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, warnings
import numpy as np
warnings.simplefilter("ignore")
sys.path[:0] = [%r, %r, %r]
import ss_compat; ss_compat.apply()
import identify_strains_L2_Enet_Pscan_new_sp as ref
from oracle import adapters
for seed in range(6):
    rng = np.random.default_rng(seed)
    n, S = 3000, 7
    X = (rng.random((n, S)) < rng.uniform(0.05, 0.5, S)).astype(np.int8)
    X[:, 5] = X[:, 1]
    y = rng.poisson(5, n).astype(np.int64) * (rng.random(n) < 0.8)
    y[y == 1] = 0
    dom, _ = adapters.optimize_dominat_y(X, y)
    assert dom == ref.optimize_dominat_y(X, y), seed
    assert adapters.get_avg_depth(dom, X, y) == ref.get_avg_depth(dom, X, y), seed
    assert [c for c, _ in adapters.stat_cov_all(X, y)] == [ref.stat_cov(X[:, j], y)[1] for j in range(S)]
    cand = adapters.candidate_counts(X, y)
    c, chk = ref.get_candidate_arr(X.T, y)
    assert cand[c] == chk == max(cand) and c == int(np.argmax(cand))
    used = (rng.random(n) < 0.3).astype(np.int64)
    rem = ref.get_remainc(0, used, X.T.astype(np.int64), y, {})
    for j, (check, all_k) in enumerate(adapters.remain_cov(used, X, y)):
        if j:
            assert rem[j] == (0 if all_k == 0 else check / all_k)
print("ok")
''' % (os.path.join(root, "baseline", "shims"), _REF_LIB, root)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-3000:]


def test_l2_mirrors_reject_a_transposed_matrix():
    """the per-strain mirrors take the reference's orientations (cal_cov_all: rows x strains,
    get_candidate_arr / get_remainc: strains x rows) and raise when the side that must equal len(y) does not --
    before anything reaches the GPU (so this runs without one)."""
    rng = np.random.default_rng(0)
    X = (rng.random((50, 7)) < 0.3).astype(np.int8)          # rows x strains
    y = rng.poisson(3, 50).astype(np.int64)
    with pytest.raises(ValueError):
        l2_shim.get_candidate_arr(X, y)                      # rows x strains where strains x rows is expected
    with pytest.raises(ValueError):
        l2_shim.get_remainc(0, np.zeros(50), X, y, {})
    with pytest.raises(ValueError):
        l2_shim.cal_cov_all(X.T, y)                          # strains x rows where rows x strains is expected
    with pytest.raises(ValueError):
        l2_shim.get_candidate_arr(l2_shim.StrainMatrix(X), y)    # the sparse form: pass X.T
    with pytest.raises(ValueError):
        l2_shim.cal_cov_all(l2_shim.StrainMatrix(X).T, y)
    with pytest.raises(ValueError):
        l2_shim.optimize_dominat_y(X.T, y)
    m = l2_shim.StrainMatrix(X)
    assert m.T.T is m and m.T.shape == (7, 50) and m.shape == (50, 7)


def test_kid_row_order_is_one_rule():
    """Rows follow the FASTA header ids only when these are an exact permutation of 1..n; every other layout keeps
    FASTA order (l2_shim.remove_1, dist.reduce_then_remove_1 and ss_l2_finalize share this rule)."""
    f = l2_shim.kid_row_order
    assert f(np.array([], dtype=np.uint64)) is None
    assert f(np.arange(1, 9, dtype=np.uint64)) is None                       # identity
    assert f(np.ones(8, dtype=np.uint64)) is None                            # kmer.fa: all ">1"
    assert f(np.array([1, 2, 2, 4], dtype=np.uint64)) is None                # duplicates
    assert f(np.array([2, 3, 4, 5], dtype=np.uint64)) is None                # not 1..n
    assert f(np.array([0, 1, 2, 3], dtype=np.uint64)) is None
    assert f(np.array([3, 1, 4, 2], dtype=np.uint64)).tolist() == [1, 3, 0, 2]


def test_dense_matrices_become_column_lists_like_numpy_says():
    """l2_shim._nonzero_lists (threaded host helper ss_dense_nonzero_lists behind the dense forms of cal_cov_all /
    get_candidate_arr / get_remainc) against np.flatnonzero: every dtype the reference's matrices come in, C / F /
    strided layouts, both orientations, empty shapes, -0.0 and NaN."""
    from strainscan_b200 import l2_shim
    rng = np.random.default_rng(0)

    def by_numpy(X, axis):
        n = X.shape[axis]
        lists = [np.flatnonzero(X[j] if axis == 0 else X[:, j]).astype(np.uint32) for j in range(n)]
        ptr = np.zeros(n + 1, dtype=np.uint64)
        if n:
            ptr[1:] = np.cumsum([x.size for x in lists])
        return ptr, (np.concatenate(lists) if lists else np.zeros(0, np.uint32))

    for dt in (np.bool_, np.int8, np.uint8, np.int16, np.int32, np.int64, np.uint64, np.float32, np.float64, np.float16):
        for shape in ((1, 1), (3, 1000), (1000, 3), (64, 5000), (257, 33), (0, 5), (5, 0)):
            X = (rng.random(shape) < 0.1).astype(dt)
            if np.dtype(dt).kind == "f" and X.size:
                X.flat[0] = -0.0
                X.flat[-1] = np.nan
            if np.dtype(dt).kind == "i" and X.size:
                X[X != 0] = -1
            for arr in (X, np.asfortranarray(X), X[:, ::2] if shape[1] > 1 else X, X.T):
                for axis in (0, 1):
                    ptr, idx = l2_shim._nonzero_lists(arr, axis)
                    rp, ri = by_numpy(arr, axis)
                    assert np.array_equal(ptr, rp) and np.array_equal(idx, ri), (dt, shape, axis)
    # the two orientations of one matrix give the same CSC
    Xt = (rng.random((9, 20_000)) < 0.3).astype(np.int64)
    a = l2_shim._csc_of_strains_by_rows(Xt)
    b = l2_shim._csc_of_rows_by_strains(np.ascontiguousarray(Xt.T))
    c = l2_shim._csc_of_rows_by_strains(Xt.T)
    assert a[:2] == b[:2] == c[:2] == (20_000, 9)
    assert all(np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]) for k in (2, 3))


class _NumpyStrainEngine:
    """Stand-in for Engine's strain-matrix handle (K5 on the GPU): the same three sums in NumPy, so that the host side
    of the dense mirrors -- orientation checks, dense -> column lists, masks, ratios -- runs without a GPU."""

    def strain_matrix_create(self, col_ptr, rows, n_rows):
        return (np.asarray(col_ptr).copy(), np.asarray(rows).copy(), int(n_rows))

    def strain_matrix_free(self, handle):
        pass

    def strain_matrix_reduce(self, handle, y, row_mask=None):
        col_ptr, rows, n_rows = handle
        y = np.asarray(y, dtype=np.int64)
        assert y.size == n_rows
        S = col_ptr.size - 1
        total, covered, ssum = (np.zeros(S, dtype=np.uint64) for _ in range(3))
        for j in range(S):
            r = rows[int(col_ptr[j]):int(col_ptr[j + 1])].astype(np.int64)
            if row_mask is not None:
                r = r[np.asarray(row_mask)[r] != 0]
            total[j] = r.size
            hit = y[r] > 1
            covered[j] = int(hit.sum())
            ssum[j] = int(y[r][hit].sum())
        return total, covered, ssum


@pytest.mark.parametrize("dtype", [np.int64, np.int8, np.float64, np.bool_])
def test_dense_mirrors_equal_the_reference_restatements_without_a_gpu(dtype):
    """cal_cov_all / get_candidate_arr / get_remainc of l2_shim on the dense matrices the reference hands them
    (identify_strains_L2_Enet_Pscan_new_sp.py:241-334), every layout, against oracle/adapters.py's restatements."""
    from strainscan_b200 import l2_shim
    rng = np.random.default_rng(3)
    eng = _NumpyStrainEngine()
    n, S = 3000, 11
    pX = (rng.random((n, S)) < 0.2).astype(dtype)                       # rows x strains, as pX = X.A
    y = (rng.poisson(3, n) * (rng.random(n) < 0.7)).astype(np.int64)
    used = (rng.random(n) < 0.3).astype(np.int64)
    want_cov = [0 if t == 0 else c / t for c, t in adapters.stat_cov_all(pX, y)]
    for ix in (pX, np.asfortranarray(pX), pX.T.copy().T):
        assert l2_shim.cal_cov_all(ix, y, engine=eng) == want_cov
    cand = adapters.candidate_counts(pX, y)
    best = int(np.argmax(cand))
    for ixt in (pX.T, np.ascontiguousarray(pX.T)):
        assert l2_shim.get_candidate_arr(ixt, y, engine=eng) == (best, cand[best])
    want = {i: (0 if a == 0 else c / a) for i, (c, a) in enumerate(adapters.remain_cov(used, pX, y)) if i != 2}
    for ixt in (pX.T, np.ascontiguousarray(pX.T)):
        assert l2_shim.get_remainc(2, used, ixt, y, {}, engine=eng) == want


def test_node_lists_parsed_in_bulk_equal_the_one_by_one_parser(tmp_path, monkeypatch):
    """identify_shim.load_node_csr: the kmers/<node> files through the threaded host parser (ss_node_lists_parse) and
    through the one-file-at-a-time NumPy parser give the same cache and the same CSR -- plain lists with and without the
    trailing blank, duplicates, an empty file, and the shapes the fast parser must leave alone (a sign, two blanks, a
    value over 2^32, a second line)."""
    rng = np.random.default_rng(4)
    db = tmp_path / "Tree_database"
    (db / "kmers").mkdir(parents=True)
    lists = {}
    for v in range(40):
        a = rng.integers(0, 5_000_000, int(rng.integers(1, 3000)))
        if v % 5 == 0:
            a = np.concatenate([a, a[:20]])
        lists[str(v)] = a
        (db / "kmers" / str(v)).write_text(" ".join(map(str, a.tolist())) + (" \n" if v % 2 else "\n"))
    (db / "kmers" / "empty").write_text("")
    odd = {"sign": "5 -3 7\n", "blanks": "5  7\n", "big": "5 99999999999 7\n", "lines": "9 8 7\n1 2 3\n", "cr": "4 2 4\r\n"}
    for name, text in odd.items():
        (db / "kmers" / name).write_text(text)
    ids = list(lists) + ["empty", "sign", "big", "lines", "cr"]
    identify_shim._NODES.clear()
    ptr, ords = identify_shim.load_node_csr(str(db), ids)
    _, cache = identify_shim._node_cache(str(db))
    fast = {k: (None if v is None else v.copy()) for k, v in cache.items()}
    identify_shim._NODES.clear()
    monkeypatch.setattr(identify_shim, "_prefetch_node_lists", lambda *a: None)
    ptr2, ords2 = identify_shim.load_node_csr(str(db), ids)
    _, slow = identify_shim._node_cache(str(db))
    assert np.array_equal(ptr, ptr2) and np.array_equal(ords, ords2)
    assert set(fast) == set(slow)
    for k in slow:
        assert (fast[k] is None and slow[k] is None) or (fast[k].dtype == slow[k].dtype and np.array_equal(fast[k], slow[k])), k
    for k, a in lists.items():
        assert np.array_equal(slow[k], np.unique(a))
    assert slow["empty"] is None and slow["lines"].tolist() == [7, 8, 9] and slow["cr"].tolist() == [2, 4]
    identify_shim._NODES.clear()
