"""Drop-in mirrors of the reference's L1 count adapter and per-node gather, backed by the GPU engine.

Same names, argument meaning and return contracts as the reference functions they replace:

  jellyfish_count(fq_path, db_dir)                       library/identify.py:73-103
                                                         library/identify_low_mem.py:67-90
                                                         library/identify_low_depth.py:46-74
  match_node(match_results, db_dir, node_id, valid_kmers)        library/identify.py:115-127
  match_node_low_depth(...)                                      library/identify_low_depth.py:86-102
  del_outlier(profile)                                           library/identify.py:106-112
  node_coverages(...)   every node at once (-b 1)                library/identify_low_depth.py:113-132

The reference returns a dict {record ordinal: count} whose keys are the valid k-mers
(identify.py:410: valid_kmers = set(match_results.keys())).  Here the same mapping is an
array-backed `CountVector` (a collections.abc.Mapping), and `ValidKmers` is a set-like over the
validity mask whose `&` with a Python set returns a Python set -- so the reference's match_node /
adjust_profile / search code runs unmodified on them, without building 10^7-entry dicts
(SURVEY.md section 7 hard part 6).  No CPU counting path exists here: everything goes through
strainscan_b200.Engine and fails loudly without the CUDA extension or a B200.
"""
import os
from collections.abc import Mapping, Set

import numpy as np

from .engine import Engine

_ENGINE = None
_SETS = {}    # (device, fasta path, k, mtime) -> KmerSet
_READS = {}   # (device, paths, mtimes, shard, n_shards) -> Reads     (the read cache for pass 2..n)
_MAX_CACHED_READS = 2


def default_engine():
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(int(os.environ.get("SS_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _ENGINE


def _db_cache_path(fasta_path, k):
    """Where the binary form of a k-mer FASTA is kept: SS_DB_CACHE=1 -> next to the FASTA (<fasta>.k<k>.ssb200),
    SS_DB_CACHE=<dir> -> in that directory (for read-only databases), unset -> no cache."""
    where = os.environ.get("SS_DB_CACHE", "")
    if not where or where == "0":
        return None
    full = os.path.abspath(fasta_path)
    if where == "1":
        return "%s.k%d.ssb200" % (full, int(k))
    import hashlib
    tag = hashlib.sha1(full.encode()).hexdigest()[:16]
    os.makedirs(where, exist_ok=True)
    return os.path.join(where, "%s_%s.k%d.ssb200" % (os.path.basename(full), tag, int(k)))


def cached_kmerset(engine, fasta_path, k):
    key = (engine.device, os.path.abspath(fasta_path), int(k), os.path.getmtime(fasta_path))
    if key not in _SETS:
        cache = _db_cache_path(fasta_path, k)
        if cache:
            _SETS[key] = engine.kmerset_from_fasta_cached(fasta_path, k, cache)[0]
        else:
            _SETS[key] = engine.kmerset_from_fasta(fasta_path, k)
    return _SETS[key]


def read_paths(fq_path):
    """The reference passes (fq1, fq2) with fq2 == '' for single-end (StrainScan.py:173-181)."""
    if isinstance(fq_path, str):
        return [p for p in fq_path.split(" ") if p]
    return [p for p in fq_path if p]


def cached_reads(engine, fq_path, shard=0, n_shards=1):
    paths = read_paths(fq_path)
    key = (engine.device, tuple(os.path.abspath(p) for p in paths), tuple(os.path.getmtime(p) for p in paths),
           shard, n_shards)
    if key not in _READS:
        while len(_READS) >= _MAX_CACHED_READS:
            _READS.pop(next(iter(_READS))).free()
        _READS[key] = engine.reads_from_files(paths, shard, n_shards)
    return _READS[key]


def drop_caches():
    for r in _READS.values():
        r.free()
    _READS.clear()
    for s in _SETS.values():
        s.free()
    _SETS.clear()


class ValidKmers(Set):
    """valid_kmers (identify.py:410) over a boolean mask; `valid & some_set` -> Python set."""

    def __init__(self, mask):
        self.mask = mask

    def __contains__(self, k):
        return 0 <= k < self.mask.size and bool(self.mask[k])

    def __iter__(self):
        return (int(i) for i in np.nonzero(self.mask)[0])

    def __len__(self):
        return int(self.mask.sum())

    def _and(self, other):
        if isinstance(other, ValidKmers):
            return ValidKmers(self.mask & other.mask)
        arr = np.fromiter(other, dtype=np.int64, count=len(other))
        arr = arr[(arr >= 0) & (arr < self.mask.size)]
        return set(arr[self.mask[arr]].tolist())

    __and__ = _and
    __rand__ = _and


class CountVector(Mapping):
    """match_results (identify.py:96-103): {record ordinal: count}, keys = valid ordinals only."""

    def __init__(self, counts, valid_mask, stats=None):
        self.counts = counts          # np.uint32[n_records], dense by ordinal
        self.valid_mask = valid_mask  # np.bool_[n_records]
        self.stats = stats

    def __getitem__(self, k):
        if not (0 <= k < self.counts.size) or not self.valid_mask[k]:
            raise KeyError(k)
        return int(self.counts[k])

    def __iter__(self):
        return (int(i) for i in np.nonzero(self.valid_mask)[0])

    def __len__(self):
        return int(self.valid_mask.sum())

    def keys(self):
        return ValidKmers(self.valid_mask)

    def valid_kmers(self):
        return ValidKmers(self.valid_mask)

    def to_dict(self):
        idx = np.nonzero(self.valid_mask)[0]
        return dict(zip(idx.tolist(), self.counts[idx].tolist()))


def jellyfish_count(fq_path, db_dir, engine=None, k=31):
    """identify.py:73-103.  fq_path: (fq1, fq2 | ''), db_dir: <DB>/Tree_database.  L1 always counts
    with k = 31 (identify.py:82,86).  gz inputs (identify.py:81) are inflated by the library."""
    eng = engine or default_engine()
    kset = cached_kmerset(eng, os.path.join(db_dir, "kmer.fa"), k)
    reads = cached_reads(eng, fq_path)
    counts, st = eng.count(kset, reads)
    return CountVector(counts, kset.valid, st)


def del_outlier(profile):
    """identify.py:106-112 on an array: drop every value >= 100 * median."""
    profile = np.asarray(profile)
    cutoff = 100 * np.median(profile)
    return profile[profile < cutoff]


def _parse_ints(line):
    """`list(map(int, line.rstrip().split(" ")))` of identify.py:118 as an int64 array (4x faster parser)."""
    try:
        a = np.fromstring(line, dtype=np.int64, sep=" ")
        if a.size == line.strip().count(" ") + 1:      # single-space separated, as the reference's split(" ") requires
            return a
    except (ValueError, DeprecationWarning):
        pass
    return np.array(line.split(), dtype=np.int64)


def _node_ordinals(db_dir, node_id):
    with open(os.path.join(db_dir, "kmers", str(node_id)), "r") as f:
        lines = f.readlines()
    if len(lines) == 0:
        return None
    return np.unique(_parse_ints(lines[0]))   # the reference builds a set


def _profile(match_results, d):
    d = d[(d >= 0) & (d < match_results.counts.size)]
    valid = d[match_results.valid_mask[d]]
    c = match_results.counts[valid]
    prof = c[c > 0]
    if prof.size > 0:
        prof = del_outlier(prof)
    return int(valid.size), prof.tolist()


def match_node(match_results, db_dir, node_id, valid_kmers=None):
    """identify.py:115-127 -> (len(valid_kmer), k_profile).  k_profile's order is unspecified in the
    reference too (set iteration); its consumers use len() and np.mean() only (identify.py:134,238)."""
    d = _node_ordinals(db_dir, node_id)
    if d is None:
        raise IndexError("list index out of range")   # the reference does lines[0] on an empty file
    return _profile(match_results, d)


def match_node_low_depth(match_results, db_dir, node_id, valid_kmers=None):
    """identify_low_depth.py:86-102: empty file or < 1000 valid k-mers -> (0, [])."""
    d = _node_ordinals(db_dir, node_id)
    if d is None:
        return 0, []
    dd = d[(d >= 0) & (d < match_results.counts.size)]
    if int(match_results.valid_mask[dd].sum()) < 1000:
        return 0, []
    return _profile(match_results, d)


def adjust_profile_gather(match_results, db_dir, node_id, delete_positions, valid_kmers=None):
    """Gather part of adjust_profile (identify.py:167-191): the node's k-mer list minus the overlap
    positions (`overlapping_info[leaf][node]`: indices INTO the node's list as written in kmers/<node>,
    identify.py:176-178).  Returns (len(valid_kmer), k_profile) when at least 1000 list entries remain
    (identify.py:181), else None -- the caller then takes the reference's Poisson branch (identify.py:196-228),
    which stays host code."""
    with open(os.path.join(db_dir, "kmers", str(node_id)), "r") as f:
        lines = f.readlines()
    d = _parse_ints(lines[0])                                       # list order matters: positions index it
    delete = np.unique(d[np.asarray(list(delete_positions), dtype=np.int64)]) if len(delete_positions) else np.zeros(0, np.int64)
    ds = np.unique(d)
    if ds.size - delete.size < 1000:
        return None
    return _profile(match_results, np.setdiff1d(ds, delete, assume_unique=True))


def node_coverage_all(match_results, db_dir, node_ids, min_valid=1000):
    """Coverage part of identify_ranks (identify_low_depth.py:113-132) for every node at once:
    {node_id: len(k_profile) / length, or -1 when the node has fewer than `min_valid` valid k-mers or an
    empty list}.  Same numbers as calling match_node_low_depth per node; the per-node vectors are gathers
    of the GPU-produced dense count vector."""
    out = {}
    for nid in node_ids:
        length, prof = match_node_low_depth(match_results, db_dir, nid) if min_valid else match_node(match_results, db_dir, nid)
        out[nid] = -1 if length == 0 else len(prof) / length
    return out


def load_node_csr(db_dir, node_ids):
    """kmers/<node> files -> CSR (ptr uint64[n+1], ordinals uint32[nnz]), lists de-duplicated."""
    ptr = np.zeros(len(node_ids) + 1, dtype=np.uint64)
    parts = []
    for i, nid in enumerate(node_ids):
        d = _node_ordinals(db_dir, nid)
        if d is None:
            d = np.zeros(0, dtype=np.int64)
        parts.append(d.astype(np.uint32))
        ptr[i + 1] = ptr[i] + d.size
    return ptr, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32))
