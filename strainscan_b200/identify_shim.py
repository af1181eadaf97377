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

from ._lib import SS_ERR_NOMEM, StrainScanB200Error
from .engine import Engine

_ENGINE = None
_SETS = {}    # (device, fasta path, k, mtime) -> KmerSet, least recently used first
_READS = {}   # (device, paths, mtimes, shard, n_shards) -> Reads     (the read cache for pass 2..n)
_NODES = {}   # (db_dir, mtime of kmers/) -> {node id: de-duplicated ordinals}; (.., "csr", ids) -> (ptr, ord, NodeIndex)
_MAX_CACHED_READS = 2
# GPU tables kept alive: the L1 set (used by every sample) plus the sets of the clusters under work.  vote_strain_L2
# uses each cluster's set once per run (Vote_...:295-296), and an E. coli-scale table is ~1.8 GB, so the cache is a
# small LRU instead of growing with the number of identified clusters (SS_MAX_CACHED_SETS to change).
_MAX_CACHED_SETS = int(os.environ.get("SS_MAX_CACHED_SETS", "4"))


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
    if key in _SETS:
        _SETS[key] = _SETS.pop(key)                      # most recently used last
        return _SETS[key]
    while len(_SETS) >= max(1, _MAX_CACHED_SETS):
        _SETS.pop(next(iter(_SETS)))                     # evict the least recently used table: its device memory is
                                                         # released when the last reference (a live CountVector?) goes
    cache = _db_cache_path(fasta_path, k)

    def load():
        if cache:
            return engine.kmerset_from_fasta_cached(fasta_path, k, cache)[0]
        return engine.kmerset_from_fasta(fasta_path, k)

    try:
        _SETS[key] = load()
    except StrainScanB200Error as e:
        if e.code != SS_ERR_NOMEM and "out of memory" not in str(e):
            raise
        import gc
        _SETS.clear()                                    # device memory is short: drop every cached table and retry once
        gc.collect()
        _SETS[key] = load()
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
    for v in _NODES.values():
        if isinstance(v, tuple) and hasattr(v[2], "free"):
            v[2].free()
    _NODES.clear()


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

    def __init__(self, counts, valid_mask, stats=None, kset=None, dev=None, engine=None):
        self.counts = counts          # np.uint32[n_records], dense by ordinal
        self.valid_mask = valid_mask  # np.bool_[n_records]
        self.stats = stats
        self.kset, self.dev, self.engine = kset, dev, engine   # the same vector on the GPU (torch.int32), for K4

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
    import torch                                          # owner of the device buffer only
    dev = torch.empty(max(kset.n_records, 1), dtype=torch.int32, device=torch.device("cuda", eng.device))
    st = eng.count_device(kset, reads, dev.data_ptr())
    counts = dev[:kset.n_records].cpu().numpy().view(np.uint32)
    return CountVector(counts, kset.valid, st, kset=kset, dev=dev, engine=eng)


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


def _node_cache(db_dir):
    """Per database: {node id: de-duplicated ordinals of kmers/<node> (None for an empty file)} -- the text of a node
    list is parsed once per process, not once per match_node call (the tree search visits nodes repeatedly and
    identify_low_depth.identify_ranks visits every node twice, identify_low_depth.py:113-132)."""
    kd = os.path.join(db_dir, "kmers")
    key = (os.path.abspath(db_dir), os.path.getmtime(kd) if os.path.isdir(kd) else 0.0)
    if key not in _NODES:
        for old in [k for k in _NODES if k[0] == key[0]]:
            v = _NODES.pop(old)
            if isinstance(v, tuple) and hasattr(v[2], "free"):
                v[2].free()
        _NODES[key] = {}
    return key, _NODES[key]


def _node_ordinals(db_dir, node_id):
    _, cache = _node_cache(db_dir)
    nid = str(node_id)
    if nid not in cache:
        with open(os.path.join(db_dir, "kmers", nid), "r") as f:
            lines = f.readlines()
        cache[nid] = None if len(lines) == 0 else np.unique(_parse_ints(lines[0]))   # the reference builds a set
    return cache[nid]


def _prefetch_node_lists(db_dir, node_ids):
    """Fill the node cache for many nodes at once: the list files parsed and de-duplicated on all host threads
    (ss_node_lists_parse).  Files that are not plain single-space separated ordinals, and unreadable ones, are left to
    _node_ordinals, which has the reference's semantics (and its errors)."""
    _, cache = _node_cache(db_dir)
    todo = [str(n) for n in node_ids if str(n) not in cache]
    if len(todo) < 4:
        return
    from . import _lib
    import ctypes as C
    paths = [os.path.join(db_dir, "kmers", n) for n in todo]
    total = 0
    for p in paths:
        try:
            total += os.path.getsize(p)
        except OSError:
            pass
    lib = _lib.load()
    ptr = np.zeros(len(todo) + 1, dtype=np.uint64)
    out = np.empty(total // 2 + 1, dtype=np.uint32)
    status = np.zeros(len(todo), dtype=np.uint8)
    arr = (C.c_char_p * len(todo))(*[p.encode() for p in paths])
    _lib.check(lib.ss_node_lists_parse(arr, len(todo), 0, ptr.ctypes.data, out.ctypes.data, out.size, status.ctypes.data))
    for i, nid in enumerate(todo):
        if status[i] == 0:
            cache[nid] = out[int(ptr[i]):int(ptr[i + 1])].astype(np.int64)
        elif status[i] == 1:
            cache[nid] = None


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
    empty list}.  Same numbers as calling match_node_low_depth per node.  With a GPU-backed CountVector the
    per-node (length, covered, max) come from ONE K4 launch over a persistent CSR of the whole tree
    (ss_node_index); only nodes whose largest count reaches 100 -- where del_outlier (identify.py:106-112) could
    remove something -- are re-done with the exact host gather."""
    node_ids = list(node_ids)
    out = {}
    if getattr(match_results, "dev", None) is not None and min_valid and node_ids:
        key, _ = _node_cache(db_dir)
        ckey = key + ("csr", tuple(str(n) for n in node_ids))
        if ckey not in _NODES:
            ptr, ords = load_node_csr(db_dir, node_ids)
            _NODES[ckey] = (ptr, ords, match_results.engine.node_index_create(ptr, ords))
        ptr, _, index = _NODES[ckey]
        length, covered, _, mx = match_results.engine.node_index_reduce(match_results.kset, index, match_results.dev.data_ptr())
        for i, nid in enumerate(node_ids):
            if ptr[i + 1] == ptr[i] or length[i] < min_valid:
                out[nid] = -1
            elif mx[i] < 100:                                  # median >= 1, so nothing is >= 100 * median
                out[nid] = int(covered[i]) / int(length[i])
            else:
                ln, prof = match_node_low_depth(match_results, db_dir, nid)
                out[nid] = -1 if ln == 0 else len(prof) / ln
        return out
    _prefetch_node_lists(db_dir, node_ids)
    for nid in node_ids:
        length, prof = match_node_low_depth(match_results, db_dir, nid) if min_valid else match_node(match_results, db_dir, nid)
        out[nid] = -1 if length == 0 else len(prof) / length
    return out


def load_node_csr(db_dir, node_ids):
    """kmers/<node> files -> CSR (ptr uint64[n+1], ordinals uint32[nnz]), lists de-duplicated."""
    ptr = np.zeros(len(node_ids) + 1, dtype=np.uint64)
    parts = []
    _prefetch_node_lists(db_dir, node_ids)
    for i, nid in enumerate(node_ids):
        d = _node_ordinals(db_dir, nid)
        if d is None:
            d = np.zeros(0, dtype=np.int64)
        parts.append(d.astype(np.uint32))
        ptr[i + 1] = ptr[i] + d.size
    return ptr, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32))
