"""Host-side helpers for the synthetic workload of BASELINE.json configs 2/3/5 (bench + full-size
parity tests).  The generators themselves are CUDA kernels (csrc/ss_synth.cu); this module only
fills the parameter struct and sizes the search-tree nodes.  Reference scale figures:
README.md:114,117 (E. coli 1433 strains / 823 clusters, S. aureus 1627 / 202) and
StrainScan_build.py:75-78 (1000..30000 k-mers per tree node)."""
import numpy as np

from ._lib import SynthParams

_M = (1 << 64) - 1


def _h3(seed, a, b):
    """ss_h3 of csrc/ss_synth.cuh in Python ints."""
    x = (seed + 0x9E3779B97F4A7C15 * (a + 1) + 0xC2B2AE3D27D4EB4F * (b + 1)) & _M
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & _M
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) & _M
    x ^= x >> 31
    return x


def depth(node):
    return int(node + 1).bit_length() - 1


def max_depth(n_leaves):
    return depth(2 * n_leaves - 2)


def _p32(x):
    return int(min(max(x, 0.0), 1.0) * 0xFFFFFFFF)


def default_params(n_leaves=823, genome_len=5_000_000, seed=1, k=31, read_len=150, header_len=24,
                   block_len=4096, sources=None, p_offtarget=0.05, p_sub=0.005, p_n=0.001, snp_rate=0.002):
    """sources: list of (cluster index, strain index, abundance); default = three clusters 60/30/10 %."""
    p = SynthParams()
    p.seed, p.n_leaves, p.genome_len, p.block_len = seed, n_leaves, genome_len, block_len
    p.k, p.read_len, p.header_len = k, read_len, header_len
    if sources is None:
        picks = sorted({n_leaves // 7, n_leaves // 2, (5 * n_leaves) // 6})
        ab = [0.6, 0.3, 0.1][:len(picks)]
        sources = [(c, i, a) for i, (c, a) in enumerate(zip(picks, ab))]
    tot = sum(a for _, _, a in sources)
    p.n_sources = len(sources)
    cum = 0.0
    for i, (leaf, strain, a) in enumerate(sources):
        cum += a / tot
        p.source_leaf[i], p.source_strain[i], p.source_cum[i] = leaf, strain, _p32(cum)
    p.source_cum[len(sources) - 1] = 0xFFFFFFFF
    p.p_offtarget, p.p_sub, p.p_n, p.snp_rate = _p32(p_offtarget), _p32(p_sub), _p32(p_n), _p32(snp_rate)
    return p


def blocks_per_depth(p):
    md = max_depth(p.n_leaves)
    out = np.zeros(md + 1, dtype=np.int64)
    for b in range(p.genome_len // p.block_len):
        out[_h3(p.seed, 0xB10C, b) % (md + 1)] += 1
    return out


def node_sizes(p, lo=1000, hi=30000, seed=0):
    """Records per tree node (both strands counted, so even), U[lo, hi] as Build_tree.py's min/max
    node k-mer numbers, clipped to what the node's own blocks can supply."""
    rng = np.random.default_rng(seed)
    n_nodes = 2 * p.n_leaves - 1
    bpd = blocks_per_depth(p)
    per_block = p.block_len - p.k + 1
    sizes = rng.integers(lo // 2, hi // 2 + 1, n_nodes) * 2
    for v in range(n_nodes):
        sizes[v] = min(int(sizes[v]), int(bpd[depth(v)]) * per_block * 2)
    return sizes.astype(np.uint32)


def node_csr(node_of_record, n_nodes):
    """CSR (node -> record ordinals) = the content of Tree_database/kmers/<node>."""
    order = np.argsort(node_of_record, kind="stable").astype(np.uint32)
    counts = np.bincount(node_of_record, minlength=n_nodes)
    ptr = np.zeros(n_nodes + 1, dtype=np.uint64)
    ptr[1:] = np.cumsum(counts)
    return ptr, order
