// ss_synth.cuh -- counter-based synthetic pan-genome shared by the DB and read generators.
// Bench/test tooling (BASELINE.json configs 2,3,5 name *synthetic* genomes and reads); it is not on
// the identification path.  Everything is a pure function of (seed, indices), so any slice of the
// database or of the read set can be regenerated independently on any GPU.
//
// Model: a complete binary search tree over n_leaves clusters in heap order (node 0 = root, children
// 2i+1 / 2i+2, leaves n_leaves-1 .. 2*n_leaves-2).  A cluster genome of genome_len bases is cut into
// blocks of block_len; block b belongs to the ancestor of the leaf at depth d_b = H(b) % (max_depth+1)
// (clamped to the leaf), and its bases are H(owner, pos) & 3 -- so all clusters below a node share that
// node's blocks and nothing else does: the k-mers of a block are specific to its owner node, which is
// what Build_tree.py extracts into Tree_database/kmers/<node>.  Strains of a cluster differ from the
// cluster genome by SNPs at rate snp_rate.
#pragma once
#include <stdint.h>

#include "../../include/strainscan_b200.h"   // ss_synth_params

#ifdef __CUDACC__
#define SS_HD __host__ __device__ __forceinline__
#else
#define SS_HD static inline
#endif

SS_HD uint64_t ss_h3(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (a + 1) + 0xC2B2AE3D27D4EB4Full * (b + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

SS_HD uint32_t ss_synth_depth(uint32_t node) {   // depth of heap node: floor(log2(node + 1))
    uint32_t d = 0, x = node + 1;
    while (x > 1) { x >>= 1; d++; }
    return d;
}
SS_HD uint32_t ss_synth_ancestor(uint32_t node, uint32_t depth_now, uint32_t depth_want) {
    while (depth_now > depth_want) { node = (node - 1) >> 1; depth_now--; }
    return node;
}
SS_HD uint32_t ss_synth_max_depth(uint32_t n_leaves) { return ss_synth_depth(2 * n_leaves - 2); }

// depth at which block b is owned
SS_HD uint32_t ss_synth_block_depth(const ss_synth_params *p, uint32_t block) {
    return (uint32_t)(ss_h3(p->seed, 0xB10Cull, block) % (ss_synth_max_depth(p->n_leaves) + 1));
}
// owner node of position pos in the genome of leaf node `leaf`
SS_HD uint32_t ss_synth_owner(const ss_synth_params *p, uint32_t leaf, uint32_t pos) {
    uint32_t dl = ss_synth_depth(leaf), db = ss_synth_block_depth(p, pos / p->block_len);
    return db >= dl ? leaf : ss_synth_ancestor(leaf, dl, db);
}
// base (0..3 = ACGT) of the cluster genome
SS_HD uint32_t ss_synth_node_base(const ss_synth_params *p, uint32_t owner, uint32_t pos) {
    return (uint32_t)(ss_h3(p->seed, 0x6E0DEull + owner, pos) >> 17) & 3u;
}
// base of strain `strain` of leaf `leaf`
SS_HD uint32_t ss_synth_strain_base(const ss_synth_params *p, uint32_t leaf, uint32_t strain, uint32_t pos) {
    uint32_t b = ss_synth_node_base(p, ss_synth_owner(p, leaf, pos), pos);
    uint64_t h = ss_h3(p->seed ^ 0x57A1ull, ((uint64_t)leaf << 20) + strain, pos);
    if ((uint32_t)h < p->snp_rate) b = (b + 1u + (uint32_t)((h >> 40) % 3u)) & 3u;
    return b;
}

// ---- record writers, shared by the device kernels (ss_synth.cu) and the host generator (tools/synth_host.cpp) ----
struct ss_synth_db_plan {
    const unsigned long long *node_off;   // n_nodes + 1 prefix of per-node record counts (even numbers)
    uint32_t n_nodes;
    const uint32_t *blk_list;             // blocks grouped by owner depth
    const uint32_t *blk_off;              // max_depth + 2 offsets into blk_list
    unsigned long long n_records;
    unsigned long long perm_a, perm_c;    // record i holds logical record (a*i + c) % n_records
};

// FASTA record i of the synthetic kmer.fa: ">1\nKMER\n" (k + 4 bytes at text + i * (k + 4))
SS_HD void ss_synth_write_db_record(const ss_synth_params &p, const ss_synth_db_plan &plan, unsigned long long i,
                                    uint8_t *text, uint32_t *node_of_record) {
    unsigned long long q = (plan.perm_a * i + plan.perm_c) % plan.n_records;
    uint32_t lo = 0, hi = plan.n_nodes;       // node = last v with node_off[v] <= q
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (plan.node_off[mid] <= q) lo = mid; else hi = mid;
    }
    uint32_t v = lo;
    unsigned long long j = q - plan.node_off[v];
    uint32_t strand = (uint32_t)(j & 1ull);
    uint32_t t = (uint32_t)(j >> 1);
    uint32_t d = ss_synth_depth(v);
    uint32_t nb = plan.blk_off[d + 1] - plan.blk_off[d];
    uint32_t blk = plan.blk_list[plan.blk_off[d] + (t % nb)];
    uint32_t pos = blk * p.block_len + (t / nb);
    const uint32_t rec = p.k + 4;
    uint8_t *o = text + i * rec;
    o[0] = '>'; o[1] = '1'; o[2] = '\n';
    for (uint32_t x = 0; x < p.k; x++) {
        uint32_t gp = strand ? pos + p.k - 1 - x : pos + x;
        uint32_t b = ss_synth_node_base(&p, v, gp);
        if (strand) b = 3u - b;
        o[3 + x] = "ACGT"[b];
    }
    o[3 + p.k] = '\n';
    if (node_of_record) node_of_record[i] = v;
}

// 4-line FASTQ record of read `rid` at o (header_len + 1 + read_len + 1 + 2 + read_len + 1 bytes)
SS_HD void ss_synth_write_read(const ss_synth_params &p, unsigned long long rid, uint8_t *o) {
    const uint32_t L = p.read_len, H = p.header_len;
    o[0] = '@'; o[1] = 'S';               // header: '@' 'S' then the zero-padded decimal read id
    {
        unsigned long long x = rid;
        for (int i = (int)H - 1; i >= 2; i--) { o[i] = '0' + (uint8_t)(x % 10); x /= 10; }
    }
    o[H] = '\n';
    uint64_t h = ss_h3(p.seed, 0x4EAD5ull, rid);
    bool off = (uint32_t)h < p.p_offtarget;
    uint32_t src = 0;
    uint64_t h2 = ss_h3(p.seed, 0x4EAD6ull, rid);
    if (!off) {
        uint32_t a = (uint32_t)(h >> 32);
        while (src + 1 < p.n_sources && a > p.source_cum[src]) src++;
    }
    uint32_t leaf = p.n_leaves - 1 + p.source_leaf[src];
    uint32_t strain = p.source_strain[src];
    uint32_t pos = (uint32_t)(h2 % (uint64_t)(p.genome_len - L + 1));
    uint32_t strand = (uint32_t)(h2 >> 63);
    uint8_t *s = o + H + 1;
    for (uint32_t i = 0; i < L; i++) {
        uint32_t b;
        if (off) {
            b = (uint32_t)(ss_h3(p.seed ^ 0x0FF7ull, rid, i) >> 9) & 3u;
        } else {
            uint32_t gp = strand ? pos + L - 1 - i : pos + i;
            b = ss_synth_strain_base(&p, leaf, strain, gp);
            if (strand) b = 3u - b;
        }
        uint64_t e = ss_h3(p.seed ^ 0xE440ull, rid, i);
        uint8_t c;
        if ((uint32_t)e < p.p_n) c = 'N';
        else {
            if ((uint32_t)(e >> 32) < p.p_sub) b = (b + 1u + (uint32_t)((e >> 20) % 3u)) & 3u;
            c = "ACGT"[b];
        }
        s[i] = c;
    }
    s[L] = '\n'; s[L + 1] = '+'; s[L + 2] = '\n';
    uint8_t *qv = s + L + 3;
    for (uint32_t i = 0; i < L; i++)   // Phred+33 range '!'..'J': includes '@' and '+' (also as first char)
        qv[i] = '!' + (uint8_t)((ss_h3(p.seed ^ 0x9A1ull, rid, i) >> 13) % 42u);
    qv[L] = '\n';
}

#ifdef __cplusplus
#include <string>
#include <vector>
// host side of the database generator: the arrays ss_synth_db_plan points at.  Returns "" or what is wrong.
struct ss_synth_db_layout {
    std::vector<unsigned long long> node_off;
    std::vector<uint32_t> blk_list, blk_off;
    unsigned long long n_records = 0, perm_a = 1, perm_c = 0;
};
static inline std::string ss_synth_db_layout_build(const ss_synth_params *p, const uint32_t *node_sizes, uint32_t n_nodes,
                                                   ss_synth_db_layout &L) {
    if (n_nodes != 2 * p->n_leaves - 1) return "n_nodes must be 2*n_leaves-1";
    uint32_t maxd = ss_synth_max_depth(p->n_leaves);
    uint32_t n_blocks = p->genome_len / p->block_len;
    std::vector<std::vector<uint32_t>> by_depth(maxd + 1);
    for (uint32_t b = 0; b < n_blocks; b++) by_depth[ss_synth_block_depth(p, b)].push_back(b);
    L.blk_list.clear(); L.blk_off.assign(maxd + 2, 0);
    for (uint32_t d = 0; d <= maxd; d++) {
        L.blk_off[d] = (uint32_t)L.blk_list.size();
        L.blk_list.insert(L.blk_list.end(), by_depth[d].begin(), by_depth[d].end());
    }
    L.blk_off[maxd + 1] = (uint32_t)L.blk_list.size();
    L.node_off.assign(n_nodes + 1, 0);
    const uint32_t per_block = p->block_len - p->k + 1;
    for (uint32_t v = 0; v < n_nodes; v++) {
        uint32_t d = ss_synth_depth(v);
        uint64_t avail = (uint64_t)by_depth[d].size() * per_block * 2;
        if (node_sizes[v] & 1u) return "node sizes must be even (both strands)";
        if (node_sizes[v] > avail) return "node larger than its owned blocks";
        L.node_off[v + 1] = L.node_off[v] + node_sizes[v];
    }
    const uint64_t N = L.node_off[n_nodes];
    L.n_records = N;
    if (N >= (1ull << 32)) return "too many records";
    if (N) {
        uint64_t a = (uint64_t)((double)N * 0.6180339887) | 1ull;
        auto gcd = [](uint64_t x, uint64_t y) { while (y) { uint64_t t = x % y; x = y; y = t; } return x; };
        while (gcd(a, N) != 1) a += 2;
        L.perm_a = a; L.perm_c = N / 3;
    }
    return "";
}
#endif

#ifdef __CUDACC__
cudaError_t ss_launch_synth_db(const ss_synth_params &p, const ss_synth_db_plan &plan, uint8_t *text,
                               uint32_t *node_of_record, cudaStream_t st);
cudaError_t ss_launch_synth_reads_impl(const ss_synth_params &p, uint8_t *text, uint64_t n_reads,
                                       uint64_t first_read, cudaStream_t st);
#endif
