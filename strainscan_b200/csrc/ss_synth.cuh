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

#ifdef __CUDACC__
struct ss_synth_db_plan {
    const unsigned long long *node_off;   // n_nodes + 1 prefix of per-node record counts (even numbers)
    uint32_t n_nodes;
    const uint32_t *blk_list;             // blocks grouped by owner depth
    const uint32_t *blk_off;              // max_depth + 2 offsets into blk_list
    unsigned long long n_records;
    unsigned long long perm_a, perm_c;    // record i holds logical record (a*i + c) % n_records
};
cudaError_t ss_launch_synth_db(const ss_synth_params &p, const ss_synth_db_plan &plan, uint8_t *text,
                               uint32_t *node_of_record, cudaStream_t st);
cudaError_t ss_launch_synth_reads_impl(const ss_synth_params &p, uint8_t *text, uint64_t n_reads,
                                       uint64_t first_read, cudaStream_t st);
#endif
