// ss_kernels.cuh -- launch wrappers of ss_probe.cu (host side declarations) and tunables.
#pragma once
#include "ss_common.cuh"

#ifndef SS_PROBE_MIN_CTAS
#define SS_PROBE_MIN_CTAS 5      // __launch_bounds__ min CTAs/SM (256 threads each)
#endif

int ss_probe_ctas_per_sm();

// K1: line index of n_tiles (+1) units; `tile_line` must hold ss_index_words(n_tiles) 32-bit words (index + scratch)
size_t ss_index_words(uint32_t n_tiles);
cudaError_t ss_launch_index(const uint8_t *text, uint32_t n_tiles, uint32_t *tile_line, uint32_t line_base,
                            int n_sm, cudaStream_t st);
cudaError_t ss_launch_probe(const uint8_t *text, uint64_t text_len, uint32_t n_tiles, const uint32_t *tile_line,
                            const ss_table_view &tv, unsigned long long *stats, unsigned long long *err, int n_sm,
                            cudaStream_t st);
cudaError_t ss_launch_probe_range(const uint8_t *text, uint64_t text_len, uint32_t tile_lo, uint32_t tile_hi,
                                  const uint32_t *sub_line, const ss_table_view &tv, unsigned long long *stats,
                                  unsigned long long *err, int n_sm, cudaStream_t st);
int ss_bin_scan_ctas_per_sm();
cudaError_t ss_launch_bin_scan(const uint8_t *text, uint64_t text_len, uint32_t tile_lo, uint32_t tile_hi,
                               const uint32_t *sub_line, const ss_table_view &tv, const ss_bin_view &bv, bool use_filter,
                               unsigned long long *stats, unsigned long long *err, int n_sm, cudaStream_t st);
cudaError_t ss_launch_bin_probe(const ss_bin_view &bv, const uint32_t *n_chunks, const ss_table_view &tv,
                                unsigned long long *stats, const unsigned long long *scan_stats, int n_sm,
                                cudaStream_t st);
cudaError_t ss_launch_insert(const uint64_t *keys, const uint8_t *rec_ok, uint64_t n, unsigned long long *slots,
                             uint64_t n_buckets, uint32_t *slot_of, uint32_t *last_ord,
                             unsigned long long *n_distinct, cudaStream_t st);
cudaError_t ss_launch_filter_build(const uint64_t *keys, const uint8_t *rec_ok, uint64_t n, ss_fword *filter,
                                   uint32_t n_words, uint64_t kmask, cudaStream_t st);
cudaError_t ss_launch_flags(const uint32_t *slot_of, const uint32_t *last_ord, uint64_t n, uint8_t *flags,
                            cudaStream_t st);
cudaError_t ss_launch_gather(const uint32_t *slot_of, const uint32_t *slot_cnt, uint64_t n, uint32_t *dense,
                             cudaStream_t st);
// K3b: dense[record] = counter of the record's slot (dense is zeroed here), counters cleared for the next pass
cudaError_t ss_launch_scatter(uint32_t *slot_cnt, const uint32_t *ord_of_slot, uint64_t n_buckets, const uint32_t *dup,
                              uint64_t n_dup, const uint32_t *slot_of, uint64_t n_records, uint32_t *dense, int n_sm,
                              cudaStream_t st);
cudaError_t ss_launch_l2_finalize(const uint32_t *dense, const uint8_t *flags, const uint32_t *row_of, uint64_t n,
                                  long long *py_o, cudaStream_t st);
cudaError_t ss_launch_node_reduce(const uint32_t *dense, const uint8_t *flags, const unsigned long long *node_ptr,
                                  const uint32_t *ordinals, uint32_t n_nodes, uint64_t n_records, uint32_t *length,
                                  uint32_t *covered, unsigned long long *sum, uint32_t *max_count, cudaStream_t st);
cudaError_t ss_launch_strain_reduce(const unsigned long long *col_ptr, const uint32_t *rows, uint32_t n_strains,
                                    const long long *y, const uint8_t *row_mask, unsigned long long *total,
                                    unsigned long long *covered, unsigned long long *sum, cudaStream_t st);
cudaError_t ss_launch_random_gather(const void *buf, uint64_t n_sectors, uint64_t n_probes, uint64_t seed,
                                    unsigned long long *sink, int n_sm, cudaStream_t st);

struct ss_member;
// device inflate of `n_members` independent gzip members: out[tab[i].out_off ...] = inflate(comp + tab[i].comp_off);
// *err receives the smallest failing member index (initialise to 0xFFFFFFFF), *next is scratch
cudaError_t ss_launch_gunzip(const uint8_t *comp, const ss_member *tab, uint32_t n_members, uint8_t *out,
                             unsigned int *next, unsigned int *err, int n_sm, cudaStream_t st);
