// ss_gunzip.cu -- DEFLATE on the device: one warp per independent gzip member (BGZF block).
//
// Compressed bytes cross PCIe (3-4x fewer than the text) and the FASTQ text is produced directly in
// HBM, in front of the probe kernel.  Replaces the host-side `zcat a b |` of library/identify.py:82 /
// library/Vote_Strain_L2_Lasso_new_sp.py:359,367 for blocked-gzip inputs; ordinary single-stream gzip
// has no independent units and stays on the ingest threads (ss_ingest.cu).  The decoder is the same
// source as the host's (ss_inflate.cuh): Huffman decoding is a serial bit-dependency chain, so lane 0
// of each warp walks the stream with the decode tables in shared memory; parallelism comes from the
// thousands of members in flight (a 32 MiB batch holds ~2-4 k of them).
#include "ss_common.cuh"
#include "ss_inflate.cuh"
#include "ss_ingest.cuh"
#include "ss_kernels.cuh"

#define SS_GUNZIP_WARPS 1        // one decoder per CTA: its tables sit at a fixed shared-memory address
#ifndef SS_GUNZIP_MINCTAS
#define SS_GUNZIP_MINCTAS 32      // 64 registers: 32 one-lane decoders per SM (the CTA limit; 7 KB of tables each)
#endif

__global__ void __launch_bounds__(SS_GUNZIP_WARPS * 32, SS_GUNZIP_MINCTAS)
ss_gunzip_kernel(const uint8_t *__restrict__ comp, const ss_member *__restrict__ tab, uint32_t n_members,
                 uint8_t *__restrict__ out, unsigned int *__restrict__ next, unsigned int *__restrict__ err) {
    __shared__ ssi_tables s_tab[SS_GUNZIP_WARPS];
    if ((threadIdx.x & 31u) != 0) return;
    ssi_tables &t = s_tab[threadIdx.x >> 5];
    while (true) {
        uint32_t m = atomicAdd(next, 1u);                      // members differ in size: dynamic hand-out
        if (m >= n_members) break;
        const ss_member mem = tab[m];
        ssi_stream s;
        ssi_stream_init(s, comp + mem.comp_off, comp + mem.comp_off + mem.comp_len);
        uint8_t *w = out + mem.out_off;
        // the decoder keeps going while a full slack of room is left; a valid member stops at isize
        int rc = ssi_inflate(s, t, &w, out + mem.out_off + mem.isize + SSI_OUT_SLACK);
        if (rc != SSI_OK || s.out_total != mem.isize) atomicMin(err, m);
    }
}

cudaError_t ss_launch_gunzip(const uint8_t *comp, const ss_member *tab, uint32_t n_members, uint8_t *out,
                             unsigned int *next, unsigned int *err, int n_sm, cudaStream_t st) {
    if (n_members == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(next, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    uint32_t grid = (n_members + SS_GUNZIP_WARPS - 1) / SS_GUNZIP_WARPS;
    uint32_t cap = (uint32_t)n_sm * 32u;
    if (grid > cap) grid = cap;
    ss_gunzip_kernel<<<grid, SS_GUNZIP_WARPS * 32, 0, st>>>(comp, tab, n_members, out, next, err);
    return cudaGetLastError();
}
