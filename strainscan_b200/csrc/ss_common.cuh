// ss_common.cuh -- shared definitions for the sm_100a match+count engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/strainscan_b200.h"

// ---- tiling of the FASTQ text ---------------------------------------------------------------
// A tile owns SS_TILE window-start positions; SS_HALO more bytes are staged with it so that a
// window that starts in the tile can finish (k-1 <= 31 bytes needed).
#define SS_TILE     8192
#define SS_HALO     32                               // >= k-1 bytes after the tile (one more run)
#define SS_RUN      32                               // text bytes classified per thread
#define SS_THREADS  (SS_TILE / SS_RUN)               // 256 consumer threads
#define SS_CONSUMERS (SS_THREADS / 32)               // 8 consumer warps, each owns SS_SUB bytes of a tile
#define SS_CTA_THREADS (SS_THREADS + 32)             // + 1 producer warp
#define SS_SUB      1024                             // line-index granularity = one consumer warp's slice
#define SS_WRUNS    33                               // runs a warp classifies: 32 + 1 halo
#define SS_QCAP     (32 + 32)                        // deferred table-probe queue entries per warp
#define SS_STAGES   3
#define SS_TEXT_PAD (SS_TILE + 256)                  // '\n' padding after the text on device

#define SS_EMPTY 0xFFFFFFFFFFFFFFFFull
#define SS_NOSLOT 0xFFFFFFFFu

// One probe bucket = one 32-byte sector = 4 full 64-bit keys (first base in the low bits).
struct __align__(32) ss_bucket { unsigned long long k[4]; };

// Invertible 64-bit mixer; bucket = mulhi(mix(key), n_buckets) so n_buckets need not be a power of 2.
__host__ __device__ __forceinline__ uint64_t ss_mix(uint64_t x) {
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    return x;
}

// The probe path's hash: one xorshift-multiply round, split into two 32-bit words.
//   hi -> table bucket   = mulhi32(hi, n_buckets)
//   lo -> filter word    = mulhi32(lo, n_filter_words), filter bits from (lo * golden) top bits
__host__ __device__ __forceinline__ uint64_t ss_mix1(uint64_t x) {
    x ^= x >> 31; x *= 0xd6e8feb86659fd93ull;
    return x;
}
// 4 bits of a 64-bit filter word: two in each 32-bit half
__host__ __device__ __forceinline__ uint64_t ss_filter_mask(uint32_t lo) {
    uint32_t m = lo * 0x9E3779B1u;
    uint32_t a = (1u << (m >> 27)) | (1u << ((m >> 22) & 31u));
    uint32_t b = (1u << ((m >> 17) & 31u)) | (1u << ((m >> 12) & 31u));
    return (uint64_t)a | ((uint64_t)b << 32);
}

struct ss_table_view {
    const ss_bucket *buckets;      // n_buckets x 32 B
    uint32_t *slot_cnt;            // 4 * n_buckets + 1 counters (last = the all-ones key, k = 32 only)
    uint64_t n_buckets;
    uint64_t kmask;                // low 2k bits
    uint32_t vmask;                // low k bits
    int k;
    int has_ones;                  // the key 0xFFFF...F (k = 32 poly-T) is in the set
    const unsigned long long *filter;   // L2-resident blocked Bloom filter (64-bit blocks) or NULL
    uint32_t n_filter_words;
};

void ss_set_error(const std::string &msg);
int ss_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define SS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ss_cuda_fail(e_, #x, __FILE__, __LINE__); } while (0)
