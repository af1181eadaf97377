// ss_common.cuh -- shared definitions for the sm_100a match+count engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/strainscan_b200.h"

// ---- tiling of the FASTQ text ---------------------------------------------------------------
// A unit ("tile") owns SS_TILE window-start positions = 31 runs of 32 bytes; one more run (SS_HALO,
// >= k-1 bytes) is staged with it so that a window that starts in the unit can finish.  Unit + halo
// = 1024 bytes = one bulk TMA copy, 32 runs = one classification pass of a warp (lane = run).
#define SS_RUN      32                               // text bytes classified per lane
#define SS_TILE     992                              // window starts per unit (31 runs)
#define SS_HALO     32
#define SS_SUB      SS_TILE                          // line-index granularity = one unit
#define SS_WRUNS    32                               // runs classified per unit (31 + halo)
#define SS_CTA_WARPS 8
#define SS_CTA_THREADS (SS_CTA_WARPS * 32)
#define SS_QCAP     (32 + 32)                        // deferred table-probe queue entries per warp
#define SS_STAGES   2                                // per-warp TMA ring depth
#define SS_TEXT_PAD (8192 + 256)                     // '\n' padding after the text on device

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

// The probe path's hash of a packed k-mer (k0 = low 32 bits, k1 = high 32 bits): two 32x32->64
// multiplies (FMA pipe) folded crosswise.  hh -> table bucket = mulhi32(hh, n_buckets);
// hl -> filter word = mulhi32(hl, n_filter_words) and the four filter bits.
#ifdef __CUDACC__
__device__ __forceinline__ void ss_hash2(uint32_t k0, uint32_t k1, uint32_t &hh, uint32_t &hl) {
    uint64_t t = (uint64_t)(k0 ^ 0x9E3779B9u) * 0xD6E8FEB9ull;
    uint64_t u = (uint64_t)(k1 ^ 0x85EBCA6Bu) * 0xC2B2AE35ull;
    hh = (uint32_t)(t >> 32) ^ (uint32_t)u;
    hl = (uint32_t)t ^ (uint32_t)(u >> 32);
}
// 4 bits of a 64-bit filter word, two in each 32-bit half; bit positions = low 5 bits of four
// high-half products (shift amounts wrap mod 32)
__device__ __forceinline__ void ss_filter_mask2(uint32_t hl, uint32_t &ma, uint32_t &mb) {
    uint32_t p = __umulhi(hl, 0x9E3779B1u), q = __umulhi(hl, 0x85EBCA77u);
    ma = __funnelshift_l(0u, 1u, p) | __funnelshift_l(0u, 1u, p >> 5);
    mb = __funnelshift_l(0u, 1u, q) | __funnelshift_l(0u, 1u, q >> 5);
}
#endif

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
