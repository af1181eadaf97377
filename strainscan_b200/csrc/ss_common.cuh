// ss_common.cuh -- shared definitions for the sm_100a match+count engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/strainscan_b200.h"

// ---- tiling of the FASTQ text ---------------------------------------------------------------
// A tile owns SS_TILE window-start positions; SS_HALO more bytes are staged with it so that a
// window that starts in the tile can finish (k-1 <= 31 bytes needed; 64 keeps 32-byte runs).
#define SS_TILE     8192
#define SS_HALO     64
#define SS_RUN      32                               // text bytes classified per thread
#define SS_THREADS  (SS_TILE / SS_RUN)               // 256
#define SS_NRUN     ((SS_TILE + SS_HALO) / SS_RUN)   // 258 runs incl. halo
#define SS_STAGES   2
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

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t ss_bucket_of(uint64_t key, uint64_t n_buckets) {
    return __umul64hi(ss_mix(key), n_buckets);
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
};

void ss_set_error(const std::string &msg);
int ss_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define SS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ss_cuda_fail(e_, #x, __FILE__, __LINE__); } while (0)
