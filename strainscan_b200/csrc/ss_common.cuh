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
#ifndef SS_PROBE_UNROLL
#define SS_PROBE_UNROLL 3                                // groups (filter / table probes) in flight per warp iteration
#endif
#define SS_QCAP     (32 + 32 * SS_PROBE_UNROLL)          // deferred table-probe queue entries per warp
#ifndef SS_STAGES
#define SS_STAGES   1                                // per-warp TMA ring depth: the next unit is requested right after
#endif                                               // classification and lands during the (long) probe phase
#define SS_TEXT_PAD (8192 + 256)                     // '\n' padding after the text on device

// Experiment, off by default (-DSS_FILTER_PAIRS=1 builds it): key the L2-resident filter by (k-1)-mers -- it then holds
// the first and the last k-1 bases of every k-mer of the set, and the scan asks it only about the (k-1)-mers at EVEN
// text positions: the one at position j answers for the k-mer that starts at j (its prefix) and for the k-mer that
// starts at j-1 (its suffix), i.e. 17 filter loads per 32 k-mers instead of 32.  Bit-exact (all parity tests pass), same
// table-probe rate (2.6 %), but SLOWER on B200: probe 7.39 ms vs 6.29 ms per 10 M reads.  Halving the L1 tag lookups and
// L2 requests does not pay for the ballot / shuffle / select added to every group and for the half-empty load
// instructions: the kernel is bound by issue slots and load latency, not by the request rate.
#ifndef SS_FILTER_PAIRS
#define SS_FILTER_PAIRS 0
#endif

// Filter word: 32-bit blocks (all four bits of a key in one 32-bit word: 32-bit load / pattern / test in the probe
// loop) or the first design's 64-bit blocks (two bits in each half).  Same bytes per key either way.
#ifndef SS_FILTER_WORD32
#define SS_FILTER_WORD32 1
#endif
#if SS_FILTER_WORD32
typedef uint32_t ss_fword;
#else
typedef unsigned long long ss_fword;
#endif

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
// two chained 32x32+64 multiply-adds (IMAD.WIDE, FMA pipe) with one XOR between them: the low word of the first
// product is folded into the second multiplicand and the whole first product is the second one's addend, so both
// output words depend on every key bit.  Three instructions (the first form swapped the halves of the first
// product instead, which cost a three-XOR register swap per k-mer for the same bucket / filter statistics).
__device__ __forceinline__ void ss_hash2(uint32_t k0, uint32_t k1, uint32_t &hh, uint32_t &hl) {
    uint64_t t = (uint64_t)k0 * 0xD6E8FEB9ull + 0x9E3779B97F4A7C15ull;
    uint64_t u = (uint64_t)(k1 ^ (uint32_t)t) * 0xC2B2AE35ull + t;
    hh = (uint32_t)u;
    hl = (uint32_t)(u >> 32);
}
// Pattern-based blocked Bloom filter: the filter word of a key is mulhi32(hl, n_filter_words), and the key sets in it
// the 4 bits of pattern ss_pat_index(hh); the probe kernel keeps the SS_NPAT patterns in shared memory (one LDS
// instead of ~8 ALU ops).  The index is taken from the bits of hh that, masked, ARE the byte offset into the pattern
// table (bits 2..11 for 4-byte words): one AND in the probe loop, no shift.  It must not come from hl: the word index
// uses hl's top ~24 bits, so keys of one word would share the top bits of such an index and 128 instead of 1024
// patterns would be in play per word -- measured as a table-probe rate of 4.6 % instead of 2.9 %.
#define SS_NPAT 1024
#define SS_FW_SHIFT (sizeof(ss_fword) == 4 ? 2 : 3)
#define ss_pat_index(hh) (((hh) >> SS_FW_SHIFT) & (SS_NPAT - 1u))

__device__ __forceinline__ ss_fword ss_filter_pattern(uint32_t i) {
    uint32_t x = (i + 1u) * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
    uint32_t a = (1u << (x & 31u)) | (1u << ((x >> 5) & 31u));
    uint32_t b = (1u << ((x >> 10) & 31u)) | (1u << ((x >> 15) & 31u));
#if SS_FILTER_WORD32
    return a | b;
#else
    return (uint64_t)a | ((uint64_t)b << 32);
#endif
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
    const ss_fword *filter;             // L2-resident blocked Bloom filter or NULL
    uint32_t n_filter_words;
};

// ---- binned probing (large table + many table probes) ------------------------------------------
// The k-mers that pass the filter are not probed where they are found: they are appended to one of P
// bins = contiguous ranges of the table (bin = top bits of the bucket hash), in chunks of
// SS_BIN_CHUNK keys that a warp owns until they are full, and a second kernel probes bin after bin, so
// that the table range (and its counters) a bin touches stays in L2 instead of costing one random
// 128-byte DRAM line per probe.
#define SS_BIN_CHUNK 128u            // keys per chunk (1 KiB, eight full 128-byte lines)
#define SS_BIN_PMAX  64u
struct ss_bin_view {
    unsigned long long *keys;        // P * cap chunks of SS_BIN_CHUNK keys
    uint32_t *fill;                  // keys in each chunk
    uint32_t *n_chunks;              // [P] chunks handed out per bin; [SS_BIN_PMAX] = overflow flag
    uint32_t cap;                    // chunks per bin
    uint32_t shift;                  // bin = hash_hi >> shift  (P = 2^(32 - shift), P >= 2)
    uint32_t P;
};

void ss_set_error(const std::string &msg);
int ss_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define SS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ss_cuda_fail(e_, #x, __FILE__, __LINE__); } while (0)
