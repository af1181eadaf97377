// ss_probe.cu -- sm_100a kernels of the match+count hot path.
//
//   K1  ss_line_index_kernel newline count per text unit + chained (decoupled look-back) scan ->
//                            line index at each unit start, one pass      (streaming HBM)
//   K3  ss_probe_kernel      fused: TMA-staged text tile -> SWAR classify (newline / ACGT / case
//                            fold) -> 2-bit pack + window-valid bitmap in shared memory ->
//                            per-position k-mer extract -> one 32-byte-sector probe ->
//                            red.global.add on the slot counter    (random HBM sectors)
//   K2  ss_insert_kernel     open-addressing table build (CAS)
//   K3b ss_gather_kernel     slot counters -> dense per-record vector
//
// What it replaces: the inside of `jellyfish count --if` (parser, rolling 2-bit encoder, hash
// lookup, atomic add) and `jellyfish dump -c` + the Python dump parse, as invoked at
// library/identify.py:82-101 and library/Vote_Strain_L2_Lasso_new_sp.py:357-403.
#include "ss_common.cuh"
#include "ss_kernels.cuh"
#include <cstdlib>

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA (cp.async.bulk -> UBLKCP), 256-bit sector load, RED
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one probe = one 32-byte sector, read-only path, no L1 allocation (random, no reuse)
__device__ __forceinline__ void ld_bucket(const ss_bucket *p, unsigned long long &a, unsigned long long &b,
                                          unsigned long long &c, unsigned long long &d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
}
// one filter word of the L2-resident blocked Bloom filter, kept in L2 with evict_last
__device__ __forceinline__ unsigned long long ld_filter(const unsigned long long *p, uint64_t policy) {
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint32_t ld_filter(const uint32_t *p, uint64_t policy) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void red_add_u32(uint32_t *p, uint32_t v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// SWAR byte classification of one 32-bit word (4 text bytes)
// ---------------------------------------------------------------------------------------------
// 0x80 in every byte of (w ^ pat) that is NON-zero (exact, no cross-byte carries); used by the line index kernel
__device__ __forceinline__ uint32_t nz7(uint32_t w, uint32_t pat) {
    uint32_t v = w ^ pat;
    return ((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v;
}
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
// truth-table bytes of lop3 for inputs a = 0xF0, b = 0xCC, c = 0xAA
#define SS_LA 0xF0
#define SS_LB 0xCC
#define SS_LC 0xAA
// Classification by bit planes.  s_i = w >> i carries bit i of every byte at that byte's bit 0 (what spills in from
// the neighbouring byte lands in the higher bits of the lane, which the final 0x01010101 mask drops), and each
// property is a Boolean function of the planes, three inputs per LOP3:
//   newline 0x0A = 00001010b                                   -> ~s7 ~s6 ~s5 ~s4 s3 ~s2 s1 ~s0
//   A 0x41, C 0x43, G 0x47, T 0x54 with bit 5 (case) ignored   -> ~s7 s6 ~s3 and (s4, s2, s1, s0) in {0001, 0011, 0111, 1100}
//   2-bit code A0 C1 G2 T3 = bits 1..2 xor bits 2..3           -> (s1 ^ s2) & 3 per byte
// The four results of a word are gathered with one multiply each (the wanted bits land in the top nibble / byte).
// 26 instructions per word against 45 for the compare-with-four-patterns form this replaces.
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &nl4, uint32_t &ok4, uint32_t &c8) {
    const uint32_t s0 = w, s1 = w >> 1, s2 = w >> 2, s3 = w >> 3, s4 = w >> 4, s5 = w >> 5, s6 = w >> 6, s7 = w >> 7;
    const uint32_t hi0 = lop3<(~SS_LA & ~SS_LB & ~SS_LC) & 0xFF>(s7, s6, s5);              // ~s7 & ~s6 & ~s5
    const uint32_t mid = lop3<(~SS_LA & SS_LB & ~SS_LC) & 0xFF>(s4, s3, s2);               // ~s4 & s3 & ~s2
    const uint32_t low = lop3<(SS_LA & ~SS_LB & SS_LC) & 0xFF>(s1, s0, hi0);               // s1 & ~s0 & hi0
    const uint32_t nl = lop3<(SS_LA & SS_LB & SS_LC) & 0xFF>(mid, low, 0x01010101u);
    const uint32_t acg = lop3<(SS_LA & (SS_LB | ~SS_LC)) & 0xFF>(s0, s1, s2);              // s0 & (s1 | ~s2): A, C, G
    const uint32_t t = lop3<(SS_LA & ~SS_LB & ~SS_LC) & 0xFF>(s2, s1, s0);                 // s2 & ~s1 & ~s0: T
    const uint32_t f = lop3<((SS_LA & SS_LB) | (~SS_LA & SS_LC)) & 0xFF>(s4, t, acg);      // s4 ? t : acg
    const uint32_t g = lop3<(SS_LA & ~SS_LB & ~SS_LC) & 0xFF>(s6, s7, s3);                 // s6 & ~s7 & ~s3
    const uint32_t ok = lop3<(SS_LA & SS_LB & SS_LC) & 0xFF>(f, g, 0x01010101u);
    const uint32_t c = lop3<((SS_LA ^ SS_LB) & SS_LC) & 0xFF>(s1, s2, 0x03030303u);        // (s1 ^ s2) & 3
    nl4 = (nl * 0x10204080u) >> 28;      // byte i's bit 0 -> bit 28 + i
    ok4 = (ok * 0x10204080u) >> 28;
    c8 = (c * 0x01041040u) >> 24;        // byte i's two bits -> bits 24 + 2i
}

struct run_bits { uint32_t nl, ok; uint64_t codes; };

__device__ __forceinline__ run_bits classify_run(const uint8_t *smem_run) {
    const uint4 *r4 = reinterpret_cast<const uint4 *>(smem_run);
    uint4 a = r4[0], b = r4[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    run_bits r; r.nl = 0; r.ok = 0; r.codes = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t nl4, ok4, c8;
        classify4(w[i], nl4, ok4, c8);
        r.nl |= nl4 << (4 * i);
        r.ok |= ok4 << (4 * i);
        r.codes |= (uint64_t)c8 << (8 * i);
    }
    return r;
}

// Bits of a 32-byte run that lie on a FASTQ sequence line (line index % 4 == 1), given the line
// index at the run's first byte and the run's newline mask.  Also checks the 4-line framing
// ('@' opens line 0, '+' opens line 2) and counts sequence lines that end in this run.
__device__ __forceinline__ uint32_t seq_line_mask(uint32_t nlmask, uint32_t line, const uint8_t *raw_tile,
                                                  uint32_t run_off, uint64_t tile_gpos, uint64_t text_len,
                                                  bool check, uint32_t &n_reads, unsigned long long *err) {
    uint32_t seqmask = 0, start = 0, rem = nlmask;
    while (true) {
        uint32_t end = rem ? (uint32_t)(__ffs(rem) - 1) : 32u;
        if ((line & 3u) == 1u) {
            uint32_t hi = (end >= 32u) ? 0xFFFFFFFFu : ((1u << end) - 1u);
            uint32_t lo = (1u << start) - 1u;   // start <= 31 here
            seqmask |= hi & ~lo;
        }
        if (!rem) break;
        rem &= rem - 1;
        if (check) {
            uint64_t gp = tile_gpos + run_off + end + 1;    // first byte of the next line
            if ((line & 3u) == 1u && gp <= text_len) n_reads++;   // not the '\n' padding after the text
            uint32_t nxt = (line + 1u) & 3u;
            if (gp < text_len && (nxt == 0u || nxt == 2u)) {
                uint8_t c = raw_tile[run_off + end + 1];    // inside tile + halo
                if (c != (nxt == 0u ? '@' : '+')) atomicMin(err, (unsigned long long)gp);
            }
        }
        line++;
        start = end + 1;
        if (start >= 32u) break;
    }
    return seqmask;
}

// ---------------------------------------------------------------------------------------------
// K1: line index at every unit start, ONE pass (newline count per 992-byte unit + chained scan).
//
// A CTA takes a ticket (dynamic block number, so that a block only ever waits for blocks that already
// run), counts the newlines of its SS_IDX_UNITS consecutive units (warp per unit, 16-byte loads, SWAR
// compare), scans the counts in shared memory, publishes its aggregate, finds its exclusive prefix by
// decoupled look-back over the predecessors' {flag, value} words (one 64-bit store / volatile load
// each) and writes sub_line[u] = base + #newlines before unit u.  Only the low two bits of a line
// index are ever used, so 32-bit wrap-around is harmless.  The first version ran the scan as a second
// kernel on a single CTA: 2.96 ms for 3.3 M units against 0.55 ms for the counting pass itself.
// ---------------------------------------------------------------------------------------------
#define SS_IDX_UNITS 512u          // units per CTA (508 KB of text)
#define SS_IDX_THREADS 256

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// state[0 .. n_blocks]: entry b + 1 belongs to block b (flag << 32 | value; flag 1 = aggregate, 2 = inclusive
// prefix); zeroed before the launch together with *ticket
__global__ void __launch_bounds__(SS_IDX_THREADS) ss_line_index_kernel(const uint8_t *__restrict__ text, uint32_t n_sub,
                                                                       uint32_t *__restrict__ sub_line, uint32_t base,
                                                                       unsigned long long *__restrict__ state,
                                                                       uint32_t *__restrict__ ticket) {
    __shared__ uint32_t s_cnt[SS_IDX_UNITS];
    __shared__ uint32_t s_wsum[SS_IDX_THREADS / 32];
    __shared__ uint32_t s_blk, s_prefix;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t blk = s_blk;
    const uint32_t u0 = blk * SS_IDX_UNITS;
    for (uint32_t i = wid; i < SS_IDX_UNITS; i += SS_IDX_THREADS / 32) {
        const uint32_t u = u0 + i;
        uint32_t c = 0;
        if (u < n_sub) {
            const uint4 *p = reinterpret_cast<const uint4 *>(text + (uint64_t)u * SS_SUB);
#pragma unroll
            for (int h = 0; h < 2; h++) {   // 62 x 16 B per unit: lane takes chunks lane and lane + 32
                if (lane + h * 32 >= SS_SUB / 16) break;
                uint4 v = __ldg(p + lane + h * 32);
                c += __popc(~nz7(v.x, 0x0A0A0A0Au) & 0x80808080u) + __popc(~nz7(v.y, 0x0A0A0A0Au) & 0x80808080u) +
                     __popc(~nz7(v.z, 0x0A0A0A0Au) & 0x80808080u) + __popc(~nz7(v.w, 0x0A0A0A0Au) & 0x80808080u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        }
        if (lane == 0) s_cnt[i] = c;
    }
    __syncthreads();
    // block scan: thread t owns entries 2t and 2t + 1
    const uint32_t a = s_cnt[2 * threadIdx.x], b = s_cnt[2 * threadIdx.x + 1];
    uint32_t inc = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += x;
    }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < SS_IDX_THREADS / 32; w++) {
        const uint32_t x = s_wsum[w];
        if (w < wid) wbase += x;
        total += x;
    }
    if (wid == 0) {
        uint32_t excl = 0;
        if (blk == 0) {
            excl = base;
        } else {
            if (lane == 0) st_state(state + blk + 1, (1ull << 32) | total);
            int j = (int)blk;                                       // state index of my nearest predecessor
            while (true) {
                const int idx = j - (int)lane;
                unsigned long long v;
                do {                                                // all 32 predecessors have at least an aggregate
                    v = idx >= 1 ? ld_state(state + idx) : (2ull << 32);
                } while (__any_sync(0xFFFFFFFFu, (v >> 32) == 0ull));
                const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, (v >> 32) == 2ull);
                const uint32_t first = m2 ? (uint32_t)(__ffs(m2) - 1) : 32u;
                uint32_t x = lane <= first ? (uint32_t)v : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
                excl += x;
                if (m2) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st_state(state + blk + 1, (2ull << 32) | (uint32_t)(excl + total));
            s_prefix = excl;
        }
    }
    __syncthreads();
    const uint32_t ex = s_prefix + wbase + inc - (a + b);
    const uint32_t u = u0 + 2 * threadIdx.x;
    if (u + 1 < n_sub) *reinterpret_cast<uint2 *>(sub_line + u) = make_uint2(ex, ex + a);
    else if (u < n_sub) sub_line[u] = ex;
}

// ---------------------------------------------------------------------------------------------
// K3: fused scan / encode / filter / probe / count
//
// Every WARP is an autonomous pipeline over 992-byte text units (31 runs of 32 bytes + one halo run
// = one 1024-byte bulk TMA copy): it prefetches its own units into its own shared-memory ring
// (lane 0 issues cp.async.bulk, completion on the warp's own mbarriers), classifies the 32 runs in
// ONE pass (lane = run), keeps the 2-bit codes / valid bits in its own slice, and probes its own
// window starts.  No CTA-wide barrier, no cross-warp coupling: warps never wait for each other.
// ---------------------------------------------------------------------------------------------
// bit L of the result = bytes p+L .. p+L+k-1 are all valid, for V = valid bits of bytes p .. p+63
__device__ __forceinline__ uint32_t window_mask(uint64_t V, int k) {
    uint64_t R = ~0ull, X = V;
    int shift = 0;
#pragma unroll
    for (int b = 0; b < 6; b++) {
        if (k & (1 << b)) { R &= X >> shift; shift += (1 << b); }
        X &= X >> (1 << b);
    }
    return (uint32_t)R;
}

struct __align__(16) ss_warp_smem {
    uint8_t raw[SS_STAGES][SS_TILE + SS_HALO];   // 1024-byte units
    uint32_t codes[2 * SS_WRUNS + 4];            // 2-bit codes of my 32 runs, 2 words per run (+ zero pad)
    uint32_t valid[SS_WRUNS + 2];                // valid bits of my 32 runs (+ zero pad)
    uint64_t q_key[SS_QCAP];                     // deferred table probes: the k-mers that passed the filter
    uint2 g_ent[32 + 8];                         // non-empty groups of the unit: {window mask, byte offset of its codes}
    uint64_t full[SS_STAGES];                    // my mbarriers
    uint32_t qn, pad_;                           // queued survivors
};

// exact table probe of one k-mer (one 32-byte sector, next sector only when the bucket is full)
__device__ __forceinline__ void table_probe(const ss_table_view &tv, uint64_t km, uint32_t &n_hits, uint32_t &n_second) {
    uint32_t hh, hl;
    ss_hash2((uint32_t)km, (uint32_t)(km >> 32), hh, hl);
    uint64_t b = __umulhi(hh, (uint32_t)tv.n_buckets);
    unsigned long long a0, a1, a2, a3;
    ld_bucket(tv.buckets + b, a0, a1, a2, a3);
    while (true) {
        int f = (a0 == km) ? 0 : (a1 == km) ? 1 : (a2 == km) ? 2 : (a3 == km) ? 3 : -1;
        if (f >= 0) {
            red_add_u32(tv.slot_cnt + 4 * b + f, 1u);
            n_hits++;
            return;
        }
        if (a3 == SS_EMPTY || a2 == SS_EMPTY || a1 == SS_EMPTY || a0 == SS_EMPTY) return;   // miss
        b = (b + 1 == tv.n_buckets) ? 0 : b + 1;   // bucket full: next sector
        n_second++;
        ld_bucket(tv.buckets + b, a0, a1, a2, a3);
    }
}

// binned mode: append one surviving k-mer per active lane to its bin.  Every warp owns one open chunk per bin:
// b_cnt[p] = keys already in it (SS_BIN_CHUNK = full / none open yet), b_base[p] = pool index of its first key.
// Common case: one shared-memory atomicAdd hands out the place and the lane stores its key.  The lanes that find
// their chunk full take a new one from the bin's pool (one global atomicAdd per SS_BIN_CHUNK keys).
__device__ __forceinline__ void bin_emit(uint32_t *b_cnt, uint32_t *b_base, const ss_bin_view &bv, bool active,
                                         uint64_t km, uint32_t hh, uint32_t lane, uint32_t lt_mask) {
    uint32_t p = 0, old = 0;
    if (active) {
        p = hh >> bv.shift;
        old = atomicAdd(b_cnt + p, 1u);
        if (old < SS_BIN_CHUNK) bv.keys[b_base[p] + old] = km;
    }
    __syncwarp();
    uint32_t ov = __ballot_sync(0xFFFFFFFFu, active && old >= SS_BIN_CHUNK);
    while (ov) {                                            // rare: once per SS_BIN_CHUNK keys of a bin
        const int l = __ffs(ov) - 1;
        const uint32_t pb = __shfl_sync(0xFFFFFFFFu, p, l);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, active && old >= SS_BIN_CHUNK && p == pb);
        if ((int)lane == l) {
            uint32_t c = atomicAdd(bv.n_chunks + pb, 1u);
            if (c >= bv.cap) { c = bv.cap - 1u; bv.n_chunks[SS_BIN_PMAX] = 1u; }   // flagged: the host redoes the round
            c += pb * bv.cap;
            bv.fill[c] = SS_BIN_CHUNK;                      // closed chunks are full; my last one is fixed up at the end
            b_base[pb] = c * SS_BIN_CHUNK;
            b_cnt[pb] = __popc(m);
        }
        __syncwarp();
        if ((m >> lane) & 1u) bv.keys[b_base[pb] + __popc(m & lt_mask)] = km;
        ov &= ~m;
    }
    __syncwarp();
}

template <bool FILTER, bool K32, int UNROLL, int MINCTAS, bool BIN>
__global__ void __launch_bounds__(SS_CTA_THREADS, MINCTAS)
ss_probe_kernel(const uint8_t *__restrict__ text, uint64_t text_len, uint32_t unit_lo, uint32_t n_units,
                const uint32_t *__restrict__ sub_line, ss_table_view tv, ss_bin_view bv,
                unsigned long long *__restrict__ stats, unsigned long long *__restrict__ err) {
    __shared__ uint32_t s_bcnt[BIN ? SS_CTA_WARPS * SS_BIN_PMAX : 1], s_bbase[BIN ? SS_CTA_WARPS * SS_BIN_PMAX : 1];
    static_assert(32 + 32 * UNROLL <= SS_QCAP, "survivor queue too small for this UNROLL");
    __shared__ ss_warp_smem s_warp[SS_CTA_WARPS];
    __shared__ ss_fword s_pat[FILTER ? SS_NPAT : 1];     // filter bit patterns (4 bits of a filter word)
    if (FILTER) {
        for (uint32_t i = threadIdx.x; i < SS_NPAT; i += SS_CTA_THREADS) s_pat[i] = ss_filter_pattern(i);
        __syncthreads();
    }

    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * SS_CTA_WARPS + wid + unit_lo;   // my first unit
    const uint64_t nw = (uint64_t)gridDim.x * SS_CTA_WARPS;            // unit stride
    constexpr uint32_t kBytes = SS_TILE + SS_HALO;
    ss_warp_smem &ws = s_warp[wid];

    uint64_t pol_stream = 0, pol_keep = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    if (FILTER) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SS_STAGES; s++) mbar_init(&ws.full[s], 1);
        fence_mbar_init();
#pragma unroll
        for (int s = 0; s < SS_STAGES; s++) {
            uint64_t u = gw + (uint64_t)s * nw;
            if (u < n_units) {
                mbar_expect_tx(&ws.full[s], kBytes);
                tma_load_1d_hint(ws.raw[s], text + u * SS_TILE, kBytes, &ws.full[s], pol_stream);
            }
        }
    }
    // zero pad behind the 32 runs (read by the last groups' funnel shifts / window masks)
    if (lane < 4) ws.codes[2 * SS_WRUNS + lane] = 0;
    if (lane < 2) ws.valid[SS_WRUNS + lane] = 0;
    if (lane == 0) ws.qn = 0;
    uint32_t *b_cnt = s_bcnt + (BIN ? wid * SS_BIN_PMAX : 0), *b_base = s_bbase + (BIN ? wid * SS_BIN_PMAX : 0);
    if (BIN) {
        b_cnt[lane] = SS_BIN_CHUNK; b_cnt[lane + 32] = SS_BIN_CHUNK;          // no chunk open yet
        b_base[lane] = 0xFFFFFFFFu; b_base[lane + 32] = 0xFFFFFFFFu;
    }
    __syncwarp();

    uint32_t n_kmers = 0, n_hits = 0, n_second = 0, n_reads = 0, n_table = 0, n_ones = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t fsh = (2u * lane) & 31u;             // funnel shift inside the 32-bit word my window starts in
    const char *const codes_lane = reinterpret_cast<const char *>(ws.codes) + 4u * (lane >> 4);   // ... which is word lane / 16 of the group
    const uint32_t khi_mask = (uint32_t)(tv.kmask >> 32), klo_mask = (uint32_t)tv.kmask;
    const uint32_t phi_mask = (uint32_t)(tv.kmask >> 34), plo_mask = (uint32_t)(tv.kmask >> 2);   // the first k-1 bases

    uint32_t it = 0;
    for (uint64_t unit = gw; unit < n_units; unit += nw, ++it) {
        const uint32_t s = it % SS_STAGES;
        mbar_wait(&ws.full[s], (it / SS_STAGES) & 1u);
        const uint8_t *rt = ws.raw[s];

        // ---- phase 1: lane = run (31 own runs + the halo run), one pass
        {
            run_bits rb = classify_run(rt + lane * SS_RUN);
            uint32_t c = __popc(rb.nl), inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (uint32_t)o) inc += x;
            }
            uint32_t line = sub_line[unit] + inc - c;
            uint32_t seqm = seq_line_mask(rb.nl, line, rt, lane * SS_RUN, unit * SS_TILE, text_len, lane < 31u, n_reads, err);
            ws.codes[2 * lane] = (uint32_t)rb.codes;
            ws.codes[2 * lane + 1] = (uint32_t)(rb.codes >> 32);
            ws.valid[lane] = rb.ok & seqm;
        }
        __syncwarp();
        if (lane == 0) {   // the raw stage is consumed: prefetch the unit SS_STAGES rounds ahead into it
            uint64_t nu = unit + (uint64_t)SS_STAGES * nw;
            if (nu < n_units) {
                fence_proxy_async();
                mbar_expect_tx(&ws.full[s], kBytes);
                tma_load_1d_hint(ws.raw[s], text + nu * SS_TILE, kBytes, &ws.full[s], pol_stream);
            }
        }

        // ---- phase 2: 31 groups of 32 window starts; lane j derives the window mask of group j and
        // the non-empty groups (header / '+' / quality text has none) are listed in shared memory
        uint32_t n_groups;
        {
            uint32_t my_w = 0;
            if (lane < 31u) {
                uint64_t V = (uint64_t)ws.valid[lane] | ((uint64_t)ws.valid[lane + 1] << 32);
                my_w = window_mask(V, tv.k);
                n_kmers += __popc(my_w);
            }
            uint32_t nonempty = __ballot_sync(0xFFFFFFFFu, my_w != 0u);
            n_groups = __popc(nonempty);
            if (my_w) ws.g_ent[__popc(nonempty & lt_mask)] = make_uint2(my_w, 8u * lane);
            if (lane < UNROLL) ws.g_ent[n_groups + lane] = make_uint2(0u, 0u);
            __syncwarp();
        }
        for (uint32_t g = 0; g < n_groups; g += UNROLL) {
            uint32_t k0[UNROLL], k1[UNROLL], hh[UNROLL], hl[UNROLL];
            bool ok[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const uint2 ge = ws.g_ent[g + u];
                ok[u] = (ge.x >> lane) & 1u;
                // my window = 2k bits starting at bit 2*lane of the group's codes: three 32-bit words
                const uint32_t *cw = reinterpret_cast<const uint32_t *>(codes_lane + ge.y);
                uint32_t w0 = cw[0], w1 = cw[1], w2 = cw[2];
                k0[u] = __funnelshift_r(w0, w1, fsh) & klo_mask;
                k1[u] = __funnelshift_r(w1, w2, fsh) & khi_mask;
                if (K32) {
                    if (ok[u] && (k0[u] & k1[u]) == 0xFFFFFFFFu) {   // poly-T 32-mer: lives outside the table
                        if (tv.has_ones) {
                            if (BIN) n_ones++;          // applied by the bin-probe kernel once the round stands
                            else { red_add_u32(tv.slot_cnt + 4 * tv.n_buckets, 1u); n_hits++; }
                        }
                        ok[u] = false;
                    }
                }
                if (FILTER && SS_FILTER_PAIRS) ss_hash2(k0[u] & plo_mask, k1[u] & phi_mask, hh[u], hl[u]);   // my (k-1)-mer
                else ss_hash2(k0[u], k1[u], hh[u], hl[u]);
            }
            if (FILTER) {
                ss_fword fw[UNROLL];
                bool need[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    if (SS_FILTER_PAIRS) {
                        // even lanes ask for their own k-mer and for the one before them; lane 31 asks for itself
                        const uint32_t okm = __ballot_sync(0xFFFFFFFFu, ok[u]);
                        need[u] = (lane & 1u) ? (lane == 31u && ok[u]) : (((okm | (okm << 1)) >> lane) & 1u);
                    } else {
                        need[u] = ok[u];
                    }
                    fw[u] = 0;
                    if (need[u]) fw[u] = ld_filter(tv.filter + __umulhi(hl[u], tv.n_filter_words), pol_keep);
                }
                bool pass[UNROLL], any = false;
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    // looked up by every lane, wanted or not: a pattern is never 0, so fw == 0 (no load) cannot pass
                    const ss_fword m = *reinterpret_cast<const ss_fword *>(
                        reinterpret_cast<const char *>(s_pat) + (hh[u] & ((SS_NPAT - 1u) * (uint32_t)sizeof(ss_fword))));
                    bool hit = (fw[u] & m) == m;
                    if (SS_FILTER_PAIRS) {
                        const bool next = __shfl_down_sync(0xFFFFFFFFu, hit, 1);
                        hit = ((lane & 1u) && lane != 31u) ? next : hit;
                    }
                    pass[u] = ok[u] && hit;
                    any |= pass[u];
                }
                if (BIN) {                              // survivors are the rule here: no queue, straight to the bins
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        n_table += pass[u];
                        uint32_t bh = hh[u];
                        if (SS_FILTER_PAIRS) { uint32_t t; ss_hash2(k0[u], k1[u], bh, t); }   // the bin is the k-mer's bucket range
                        bin_emit(b_cnt, b_base, bv, pass[u], (uint64_t)k0[u] | ((uint64_t)k1[u] << 32), bh, lane, lt_mask);
                    }
                } else if (any) {                       // rare per lane: queue survivors for an exact table probe
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        if (pass[u]) {
                            uint32_t p = atomicAdd(&ws.qn, 1u);
                            ws.q_key[p] = (uint64_t)k0[u] | ((uint64_t)k1[u] << 32);   // re-hashed at probe time
                        }
                    }
                }
                __syncwarp();
                uint32_t qn = BIN ? 0u : ws.qn;         // <= 31 + 32 * UNROLL
                while (qn >= 32u) {                     // probe the table 32 survivors at a time
                    qn -= 32u;
                    n_table++;
                    table_probe(tv, ws.q_key[qn + lane], n_hits, n_second);
                    __syncwarp();
                    if (lane == 0) ws.qn = qn;
                    __syncwarp();
                }
            } else if (BIN) {                           // most k-mers are in the set: the filter is skipped
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    n_table += ok[u];
                    bin_emit(b_cnt, b_base, bv, ok[u], (uint64_t)k0[u] | ((uint64_t)k1[u] << 32), hh[u], lane, lt_mask);
                }
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    if (!ok[u]) continue;
                    table_probe(tv, (uint64_t)k0[u] | ((uint64_t)k1[u] << 32), n_hits, n_second);
                }
            }
        }
        __syncwarp();   // my codes/valid are rewritten by my next unit's phase 1
    }
    __syncwarp();
    if (BIN) {
        for (uint32_t p = lane; p < bv.P; p += 32u)       // my open chunks are partly filled
            if (b_base[p] != 0xFFFFFFFFu) bv.fill[b_base[p] / SS_BIN_CHUNK] = b_cnt[p];
    } else if (FILTER) {       // drain the queue
        if (lane < ws.qn) {
            n_table++;
            table_probe(tv, ws.q_key[lane], n_hits, n_second);
        }
    }

    // ---- statistics
    unsigned long long st[5] = {n_kmers, n_hits, n_second, n_reads, n_table};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long x = st[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0 && x) atomicAdd(stats + (i < 4 ? i : 5), x);
    }
    if (BIN && K32) {
        unsigned long long x = n_ones;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0 && x) atomicAdd(stats + 6, x);
    }
}

// ---------------------------------------------------------------------------------------------
// K3c: probe the binned k-mers, bin after bin (CTAs are numbered bin-major, so the CTAs in flight work on
// one or two neighbouring bins and the table range + counters they touch stay in L2).  One warp per
// chunk: 4 keys per lane, their 4 sector loads in flight together.
// ---------------------------------------------------------------------------------------------
struct ss_bin_sched { uint32_t start[SS_BIN_PMAX + 1]; };   // first chunk number of each bin (prefix sums), host-built

__device__ __forceinline__ unsigned long long ld_key_stream(const unsigned long long *p, uint64_t policy) {
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}

// Persistent: warp w takes chunk numbers w, w + W, w + 2W, ... of the bin-major numbering, so all warps in
// flight work on the same one or two bins.  The keys of the next chunk are requested before the current one
// is probed.
__global__ void __launch_bounds__(SS_CTA_THREADS)
ss_bin_probe_kernel(ss_bin_view bv, ss_bin_sched sched, ss_table_view tv, unsigned long long *__restrict__ stats,
                    const unsigned long long *__restrict__ scan_stats) {
    constexpr int NK = SS_BIN_CHUNK / 32;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && tv.has_ones && scan_stats[6]) {   // the scan's poly-T 32-mers (side slot)
        red_add_u32(tv.slot_cnt + 4 * tv.n_buckets, (uint32_t)scan_stats[6]);
        atomicAdd(stats + 1, scan_stats[6]);
    }
    const uint32_t total = sched.start[SS_BIN_PMAX];
    const uint32_t W = gridDim.x * SS_CTA_WARPS;
    uint32_t n_hits = 0, n_second = 0;
    uint64_t pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    uint32_t p = 0, n_next = 0;
    unsigned long long kn[NK];
    auto fetch = [&](uint32_t w) {          // request chunk number w's keys (bin p only moves forward)
        while (sched.start[p + 1] <= w) p++;
        const uint32_t c = p * bv.cap + (w - sched.start[p]);
        n_next = bv.fill[c];
        const unsigned long long *kp = bv.keys + (uint64_t)c * SS_BIN_CHUNK;
#pragma unroll
        for (int j = 0; j < NK; j++) kn[j] = (j * 32u + lane < n_next) ? ld_key_stream(kp + j * 32 + lane, pol_stream) : 0ull;
    };
    uint32_t w = blockIdx.x * SS_CTA_WARPS + wid;
    if (w < total) fetch(w);
    for (; w < total; w += W) {
        unsigned long long km[NK], a[NK][4];
        uint64_t b[NK];
        const uint32_t n = n_next;
#pragma unroll
        for (int j = 0; j < NK; j++) km[j] = kn[j];
#pragma unroll
        for (int j = 0; j < NK; j++) {
            uint32_t hh, hl;
            ss_hash2((uint32_t)km[j], (uint32_t)(km[j] >> 32), hh, hl);
            b[j] = __umulhi(hh, (uint32_t)tv.n_buckets);
            if (j * 32u + lane < n) ld_bucket(tv.buckets + b[j], a[j][0], a[j][1], a[j][2], a[j][3]);
        }
        if (w + W < total) fetch(w + W);
#pragma unroll
        for (int j = 0; j < NK; j++) {
            if (j * 32u + lane >= n) continue;
            unsigned long long a0 = a[j][0], a1 = a[j][1], a2 = a[j][2], a3 = a[j][3];
            uint64_t bb = b[j];
            while (true) {
                int f = (a0 == km[j]) ? 0 : (a1 == km[j]) ? 1 : (a2 == km[j]) ? 2 : (a3 == km[j]) ? 3 : -1;
                if (f >= 0) { red_add_u32(tv.slot_cnt + 4 * bb + f, 1u); n_hits++; break; }
                if (a3 == SS_EMPTY || a2 == SS_EMPTY || a1 == SS_EMPTY || a0 == SS_EMPTY) break;
                bb = (bb + 1 == tv.n_buckets) ? 0 : bb + 1;
                n_second++;
                ld_bucket(tv.buckets + bb, a0, a1, a2, a3);
            }
        }
    }
    unsigned long long st[2] = {n_hits, n_second};
#pragma unroll
    for (int i = 0; i < 2; i++) {
        unsigned long long x = st[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0 && x) atomicAdd(stats + 1 + i, x);
    }
}

// ---------------------------------------------------------------------------------------------
// K2: table build.  One thread per record; CAS into the first free slot of the home bucket,
// spilling to the next bucket when all four slots are taken.
// ---------------------------------------------------------------------------------------------
__global__ void ss_insert_kernel(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ rec_ok, uint64_t n,
                                 unsigned long long *__restrict__ slots, uint64_t n_buckets,
                                 uint32_t *__restrict__ slot_of, uint32_t *__restrict__ last_ord,
                                 unsigned long long *__restrict__ n_distinct) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!rec_ok[i]) { slot_of[i] = SS_NOSLOT; return; }
    unsigned long long key = keys[i];
    uint64_t slot;
    if (key == SS_EMPTY) {                      // k = 32 poly-T
        slot = 4 * n_buckets;               // counted as distinct by the host
    } else {
        uint32_t hh, hl;
        ss_hash2((uint32_t)key, (uint32_t)(key >> 32), hh, hl);
        uint64_t b = (uint64_t)__umulhi(hh, (uint32_t)n_buckets);
        while (true) {
            bool done = false;
#pragma unroll
            for (int j = 0; j < 4 && !done; j++) {
                unsigned long long cur = slots[4 * b + j];
                if (cur == SS_EMPTY) {
                    unsigned long long old = atomicCAS(slots + 4 * b + j, SS_EMPTY, key);
                    if (old == SS_EMPTY) { atomicAdd(n_distinct, 1ull); cur = key; }
                    else cur = old;
                }
                if (cur == key) { slot = 4 * b + j; done = true; }
            }
            if (done) break;
            b = (b + 1 == n_buckets) ? 0 : b + 1;
        }
    }
    slot_of[i] = (uint32_t)slot;
    atomicMax(last_ord + slot, (uint32_t)i);
}

// L2-resident prefilter: every key sets 4 bits of one 64-bit word
__global__ void ss_filter_build_kernel(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ rec_ok, uint64_t n,
                                       ss_fword *__restrict__ filter, uint32_t n_words, uint64_t kmask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !rec_ok[i] || keys[i] == SS_EMPTY) return;
    uint32_t hh, hl;
#if SS_FILTER_PAIRS
    const uint64_t pre = keys[i] & (kmask >> 2), suf = keys[i] >> 2;      // first / last k-1 bases
    ss_hash2((uint32_t)pre, (uint32_t)(pre >> 32), hh, hl);
    atomicOr(filter + __umulhi(hl, n_words), ss_filter_pattern(ss_pat_index(hh)));
    ss_hash2((uint32_t)suf, (uint32_t)(suf >> 32), hh, hl);
    atomicOr(filter + __umulhi(hl, n_words), ss_filter_pattern(ss_pat_index(hh)));
#else
    ss_hash2((uint32_t)keys[i], (uint32_t)(keys[i] >> 32), hh, hl);
    atomicOr(filter + __umulhi(hl, n_words), ss_filter_pattern(ss_pat_index(hh)));
#endif
}

// flags[i] |= IS_LAST where record i is the highest ordinal stored at its slot
__global__ void ss_flags_kernel(const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ last_ord,
                                uint64_t n, uint8_t *__restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = slot_of[i];
    if (s == SS_NOSLOT) { flags[i] = 0; return; }
    uint8_t f = flags[i] | SS_REC_IN_SET;
    if (last_ord[s] == (uint32_t)i) f |= SS_REC_IS_LAST;
    flags[i] = f;
}

// K3b: dense[i] = slot_cnt[slot_of[i]]
__global__ void ss_gather_kernel(const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ slot_cnt,
                                 uint64_t n, uint32_t *__restrict__ dense) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = slot_of[i];
    dense[i] = (s == SS_NOSLOT) ? 0u : slot_cnt[s];
}

// K3b, streaming form: the slot counters are read in order (16 bytes = one bucket's four counters per load), the
// non-zero ones are written to dense[ordinal of the slot's record] (dense is zeroed beforehand and, at 4 bytes per
// record, mostly lives in L2) and cleared for the next pass.  Against the gather above this reads the counter array
// once as a stream instead of fetching one 128-byte DRAM line per record.
__global__ void __launch_bounds__(256) ss_scatter_kernel(uint32_t *__restrict__ slot_cnt,
                                                         const uint32_t *__restrict__ ord_of_slot, uint64_t n_buckets,
                                                         uint32_t *__restrict__ dense) {
    uint4 *c4 = reinterpret_cast<uint4 *>(slot_cnt);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += stride) {
        uint4 v = c4[b];
        if (v.x | v.y | v.z | v.w) {
            if (v.x) dense[ord_of_slot[4 * b + 0]] = v.x;
            if (v.y) dense[ord_of_slot[4 * b + 1]] = v.y;
            if (v.z) dense[ord_of_slot[4 * b + 2]] = v.z;
            if (v.w) dense[ord_of_slot[4 * b + 3]] = v.w;
            c4[b] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {          // the side slot of the all-ones key (k = 32 poly-T)
        uint32_t v = slot_cnt[4 * n_buckets];
        if (v) { dense[ord_of_slot[4 * n_buckets]] = v; slot_cnt[4 * n_buckets] = 0u; }
    }
}

// records that repeat an earlier/later record's k-mer take the count of the slot's representative (its last record)
__global__ void ss_dup_fill_kernel(const uint32_t *__restrict__ dup, uint64_t n_dup, const uint32_t *__restrict__ slot_of,
                                   const uint32_t *__restrict__ ord_of_slot, uint32_t *__restrict__ dense) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dup) return;
    uint32_t r = dup[i];
    dense[r] = dense[ord_of_slot[slot_of[r]]];
}

// L2 adapter: py_o[kid-1] = (raw_upper && c != 1) ? c : 0   (remove_1)
__global__ void ss_l2_finalize_kernel(const uint32_t *__restrict__ dense, const uint8_t *__restrict__ flags,
                                      const uint32_t *__restrict__ row_of, uint64_t n, long long *__restrict__ py_o) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = (flags[i] & SS_REC_RAW_UPPER) ? dense[i] : 0u;
    if (c == 1u) c = 0u;
    uint64_t r = row_of ? row_of[i] : i;
    py_o[r] = (long long)c;
}

// ---------------------------------------------------------------------------------------------
// K4: per-node reducer.  One warp per node walks its ordinal list.
// The reference intersects a SET of ordinals with valid_kmers, so duplicates inside one node list
// count once; the host passes de-duplicated lists (strainscan_b200/identify_shim.py).
// ---------------------------------------------------------------------------------------------
__global__ void ss_node_reduce_kernel(const uint32_t *__restrict__ dense, const uint8_t *__restrict__ flags,
                                      const unsigned long long *__restrict__ node_ptr,
                                      const uint32_t *__restrict__ ordinals, uint32_t n_nodes, uint64_t n_records,
                                      uint32_t *__restrict__ length, uint32_t *__restrict__ covered,
                                      unsigned long long *__restrict__ sum, uint32_t *__restrict__ max_count) {
    uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (node >= n_nodes) return;
    unsigned long long lo = node_ptr[node], hi = node_ptr[node + 1];
    uint32_t len = 0, cov = 0, mx = 0;
    unsigned long long s = 0;
    for (unsigned long long i = lo + lane; i < hi; i += 32) {
        uint32_t o = ordinals[i];
        if (o >= n_records) continue;
        uint8_t f = flags[o];
        if ((f & (SS_REC_IN_SET | SS_REC_IS_LAST)) == (SS_REC_IN_SET | SS_REC_IS_LAST)) {
            len++;
            uint32_t c = dense[o];
            if (c > 0) { cov++; s += c; mx = max(mx, c); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        len += __shfl_xor_sync(0xFFFFFFFFu, len, o);
        cov += __shfl_xor_sync(0xFFFFFFFFu, cov, o);
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if (lane == 0) { length[node] = len; covered[node] = cov; sum[node] = s; if (max_count) max_count[node] = mx; }
}

// K5: per-strain reducer over the CSC form of the 0/1 strain matrix.  One warp per strain column.
__global__ void ss_strain_reduce_kernel(const unsigned long long *__restrict__ col_ptr,
                                        const uint32_t *__restrict__ rows, uint32_t n_strains,
                                        const long long *__restrict__ y, const uint8_t *__restrict__ row_mask,
                                        unsigned long long *__restrict__ total,
                                        unsigned long long *__restrict__ covered,
                                        unsigned long long *__restrict__ sum) {
    uint32_t col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (col >= n_strains) return;
    unsigned long long lo = col_ptr[col], hi = col_ptr[col + 1];
    unsigned long long t = 0, c = 0, s = 0;
    for (unsigned long long i = lo + lane; i < hi; i += 32) {
        uint32_t r = rows[i];
        if (row_mask && !row_mask[r]) continue;
        t++;
        long long v = y[r];
        if (v > 1) { c++; s += (unsigned long long)v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    }
    if (lane == 0) { total[col] = t; covered[col] = c; sum[col] = s; }
}

// ---------------------------------------------------------------------------------------------
// K0: random 32-byte sector gather (the roofline denominator for K3's probes)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ss_random_gather_kernel(const ss_bucket *__restrict__ buf, uint64_t n_sectors,
                                                                uint64_t n_probes, uint64_t seed,
                                                                unsigned long long *__restrict__ sink) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i + 3 * stride < n_probes; i += 4 * stride) {
        unsigned long long a[4], b[4], c[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint64_t s = __umul64hi(ss_mix(seed + i + u * stride), n_sectors);
            ld_bucket(buf + s, a[u], b[u], c[u], d[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += a[u] ^ b[u] ^ c[u] ^ d[u];
    }
    for (; i < n_probes; i += stride) {
        unsigned long long a, b, c, d;
        ld_bucket(buf + __umul64hi(ss_mix(seed + i), n_sectors), a, b, c, d);
        acc += a ^ b ^ c ^ d;
    }
    if (acc == 0x123456789ull) *sink = acc;   // keep the loads alive
}

// ---------------------------------------------------------------------------------------------
// launch wrappers (host)
// ---------------------------------------------------------------------------------------------
static int g_probe_ctas_per_sm = 0;

#define SS_KERNEL(F, K) ss_probe_kernel<F, K, SS_PROBE_UNROLL, SS_PROBE_MIN_CTAS, false>
#define SS_KERNEL_BIN(F, K) ss_probe_kernel<F, K, SS_PROBE_UNROLL, SS_PROBE_MIN_CTAS, true>

int ss_probe_ctas_per_sm() {
    if (g_probe_ctas_per_sm == 0) {
        int best = 1 << 30;
        const void *fns[4] = {(const void *)SS_KERNEL(true, false), (const void *)SS_KERNEL(true, true),
                              (const void *)SS_KERNEL(false, false), (const void *)SS_KERNEL(false, true)};
        for (int i = 0; i < 4; i++) {
            int n = 0;
            // shared-memory carve-out (percent of 228 KB).  The rest of the 256 KB is L1, whose lines are the
            // landing slots of the in-flight filter loads: a kernel that takes all of it for shared memory
            // loses its memory-level parallelism (measured: 1.6e11 -> 1.0e11 k-mers/s at 100 %).
            if (const char *e = getenv("SS_CARVEOUT")) {
                int pct = atoi(e);
                if (pct >= 0 && pct <= 100) cudaFuncSetAttribute(fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            }
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fns[i], SS_CTA_THREADS, 0) != cudaSuccess || n < 1) n = 1;
            best = n < best ? n : best;
        }
        if (const char *e = getenv("SS_PROBE_CTAS")) { int v = atoi(e); if (v >= 1 && v <= best) best = v; }
        g_probe_ctas_per_sm = best;
    }
    return g_probe_ctas_per_sm;
}

// words of the sub_line allocation: (n_tiles + 1) index entries, then the look-back state of K1
// ((blocks + 1) 64-bit words) and its ticket; the text buffer is padded by one tile
size_t ss_index_words(uint32_t n_tiles) {
    const size_t n_sub = (size_t)n_tiles + 1;
    const size_t blocks = (n_sub + SS_IDX_UNITS - 1) / SS_IDX_UNITS;
    return ((n_sub + 1) & ~(size_t)1) + 2 * (blocks + 1) + 2;
}

cudaError_t ss_launch_index(const uint8_t *text, uint32_t n_tiles, uint32_t *sub_line, uint32_t line_base,
                            int n_sm, cudaStream_t st) {
    if (n_tiles == 0) return cudaSuccess;
    const uint32_t n_sub = n_tiles + 1;
    const uint32_t blocks = (n_sub + SS_IDX_UNITS - 1) / SS_IDX_UNITS;
    uint32_t *scratch = sub_line + ((n_sub + 1u) & ~1u);            // 8-byte aligned: sub_line comes from cudaMalloc
    const size_t scratch_words = 2 * ((size_t)blocks + 1) + 2;
    cudaError_t e = cudaMemsetAsync(scratch, 0, scratch_words * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    (void)n_sm;
    ss_line_index_kernel<<<blocks, SS_IDX_THREADS, 0, st>>>(text, n_sub, sub_line, line_base,
                                                           reinterpret_cast<unsigned long long *>(scratch),
                                                           scratch + 2 * ((size_t)blocks + 1));
    return cudaGetLastError();
}

cudaError_t ss_launch_probe_range(const uint8_t *text, uint64_t text_len, uint32_t tile_lo, uint32_t tile_hi,
                                  const uint32_t *sub_line, const ss_table_view &tv, unsigned long long *stats,
                                  unsigned long long *err, int n_sm, cudaStream_t st) {
    if (tile_hi <= tile_lo) return cudaSuccess;
    uint32_t n = tile_hi - tile_lo;
    uint32_t grid = min((n + SS_CTA_WARPS - 1) / SS_CTA_WARPS, (uint32_t)(n_sm * ss_probe_ctas_per_sm()));
    const bool k32 = tv.k == 32;
    ss_bin_view nb = {};
    if (tv.filter) {
        if (k32) SS_KERNEL(true, true)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, nb, stats, err);
        else SS_KERNEL(true, false)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, nb, stats, err);
    } else {
        if (k32) SS_KERNEL(false, true)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, nb, stats, err);
        else SS_KERNEL(false, false)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, nb, stats, err);
    }
    return cudaGetLastError();
}

cudaError_t ss_launch_probe(const uint8_t *text, uint64_t text_len, uint32_t n_tiles, const uint32_t *sub_line,
                            const ss_table_view &tv, unsigned long long *stats, unsigned long long *err, int n_sm,
                            cudaStream_t st) {
    return ss_launch_probe_range(text, text_len, 0, n_tiles, sub_line, tv, stats, err, n_sm, st);
}

// binned mode, phase 1: scan tiles [tile_lo, tile_hi) and append the k-mers that pass the filter to the bins
int ss_bin_scan_ctas_per_sm() {
    static int n_cached = 0;
    if (n_cached == 0) {
        int best = 1 << 30;
        const void *fns[4] = {(const void *)SS_KERNEL_BIN(true, false), (const void *)SS_KERNEL_BIN(true, true),
                              (const void *)SS_KERNEL_BIN(false, false), (const void *)SS_KERNEL_BIN(false, true)};
        for (int i = 0; i < 4; i++) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fns[i], SS_CTA_THREADS, 0) != cudaSuccess || n < 1) n = 1;
            best = n < best ? n : best;
        }
        n_cached = best;
    }
    return n_cached;
}

cudaError_t ss_launch_bin_scan(const uint8_t *text, uint64_t text_len, uint32_t tile_lo, uint32_t tile_hi,
                               const uint32_t *sub_line, const ss_table_view &tv, const ss_bin_view &bv, bool use_filter,
                               unsigned long long *stats, unsigned long long *err, int n_sm, cudaStream_t st) {
    if (tile_hi <= tile_lo) return cudaSuccess;
    uint32_t n = tile_hi - tile_lo;
    uint32_t grid = min((n + SS_CTA_WARPS - 1) / SS_CTA_WARPS, (uint32_t)(n_sm * ss_bin_scan_ctas_per_sm()));
    const bool k32 = tv.k == 32;
    if (use_filter && tv.filter) {
        if (k32) SS_KERNEL_BIN(true, true)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, bv, stats, err);
        else SS_KERNEL_BIN(true, false)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, bv, stats, err);
    } else {
        if (k32) SS_KERNEL_BIN(false, true)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, bv, stats, err);
        else SS_KERNEL_BIN(false, false)<<<grid, SS_CTA_THREADS, 0, st>>>(text, text_len, tile_lo, tile_hi, sub_line, tv, bv, stats, err);
    }
    return cudaGetLastError();
}

// binned mode, phase 2: n_chunks[p] = chunks handed out in bin p (host copy, no bin overflowed)
cudaError_t ss_launch_bin_probe(const ss_bin_view &bv, const uint32_t *n_chunks, const ss_table_view &tv,
                                unsigned long long *stats, const unsigned long long *scan_stats, int n_sm,
                                cudaStream_t st) {
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ss_bin_probe_kernel, SS_CTA_THREADS, 0) != cudaSuccess || n < 1) n = 1;
        ctas_per_sm = n;
    }
    ss_bin_sched sched;
    uint32_t total = 0;
    for (uint32_t p = 0; p < SS_BIN_PMAX; p++) {
        sched.start[p] = total;
        if (p < bv.P) total += n_chunks[p];
    }
    sched.start[SS_BIN_PMAX] = total;
    if (total == 0 && !tv.has_ones) return cudaSuccess;
    uint32_t grid = min((total + SS_CTA_WARPS - 1) / SS_CTA_WARPS, (uint32_t)(n_sm * ctas_per_sm));
    ss_bin_probe_kernel<<<grid ? grid : 1, SS_CTA_THREADS, 0, st>>>(bv, sched, tv, stats, scan_stats);
    return cudaGetLastError();
}

cudaError_t ss_launch_filter_build(const uint64_t *keys, const uint8_t *rec_ok, uint64_t n, ss_fword *filter,
                                   uint32_t n_words, uint64_t kmask, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ss_filter_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, rec_ok, n, filter, n_words, kmask);
    return cudaGetLastError();
}

cudaError_t ss_launch_insert(const uint64_t *keys, const uint8_t *rec_ok, uint64_t n, unsigned long long *slots,
                             uint64_t n_buckets, uint32_t *slot_of, uint32_t *last_ord,
                             unsigned long long *n_distinct, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ss_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, rec_ok, n, slots, n_buckets, slot_of,
                                                                   last_ord, n_distinct);
    return cudaGetLastError();
}

cudaError_t ss_launch_flags(const uint32_t *slot_of, const uint32_t *last_ord, uint64_t n, uint8_t *flags,
                            cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ss_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slot_of, last_ord, n, flags);
    return cudaGetLastError();
}

cudaError_t ss_launch_gather(const uint32_t *slot_of, const uint32_t *slot_cnt, uint64_t n, uint32_t *dense,
                             cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ss_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slot_of, slot_cnt, n, dense);
    return cudaGetLastError();
}

cudaError_t ss_launch_scatter(uint32_t *slot_cnt, const uint32_t *ord_of_slot, uint64_t n_buckets, const uint32_t *dup,
                              uint64_t n_dup, const uint32_t *slot_of, uint64_t n_records, uint32_t *dense, int n_sm,
                              cudaStream_t st) {
    if (n_records == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dense, 0, n_records * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (n_buckets + 255) / 256;
    unsigned grid = (unsigned)(want < (uint64_t)n_sm * 16 ? want : (uint64_t)n_sm * 16);
    ss_scatter_kernel<<<grid ? grid : 1, 256, 0, st>>>(slot_cnt, ord_of_slot, n_buckets, dense);
    if (n_dup) ss_dup_fill_kernel<<<(unsigned)((n_dup + 255) / 256), 256, 0, st>>>(dup, n_dup, slot_of, ord_of_slot, dense);
    return cudaGetLastError();
}

cudaError_t ss_launch_l2_finalize(const uint32_t *dense, const uint8_t *flags, const uint32_t *row_of, uint64_t n,
                                  long long *py_o, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ss_l2_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dense, flags, row_of, n, py_o);
    return cudaGetLastError();
}

cudaError_t ss_launch_node_reduce(const uint32_t *dense, const uint8_t *flags, const unsigned long long *node_ptr,
                                  const uint32_t *ordinals, uint32_t n_nodes, uint64_t n_records, uint32_t *length,
                                  uint32_t *covered, unsigned long long *sum, uint32_t *max_count, cudaStream_t st) {
    if (n_nodes == 0) return cudaSuccess;
    unsigned blocks = (unsigned)(((uint64_t)n_nodes * 32 + 255) / 256);
    ss_node_reduce_kernel<<<blocks, 256, 0, st>>>(dense, flags, node_ptr, ordinals, n_nodes, n_records, length,
                                                   covered, sum, max_count);
    return cudaGetLastError();
}

cudaError_t ss_launch_strain_reduce(const unsigned long long *col_ptr, const uint32_t *rows, uint32_t n_strains,
                                    const long long *y, const uint8_t *row_mask, unsigned long long *total,
                                    unsigned long long *covered, unsigned long long *sum, cudaStream_t st) {
    if (n_strains == 0) return cudaSuccess;
    unsigned blocks = (unsigned)(((uint64_t)n_strains * 32 + 255) / 256);
    ss_strain_reduce_kernel<<<blocks, 256, 0, st>>>(col_ptr, rows, n_strains, y, row_mask, total, covered, sum);
    return cudaGetLastError();
}

cudaError_t ss_launch_random_gather(const void *buf, uint64_t n_sectors, uint64_t n_probes, uint64_t seed,
                                    unsigned long long *sink, int n_sm, cudaStream_t st) {
    ss_random_gather_kernel<<<n_sm * 8, 256, 0, st>>>((const ss_bucket *)buf, n_sectors, n_probes, seed, sink);
    return cudaGetLastError();
}
