// ss_dgz.cu -- kernels K7..K10 and the host driver of the device inflate of ordinary gzip streams (ss_dgz.cuh).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ss_common.cuh"
#include "ss_dgz.cuh"
#include "ss_dgz2.cuh"
#include "ss_dgz_host.h"

#ifndef SS_DGZ_DECODERS_PER_SM
#define SS_DGZ_DECODERS_PER_SM 24u       // one-lane decoders resident per SM (80 registers each; 7 KB of tables)
#endif

// ---------------------------------------------------------------------------------------------
// K7: find a block start for every piece but the first.  One warp per piece: the lanes try 32 consecutive bit
// positions with the cheap test, the (rare) survivors are put through the full header test one after the other.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) ss_dgz_find_kernel(const uint8_t *__restrict__ comp, size_t comp_size, dgz_piece *pieces,
                                                          uint32_t n_pieces, uint64_t first_byte, uint64_t limit_bit, uint32_t piece) {
    __shared__ ssi_tables s_tab;
    const uint32_t lane = threadIdx.x;
    for (uint32_t j = blockIdx.x + 1; j < n_pieces; j += gridDim.x) {
        const uint64_t from = (first_byte + (uint64_t)j * piece) * 8u;
        uint64_t to = from + (uint64_t)piece * 8u;
        if (to > limit_bit) to = limit_bit;
        uint64_t found = ~0ull;
        for (uint64_t base = from; base < to && found == ~0ull; base += 32u) {
            const uint64_t p = base + lane;
            const bool ok = p < to && dgz_quick_test(comp, comp_size, p);
            uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
            while (m) {
                const int l = __ffs(m) - 1;
                int good = 0;
                if ((int)lane == l) good = dgz_full_test(comp, comp_size, p, s_tab) ? 1 : 0;
                good = __shfl_sync(0xFFFFFFFFu, good, l);
                if (good) { found = base + (uint64_t)l; break; }
                m &= m - 1;
            }
        }
        if (lane == 0) pieces[j].start_bit = found;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K8: marker-mode decode, one decoder (lane 0 of a one-warp CTA, tables in shared memory) per piece, pieces handed
// out dynamically.  Huffman decoding is a serial bit chain: parallelism = pieces in flight.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, SS_DGZ_DECODERS_PER_SM) ss_dgz_decode_kernel(const uint8_t *__restrict__ comp, size_t comp_size, size_t true_size, dgz_piece *pieces,
                                                                uint32_t n_pieces, uint64_t limit_bit, uint64_t stop_byte,
                                                                uint16_t *__restrict__ sym_pool, uint32_t cap,
                                                                unsigned int *__restrict__ next) {
    __shared__ ssi_tables s_tab;
    if (threadIdx.x != 0) return;
    while (true) {
        const uint32_t j = atomicAdd(next, 1u);
        if (j >= n_pieces) break;
        dgz_decode_piece(comp, comp_size, true_size, pieces, n_pieces, j, limit_bit, stop_byte, sym_pool + (uint64_t)j * cap, cap, s_tab);
    }
}

// ---------------------------------------------------------------------------------------------
// K8, several decoders per warp (ss_dgz2.cuh): one CTA per SM, LANES decoders in each of its warps, every decoder's
// 2.5 KB of tables in shared memory.  Between rounds a decoder does whatever its state asks for (fetch a piece, read a
// block header, build tables, end a piece); rounds are entered by all decoders of a warp together.
// ---------------------------------------------------------------------------------------------
#define SS_DGZ2_ROUND 256u
#define SS_DGZ2_SMEM_MAX 232448u
constexpr uint32_t dgz2_smem(uint32_t lanes, uint32_t warps, bool ring) {
    return 256u + warps * lanes * ((uint32_t)sizeof(dgz_ctables) + (ring ? DGZ2_RING * 2u : 0u));
}

template <int LANES, int WARPS, bool LOCKSTEP, bool RING>
__global__ void __launch_bounds__(WARPS * 32, 1) ss_dgz_decode_lanes_kernel(dgz2_job J, unsigned int *__restrict__ next) {
    extern __shared__ __align__(16) uint8_t dgz2_shared[];
    uint32_t *base_tab = reinterpret_cast<uint32_t *>(dgz2_shared);
    if (threadIdx.x < 64u) base_tab[threadIdx.x] = dgc_base_entry(threadIdx.x);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane >= (uint32_t)LANES) return;
    const uint32_t me = warp * LANES + lane;
    dgz_ctables &t = *reinterpret_cast<dgz_ctables *>(dgz2_shared + 256u + me * sizeof(dgz_ctables));
    uint16_t *ring = reinterpret_cast<uint16_t *>(dgz2_shared + 256u + WARPS * LANES * sizeof(dgz_ctables)) + (RING ? me * DGZ2_RING : 0u);
    const uint32_t mask = LANES >= 32 ? 0xFFFFFFFFu : (1u << LANES) - 1u;
    dgz2_lane L;
    L.mode = DGZ2_IDLE;
    while (true) {
        uint32_t budget;
        while (true) {
            budget = dgz2_advance(L, J, t);
            if (L.mode != DGZ2_IDLE) break;
            const uint32_t j = atomicAdd(next, 1u);
            if (j >= J.n_pieces) { L.mode = DGZ2_DONE; break; }
            dgz2_begin_piece(L, J, j);
        }
        if (__all_sync(mask, L.mode == DGZ2_DONE)) break;
        int rc;
        if constexpr (LOCKSTEP) rc = dgz2_round<RING>(L, t, base_tab, ring, budget, J.rounds, [mask](bool a) { return __any_sync(mask, a) != 0; });
        else rc = dgz2_round<RING>(L, t, base_tab, ring, budget, J.rounds, dgz2_vote_alone());        // every lane at its own pace
        dgz2_after_round(L, J, rc);
    }
}

// the shapes K8 is built in: decoders per warp x warps per CTA (one CTA per SM), the decoders of a warp in lockstep
// (one vote per iteration) or each at its own pace inside a round, with or without the ring of recent symbols in
// shared memory; SS_DGZ_LANES / SS_DGZ_WARPS / SS_DGZ_LOCKSTEP / SS_DGZ_RING pick one
static const ss_dgz_shape dgz2_shapes[] = {{2, 24, 1, 0}, {2, 24, 1, 1}, {1, 24, 1, 0}, {3, 24, 1, 0}, {4, 23, 1, 0}, {2, 24, 0, 0}, {3, 24, 0, 0}};

ss_dgz_shape ss_dgz::shape_from_env() {
    int lanes = SS_DGZ_LANES_DEFAULT, warps = 0, lockstep = SS_DGZ_LOCKSTEP_DEFAULT, ring = SS_DGZ_RING_DEFAULT;
    if (const char *e = getenv("SS_DGZ_LANES")) lanes = atoi(e);
    if (const char *e = getenv("SS_DGZ_WARPS")) warps = atoi(e);
    if (const char *e = getenv("SS_DGZ_LOCKSTEP")) lockstep = atoi(e) != 0;
    if (const char *e = getenv("SS_DGZ_RING")) ring = atoi(e) != 0;
    if (lanes == 1) lockstep = 1;
    if (lanes != 2 || !lockstep) ring = 0;                           // (the rings of more than 50 decoders do not fit beside the tables)
    ss_dgz_shape r = {0, 0, 1, 0};
    for (const ss_dgz_shape &sh : dgz2_shapes)
        if (r.lanes == 0 && sh.lanes == lanes && sh.lockstep == lockstep && sh.ring == ring && (warps == 0 || sh.warps == warps)) r = sh;
    return r;                                                        // {0, 0}: one decoder per one-warp CTA (ss_dgz.cuh)
}

uint32_t ss_dgz::decoders_per_sm(ss_dgz_shape sh) { return sh.lanes ? (uint32_t)(sh.lanes * sh.warps) : SS_DGZ_DECODERS_PER_SM; }

template <int LANES, int WARPS, bool LOCKSTEP, bool RING>
static cudaError_t dgz2_launch(const dgz2_job &J, unsigned int *next, int n_sm, cudaStream_t st) {
    static_assert(dgz2_smem(LANES, WARPS, RING) <= SS_DGZ2_SMEM_MAX, "the decode tables (and rings) of a CTA must fit the shared memory of an SM");
    cudaError_t e = cudaFuncSetAttribute(ss_dgz_decode_lanes_kernel<LANES, WARPS, LOCKSTEP, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dgz2_smem(LANES, WARPS, RING));
    if (e != cudaSuccess) return e;
    const uint32_t per_cta = WARPS * LANES;
    const uint32_t grid = std::min<uint32_t>((J.n_pieces + per_cta - 1) / per_cta, (uint32_t)n_sm);
    ss_dgz_decode_lanes_kernel<LANES, WARPS, LOCKSTEP, RING><<<grid, WARPS * 32, dgz2_smem(LANES, WARPS, RING), st>>>(J, next);
    return cudaGetLastError();
}

static cudaError_t dgz2_launch_shape(ss_dgz_shape sh, const dgz2_job &J, unsigned int *next, int n_sm, cudaStream_t st) {
#define DGZ2_CASE(l, w, k, r) if (sh.lanes == l && sh.warps == w && sh.lockstep == k && sh.ring == r) return dgz2_launch<l, w, k != 0, r != 0>(J, next, n_sm, st)
    DGZ2_CASE(2, 24, 1, 0); DGZ2_CASE(2, 24, 1, 1); DGZ2_CASE(1, 24, 1, 0); DGZ2_CASE(3, 24, 1, 0); DGZ2_CASE(4, 23, 1, 0);
    DGZ2_CASE(2, 24, 0, 0); DGZ2_CASE(3, 24, 0, 0);
#undef DGZ2_CASE
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// K9: the window in front of every accepted piece.  windows[k] = the 32 KiB in front of accepted piece k;
// windows[0] is given.  The dependency runs through the whole stream (headers copy headers copy headers ...: the last
// 32 KiB of a piece always hold markers), but it composes: the TAIL MAP of piece k sends every position of the window
// behind it either to a literal or to a position of the window in front of it, and (B after A)[i] = B[i] is a literal
// ? B[i] : A[B[i]].  So the chain is a scan over maps, done in three steps: (a) groups of SS_DGZ_GROUP consecutive
// pieces, one CTA each, compose their maps in order (relative to the group's first window); (b) one CTA composes the
// group totals in order; (c) every piece applies the prefix of the groups before it and looks the rest up in
// windows[0].  Every sequential step is a 32 KiB gather through two dependent loads, written so that the 32 loads of
// a kind a thread owns are in flight together (4.9 us per step; a first form that walked its positions one dependent
// pair at a time took 24 us).  7104 pieces: 32 + 222 sequential steps instead of 7104.
// ---------------------------------------------------------------------------------------------
#define SS_DGZ_GROUP 32u

// positions tid, tid + 1024, ... of the tail map of a piece
__device__ __forceinline__ void dgz_tail_map(const uint16_t *__restrict__ ps, uint32_t n, uint16_t (&x)[32]) {
    const int64_t e0 = (int64_t)n - (int64_t)SS_DGZ_WINDOW + (int64_t)threadIdx.x;
#pragma unroll
    for (uint32_t j = 0; j < 32; j++) {
        const int64_t e = e0 + (int64_t)(j * 1024u);
        x[j] = e >= 0 ? ps[e] : (uint16_t)(256 + (int64_t)SS_DGZ_WINDOW + e);   // still in the old window: points at itself, shifted
    }
}

__global__ void __launch_bounds__(1024) ss_dgz_win_local_kernel(const uint32_t *__restrict__ order, uint32_t n_acc,
                                                                 const dgz_piece *__restrict__ pieces,
                                                                 const uint16_t *__restrict__ sym_pool, uint32_t cap,
                                                                 uint16_t *__restrict__ maps) {
    const uint32_t k0 = blockIdx.x * SS_DGZ_GROUP, k1 = min(n_acc, k0 + SS_DGZ_GROUP);
    for (uint32_t k = k0; k < k1; k++) {
        const uint32_t pi = order[k];
        uint16_t x[32];
        dgz_tail_map(sym_pool + (uint64_t)pi * cap, pieces[pi].n_sym, x);
        uint16_t *__restrict__ out = maps + (uint64_t)k * SS_DGZ_WINDOW;
        if (k > k0) {
            const uint16_t *__restrict__ prev = maps + (uint64_t)(k - 1) * SS_DGZ_WINDOW;
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) if (x[j] >= 256) x[j] = prev[x[j] - 256u];
        }
#pragma unroll
        for (uint32_t j = 0; j < 32; j++) out[j * 1024u + threadIdx.x] = x[j];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) ss_dgz_win_groups_kernel(const uint16_t *__restrict__ maps, uint32_t n_acc,
                                                                  uint16_t *__restrict__ gpre) {
    const uint32_t n_groups = (n_acc + SS_DGZ_GROUP - 1) / SS_DGZ_GROUP;
    for (uint32_t g = 0; g < n_groups; g++) {
        const uint32_t last = min(n_acc, (g + 1) * SS_DGZ_GROUP) - 1u;
        const uint16_t *__restrict__ G = maps + (uint64_t)last * SS_DGZ_WINDOW;
        uint16_t *__restrict__ out = gpre + (uint64_t)g * SS_DGZ_WINDOW;
        uint16_t x[32];
#pragma unroll
        for (uint32_t j = 0; j < 32; j++) x[j] = G[j * 1024u + threadIdx.x];
        if (g > 0) {
            const uint16_t *__restrict__ prev = gpre + (uint64_t)(g - 1) * SS_DGZ_WINDOW;
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) if (x[j] >= 256) x[j] = prev[x[j] - 256u];
        }
#pragma unroll
        for (uint32_t j = 0; j < 32; j++) out[j * 1024u + threadIdx.x] = x[j];
        __syncthreads();
    }
}

// grid.y = accepted piece k (writes windows[k + 1]), grid.x * 256 threads cover the 32768 positions
__global__ void __launch_bounds__(256) ss_dgz_win_apply_kernel(const uint16_t *__restrict__ maps, const uint16_t *__restrict__ gpre,
                                                                uint8_t *__restrict__ windows) {
    const uint32_t k = blockIdx.y, g = k / SS_DGZ_GROUP;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = maps[(uint64_t)k * SS_DGZ_WINDOW + i];
    if (v >= 256u && g > 0) v = gpre[(uint64_t)(g - 1) * SS_DGZ_WINDOW + (v - 256u)];
    if (v >= 256u) v = windows[v - 256u];                     // windows[0]: the window in front of the batch
    windows[(uint64_t)(k + 1) * SS_DGZ_WINDOW + i] = (uint8_t)v;
}

// K10: symbols -> bytes at their place in the text.  grid.y = accepted piece, grid.x strides over its symbols; a
// thread resolves the 8 symbols behind an 8-byte aligned OUTPUT address and writes them with one store.
__global__ void __launch_bounds__(256) ss_dgz_resolve_kernel(const uint32_t *__restrict__ order, const uint64_t *__restrict__ text_off,
                                                              const dgz_piece *__restrict__ pieces,
                                                              const uint16_t *__restrict__ sym_pool, uint32_t cap,
                                                              const uint8_t *__restrict__ windows, const uint32_t *__restrict__ win_len,
                                                              uint8_t *__restrict__ out, unsigned int *__restrict__ err) {
    const uint32_t k = blockIdx.y;
    const uint32_t n = pieces[order[k]].n_sym;
    const uint32_t invalid_below = 256u + SS_DGZ_WINDOW - win_len[k];     // a marker below this reaches in front of the stream start
    const uint16_t *__restrict__ sym = sym_pool + (uint64_t)order[k] * cap;
    const uint8_t *__restrict__ w = windows + (uint64_t)k * SS_DGZ_WINDOW;
    uint8_t *o = out + text_off[k];
    const uint32_t head = min(n, (uint32_t)((8u - (uint32_t)((uintptr_t)o & 7u)) & 7u));      // bytes in front of the first aligned address
    if (blockIdx.x == 0 && threadIdx.x < head) {
        const uint16_t x = sym[threadIdx.x];
        if (x >= 256 && x < invalid_below) atomicMin(err, k);
        o[threadIdx.x] = dgz_resolve1(x, w);
    }
    const uint32_t n8 = (n - head) >> 3;                                                      // whole aligned groups of 8
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += gridDim.x * blockDim.x) {
        const uint16_t *ps = sym + head + 8u * g;
        uint16_t x[8];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 8; j++) { x[j] = ps[j]; bad |= x[j] >= 256 && x[j] < invalid_below; }
        if (bad) atomicMin(err, k);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) lo |= (uint32_t)dgz_resolve1(x[j], w) << (8 * j);
#pragma unroll
        for (int j = 0; j < 4; j++) hi |= (uint32_t)dgz_resolve1(x[4 + j], w) << (8 * j);
        *reinterpret_cast<uint2 *>(o + head + 8u * g) = make_uint2(lo, hi);
    }
    const uint32_t tail0 = head + 8u * n8;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < n && threadIdx.x < 8) {
        const uint16_t x = sym[tail0 + threadIdx.x];
        if (x >= 256 && x < invalid_below) atomicMin(err, k);
        o[tail0 + threadIdx.x] = dgz_resolve1(x, w);
    }
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static double dgz_now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define DGZ_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err_ = std::string("CUDA error in the device gzip inflate: ") + cudaGetErrorString(e_) + " (" #x ")"; return SS_ERR_CUDA; } } while (0)

ss_dgz::~ss_dgz() { close(); }

void ss_dgz::close() {
    cudaFree(d_pieces_); cudaFree(d_sym_); cudaFree(d_windows_); cudaFree(d_order_); cudaFree(d_off_); cudaFree(d_ctr_);
    cudaFree(d_maps_); cudaFree(d_gpre_); cudaFree(d_wlen_);
    d_pieces_ = nullptr; d_sym_ = nullptr; d_windows_ = nullptr; d_order_ = nullptr; d_off_ = nullptr; d_ctr_ = nullptr;
    d_maps_ = nullptr; d_gpre_ = nullptr; d_wlen_ = nullptr;
    alloc_pieces_ = 0; alloc_sym_ = 0;
}

ss_dgz_plan ss_dgz_make_plan(size_t n_up, int n_sm, ss_dgz_shape shape, double ratio) {
    ss_dgz_plan pl;
    pl.ratio = std::min(64.0, std::max(1.0, ratio));
    const uint32_t wave = (uint32_t)std::max(1, n_sm) * ss_dgz::decoders_per_sm(shape);
    pl.scratch = (shape.lanes ? 4096ull : 2048ull) << 20;
    if (const char *e = getenv("SS_DGZ_BATCH_MB")) { long long v = atoll(e); if (v >= 8 && v <= 16384) pl.scratch = (size_t)v << 20; }
    // symbol room: 2.5 x the head's ratio (a piece runs on to the next FOUND start, which may lie well behind the next
    // nominal one, and the head under-estimates), never under the 5 of the first version
    pl.expand = (uint32_t)std::min(48.0, std::max((double)SS_DGZ_EXPAND_DEFAULT, std::ceil(2.5 * pl.ratio)));
    if (const char *e = getenv("SS_DGZ_SYM_PER_BYTE")) { int v = atoi(e); if (v >= 1 && v <= 64) pl.expand = (uint32_t)v; }
    // piece size: a wave's text (1.3 x the head's ratio per compressed byte, at least what 3.5 would give) fits the batch buffer
    const double per_byte = 1.3 * std::max(3.5, pl.ratio);
    const double p_max = std::max(16384.0, std::min<double>(SS_DGZ_PIECE_DEFAULT, std::floor((double)pl.scratch / ((double)wave * per_byte) / 4096.0) * 4096.0));
    const double waves = std::max(1.0, std::ceil((double)n_up / ((double)wave * p_max)));
    double fit = std::ceil((double)n_up / (waves * wave) / 256.0) * 256.0;
    if (fit > p_max) fit = p_max;
    if (fit < std::min(p_max, 32768.0)) fit = std::min(p_max, 32768.0);       // small inputs: fewer pieces than decoders
    pl.piece = (uint32_t)fit;
    if (const char *e = getenv("SS_DGZ_PIECE_BYTES")) { long long v = atoll(e); if (v >= 4096 && v <= (16 << 20)) pl.piece = (uint32_t)v; }
    pl.max_pieces = std::min<uint32_t>(wave + wave / 8u, SS_DGZ_MAX_PIECES);   // (two waves would need a ratio under 2 to fit the batch buffer)
    if (const char *e = getenv("SS_DGZ_MAX_PIECES")) { long long v = atoll(e); if (v >= 2 && v <= (long long)SS_DGZ_MAX_PIECES) pl.max_pieces = (uint32_t)v; }
    pl.max_pieces = std::max(2u, pl.max_pieces);
    pl.piece = std::max(4096u, pl.piece);
    pl.device_bytes = pl.scratch + (size_t)pl.max_pieces * ((size_t)pl.piece * pl.expand * sizeof(uint16_t) + SS_DGZ_WINDOW * 3u + 64u) + (64u << 20);
    return pl;
}

size_t ss_dgz::held_bytes() const {
    return d_pieces_ ? alloc_sym_ * sizeof(uint16_t) + alloc_pieces_ * ((size_t)SS_DGZ_WINDOW * 3u + 64u) : 0;
}

int ss_dgz::open(int n_sm, cudaStream_t st, const uint8_t *d_comp, const uint8_t *h_comp, size_t comp_size, size_t first_member,
                 size_t stop_member_at, const ss_dgz_plan &plan) {
    n_sm_ = n_sm; st_ = st; d_comp_ = d_comp; h_comp_ = h_comp; size_ = comp_size; stop_at_ = std::min(stop_member_at, comp_size);
    max_pieces_ = std::max(2u, std::min(plan.max_pieces, SS_DGZ_MAX_PIECES));
    shape_ = shape_from_env();
    rounds_ = SS_DGZ2_ROUND;
    if (const char *e = getenv("SS_DGZ_ROUND")) { int v = atoi(e); if (v >= 1) rounds_ = (uint32_t)v; }
    piece_ = std::max(4096u, plan.piece);
    cap_ = piece_ * std::max(1u, plan.expand);
    done_ = false; win_len_ = 0; members_ = 0; pieces_used_ = pieces_found_ = batches_ = 0;
    gate_slack_ = SS_DGZ_GATE_SLACK;
    if (const char *e = getenv("SS_DGZ_GATE_SLACK")) { long long v = atoll(e); if (v >= 0) gate_slack_ = (size_t)v; }   // tests: small slack = retries
    ssi_gz_header h;
    if (first_member >= comp_size || ssi_gz_parse_header(h_comp + first_member, h_comp + comp_size, &h) != SSI_OK) {
        err_ = "device gzip inflate: no gzip member at the start of the range";
        return SS_ERR_IO;
    }
    cur_bit_ = (uint64_t)(first_member + h.header_len) * 8u;
    ratio_ = std::max(2.0, 1.3 * plan.ratio); ms_decode_ = ms_resolve_ = ms_windows_ = 0;     // sizes the first batch; the stream corrects it
    const size_t need_sym = (size_t)max_pieces_ * cap_;
    if (d_pieces_ && alloc_pieces_ >= max_pieces_ && alloc_sym_ >= need_sym) {      // the buffers of an earlier stream do: keep them
        DGZ_CUDA(cudaMemsetAsync(d_windows_, 0, SS_DGZ_WINDOW, st_));
        h_pieces_.resize(max_pieces_);
        return SS_OK;
    }
    close();
    DGZ_CUDA(cudaMalloc(&d_pieces_, (size_t)max_pieces_ * sizeof(dgz_piece)));
    DGZ_CUDA(cudaMalloc(&d_sym_, need_sym * sizeof(uint16_t)));
    DGZ_CUDA(cudaMalloc(&d_windows_, ((size_t)max_pieces_ + 1) * SS_DGZ_WINDOW));
    DGZ_CUDA(cudaMalloc(&d_order_, (size_t)max_pieces_ * sizeof(uint32_t)));
    DGZ_CUDA(cudaMalloc(&d_off_, (size_t)max_pieces_ * sizeof(uint64_t)));
    DGZ_CUDA(cudaMalloc(&d_ctr_, 2 * sizeof(unsigned int)));
    DGZ_CUDA(cudaMalloc(&d_maps_, (size_t)max_pieces_ * SS_DGZ_WINDOW * sizeof(uint16_t)));
    DGZ_CUDA(cudaMalloc(&d_gpre_, ((size_t)max_pieces_ / SS_DGZ_GROUP + 1) * SS_DGZ_WINDOW * sizeof(uint16_t)));
    DGZ_CUDA(cudaMalloc(&d_wlen_, (size_t)max_pieces_ * sizeof(uint32_t)));
    DGZ_CUDA(cudaMemsetAsync(d_windows_, 0, SS_DGZ_WINDOW, st_));
    alloc_pieces_ = max_pieces_; alloc_sym_ = need_sym;
    h_pieces_.resize(max_pieces_);
    return SS_OK;
}

size_t ss_dgz::batch_text_capacity() const { return (size_t)max_pieces_ * cap_; }

// Inflate the next batch into d_out (out_cap bytes of room; any size of at least one piece's worth works: a batch is
// sized to what the stream has been inflating to, and pieces that do not fit are decoded again next time).  *n_out bytes
// were produced; *done: the stream (or this range's members) is finished.  The text continues the previous batch's.
int ss_dgz::next(uint8_t *d_out, size_t out_cap, size_t *n_out, bool *done) {
    *n_out = 0; *done = done_;
    if (done_) return SS_OK;
    if (out_cap < cap_) { err_ = "device gzip inflate: output buffer smaller than one piece's text"; return SS_ERR_ARG; }
    double t0 = dgz_now_ms();
    const uint64_t first_byte = cur_bit_ >> 3;
    uint64_t span = size_ - first_byte;
    uint32_t P = (uint32_t)std::min<uint64_t>(max_pieces_, (span + piece_ - 1) / piece_);
    P = (uint32_t)std::min<double>((double)P, (double)out_cap / ((double)piece_ * ratio_ * 1.15));
    {   // whole waves of decoders: a last wave with a few pieces costs as much as a full one
        const uint32_t wave = (uint32_t)n_sm_ * decoders_per_sm(shape_);
        if (P > wave && P < (uint32_t)std::min<uint64_t>(max_pieces_, (span + piece_ - 1) / piece_)) P -= P % wave;
    }
    if (P == 0) P = 1;
    const uint64_t limit_bit = std::min<uint64_t>((uint64_t)size_ * 8u, (first_byte + (uint64_t)P * piece_) * 8u);
    bool input_complete = true;
    size_t avail = size_;                                            // the kernels of this batch see the input end here
    if (gate_wait_) {
        const size_t need = (size_t)(limit_bit >> 3) + 1 + gate_slack_;
        if (!gate_wait_(need)) { err_ = "the upload of the compressed bytes failed"; return SS_ERR_IO; }
        input_complete = !gate_complete_ || gate_complete_();
        if (!input_complete) avail = std::min(size_, need);          // bytes behind it may not have arrived: a decoder that
    }                                                                // wants them reports a truncated stream, never garbage
    for (uint32_t j = 0; j < P; j++) { h_pieces_[j] = dgz_piece(); h_pieces_[j].start_bit = j == 0 ? cur_bit_ : ~0ull; }
    DGZ_CUDA(cudaMemcpyAsync(d_pieces_, h_pieces_.data(), (size_t)P * sizeof(dgz_piece), cudaMemcpyHostToDevice, st_));
    DGZ_CUDA(cudaMemsetAsync(d_ctr_, 0, sizeof(unsigned int), st_));
    DGZ_CUDA(cudaMemsetAsync(d_ctr_ + 1, 0xFF, sizeof(unsigned int), st_));
    if (P > 1) ss_dgz_find_kernel<<<std::min<uint32_t>(P - 1, (uint32_t)n_sm_ * 32u), 32, 0, st_>>>(d_comp_, avail, d_pieces_, P, first_byte, limit_bit, piece_);
    if (shape_.lanes) {
        dgz2_job J;
        J.comp = d_comp_; J.comp_size = avail; J.true_size = size_; J.pieces = d_pieces_; J.n_pieces = P;
        J.limit_bit = limit_bit; J.stop_byte = (uint64_t)stop_at_; J.sym_pool = d_sym_; J.cap = cap_; J.rounds = rounds_;
        DGZ_CUDA(dgz2_launch_shape(shape_, J, d_ctr_, n_sm_, st_));
    } else {
        ss_dgz_decode_kernel<<<std::min<uint32_t>(P, (uint32_t)n_sm_ * SS_DGZ_DECODERS_PER_SM), 32, 0, st_>>>(d_comp_, avail, size_, d_pieces_, P, limit_bit, (uint64_t)stop_at_,
                                                                                        d_sym_, cap_, d_ctr_);
        DGZ_CUDA(cudaGetLastError());
    }
    DGZ_CUDA(cudaMemcpyAsync(h_pieces_.data(), d_pieces_, (size_t)P * sizeof(dgz_piece), cudaMemcpyDeviceToHost, st_));
    DGZ_CUDA(cudaStreamSynchronize(st_));
    batches_++;
    ms_decode_ += dgz_now_ms() - t0; t0 = dgz_now_ms();
    // ---- the chain: a piece counts only if the accepted piece before it ended exactly on its start
    std::vector<uint32_t> order, wlen;
    std::vector<uint64_t> off;
    uint64_t total = 0;
    uint32_t cur = 0;
    bool stream_end = false;
    for (uint32_t j = 1; j < P; j++) pieces_found_ += h_pieces_[j].start_bit != ~0ull;
    while (true) {
        const dgz_piece &pc = h_pieces_[cur];
        if (pc.status == SS_DGZ_ERROR || pc.status == SS_DGZ_NOSTART) {
            if (!input_complete && order.empty()) {                  // bytes that had not arrived yet?  once more, with all of them
                if (!gate_wait_(size_)) { err_ = "the upload of the compressed bytes failed"; return SS_ERR_IO; }
                return next(d_out, out_cap, n_out, done);
            }
            if (!input_complete) break;                              // keep what stands; the rest is decoded again by the next batch
            err_ = "invalid compressed data";
            return SS_ERR_IO;
        }
        if (pc.status == SS_DGZ_FULL && pc.n_sym == 0) {
            err_ = "a single deflate block inflates to more than the device decoder's piece buffer";
            return SS_ERR_UNSUPPORTED;
        }
        if (total + pc.n_sym > out_cap) {                           // does not fit: the next batch starts here
            if (order.empty()) { err_ = "device gzip inflate: output buffer smaller than one piece's text"; return SS_ERR_ARG; }
            break;
        }
        order.push_back(cur); off.push_back(total);
        wlen.push_back((uint32_t)std::min<uint64_t>(SS_DGZ_WINDOW, (uint64_t)win_len_ + total));   // window bytes that exist in front of it
        total += pc.n_sym;
        members_ += pc.members;
        cur_bit_ = pc.end_bit;
        if (pc.status == SS_DGZ_END) { stream_end = true; break; }
        if (pc.status == SS_DGZ_FULL || pc.next >= P) break;
        cur = pc.next;
    }
    pieces_used_ += order.size();
    const uint32_t n_acc = (uint32_t)order.size();
    DGZ_CUDA(cudaMemcpyAsync(d_order_, order.data(), n_acc * sizeof(uint32_t), cudaMemcpyHostToDevice, st_));
    DGZ_CUDA(cudaMemcpyAsync(d_off_, off.data(), n_acc * sizeof(uint64_t), cudaMemcpyHostToDevice, st_));
    DGZ_CUDA(cudaMemcpyAsync(d_wlen_, wlen.data(), n_acc * sizeof(uint32_t), cudaMemcpyHostToDevice, st_));
    const uint32_t n_groups = (n_acc + SS_DGZ_GROUP - 1) / SS_DGZ_GROUP;
    ss_dgz_win_local_kernel<<<n_groups, 1024, 0, st_>>>(d_order_, n_acc, d_pieces_, d_sym_, cap_, d_maps_);
    ss_dgz_win_groups_kernel<<<1, 1024, 0, st_>>>(d_maps_, n_acc, d_gpre_);
    ss_dgz_win_apply_kernel<<<dim3(SS_DGZ_WINDOW / 256u, n_acc), 256, 0, st_>>>(d_maps_, d_gpre_, d_windows_);
    if (getenv("SS_DEBUG_TIMING")) { cudaStreamSynchronize(st_); ms_windows_ += dgz_now_ms() - t0; }
    ss_dgz_resolve_kernel<<<dim3(16, n_acc), 256, 0, st_>>>(d_order_, d_off_, d_pieces_, d_sym_, cap_, d_windows_, d_wlen_, d_out, d_ctr_ + 1);
    // the last window becomes the first one of the next batch
    DGZ_CUDA(cudaMemcpyAsync(d_windows_, d_windows_ + (size_t)n_acc * SS_DGZ_WINDOW, SS_DGZ_WINDOW, cudaMemcpyDeviceToDevice, st_));
    unsigned int bad = 0xFFFFFFFFu;
    DGZ_CUDA(cudaMemcpyAsync(&bad, d_ctr_ + 1, sizeof bad, cudaMemcpyDeviceToHost, st_));
    DGZ_CUDA(cudaStreamSynchronize(st_));
    if (bad != 0xFFFFFFFFu) { err_ = "invalid compressed data (a match reaches in front of the stream start)"; return SS_ERR_IO; }
    ms_resolve_ += dgz_now_ms() - t0;
    win_len_ = (uint32_t)std::min<uint64_t>(SS_DGZ_WINDOW, (uint64_t)win_len_ + total);
    {   // what a compressed byte has been inflating to (sizes the next batch)
        const double consumed = (double)((cur_bit_ >> 3) - first_byte);
        if (consumed > 65536.0) ratio_ = std::max(1.0, 0.5 * ratio_ + 0.5 * (double)total / consumed);
    }
    *n_out = (size_t)total;
    if (stream_end) { done_ = true; }
    *done = done_;
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// the same pipeline on the host (no GPU): tests of the algorithms against zlib
// ---------------------------------------------------------------------------------------------
int ss_dgz_host_inflate(const uint8_t *comp, size_t comp_size, size_t first_member, size_t stop_member_at,
                        uint32_t max_pieces, uint32_t piece_bytes, uint32_t sym_per_byte, std::vector<uint8_t> &out,
                        size_t *stopped_at, uint64_t *stats4, std::string &err) {
    out.clear();
    const uint32_t piece = std::max(4096u, piece_bytes), cap = piece * std::max(1u, sym_per_byte);
    max_pieces = std::max(2u, max_pieces);
    const uint64_t stop_byte = std::min(stop_member_at, comp_size);
    ssi_gz_header h;
    if (first_member >= comp_size || ssi_gz_parse_header(comp + first_member, comp + comp_size, &h) != SSI_OK) { err = "no gzip member at the start of the range"; return SS_ERR_IO; }
    // the device buffers are padded by 16 readable bytes; so is this copy
    std::vector<uint8_t> padded(comp_size + 32, 0);
    memcpy(padded.data(), comp, comp_size);
    comp = padded.data();
    uint64_t cur_bit = (uint64_t)(first_member + h.header_len) * 8u;
    std::vector<uint8_t> win(SS_DGZ_WINDOW, 0), win2(SS_DGZ_WINDOW, 0);
    uint32_t win_len = 0;
    std::vector<dgz_piece> pieces(max_pieces);
    std::vector<uint16_t> sym((size_t)max_pieces * cap);
    ssi_tables *tab = new ssi_tables;
    const int lanes = ss_dgz::shape_from_env().lanes;
    bool with_ring = SS_DGZ_RING_DEFAULT != 0;                        // (the CPU form has room for a ring whatever the lane count)
    if (const char *e = getenv("SS_DGZ_RING")) with_ring = atoi(e) != 0;
    uint32_t rounds = SS_DGZ2_ROUND;
    if (const char *e = getenv("SS_DGZ_ROUND")) { int v = atoi(e); if (v >= 1) rounds = (uint32_t)v; }    // tests: short rounds
    uint64_t found = 0, used = 0, batches = 0, members = 0;
    int rc = SS_OK;
    while (true) {
        const uint64_t first_byte = cur_bit >> 3;
        uint32_t P = (uint32_t)std::min<uint64_t>(max_pieces, (comp_size - first_byte + piece - 1) / piece);
        if (P == 0) P = 1;
        const uint64_t limit_bit = std::min<uint64_t>((uint64_t)comp_size * 8u, (first_byte + (uint64_t)P * piece) * 8u);
        for (uint32_t j = 0; j < P; j++) { pieces[j] = dgz_piece(); pieces[j].start_bit = ~0ull; }
        pieces[0].start_bit = cur_bit;
        for (uint32_t j = 1; j < P; j++) {
            const uint64_t from = (first_byte + (uint64_t)j * piece) * 8u, to = std::min<uint64_t>(from + (uint64_t)piece * 8u, limit_bit);
            for (uint64_t p = from; p < to; p++)
                if (dgz_quick_test(comp, comp_size, p) && dgz_full_test(comp, comp_size, p, *tab)) { pieces[j].start_bit = p; found++; break; }
        }
        if (lanes) {                                                  // the decoder of ss_dgz2.cuh (any lane count: the CPU runs one)
            dgz2_job J;
            J.comp = comp; J.comp_size = comp_size; J.true_size = comp_size; J.pieces = pieces.data(); J.n_pieces = P;
            J.limit_bit = limit_bit; J.stop_byte = stop_byte; J.sym_pool = sym.data(); J.cap = cap; J.rounds = rounds;
            dgz2_decode_pieces_host(J, with_ring);
        } else
            for (uint32_t j = 0; j < P; j++) dgz_decode_piece(comp, comp_size, comp_size, pieces.data(), P, j, limit_bit, stop_byte, sym.data() + (size_t)j * cap, cap, *tab);
        batches++;
        uint32_t cur = 0;
        bool end = false;
        while (true) {
            const dgz_piece &pc = pieces[cur];
            if (pc.status == SS_DGZ_ERROR || pc.status == SS_DGZ_NOSTART) { err = "invalid compressed data"; rc = SS_ERR_IO; break; }
            if (pc.status == SS_DGZ_FULL && pc.n_sym == 0) { err = "a single deflate block inflates to more than the decoder's piece buffer"; rc = SS_ERR_UNSUPPORTED; break; }
            const uint16_t *ps = sym.data() + (size_t)cur * cap;
            if (!dgz_next_window_part(win.data(), win_len, ps, pc.n_sym, win2.data(), 0, SS_DGZ_WINDOW)) { err = "invalid compressed data (a match reaches in front of the stream start)"; rc = SS_ERR_IO; break; }
            const size_t o = out.size();
            out.resize(o + pc.n_sym);
            for (uint32_t i = 0; i < pc.n_sym; i++) out[o + i] = dgz_resolve1(ps[i], win.data());
            win.swap(win2);
            win_len = (uint32_t)std::min<uint64_t>(SS_DGZ_WINDOW, (uint64_t)win_len + pc.n_sym);
            used++; members += pc.members;
            cur_bit = pc.end_bit;
            if (pc.status == SS_DGZ_END) { end = true; break; }
            if (pc.status == SS_DGZ_FULL || pc.next >= P) break;
            cur = pc.next;
        }
        if (rc || end) break;
    }
    delete tab;
    if (stopped_at) *stopped_at = (size_t)(cur_bit >> 3);
    if (stats4) { stats4[0] = found; stats4[1] = used; stats4[2] = batches; stats4[3] = members; }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// host-only self-test: the 16-bit decode tables of ss_dgz2.cuh against the 32-bit ones of ss_inflate.cuh on RANDOM
// prefix codes (zlib's own streams only show the code shapes zlib's heuristics produce; other compressors make others)
// ---------------------------------------------------------------------------------------------
namespace {
struct dgz_rng {
    uint64_t s;
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
    uint32_t below(uint32_t n) { return next() % n; }
};

// code lengths of a complete prefix code with `used` of `n` symbols (used >= 2): leaves split at random, deep ones preferred
void dgz_random_code(dgz_rng &r, uint8_t *lens, int n, int used, int max_len) {
    std::vector<int> depth(1, 0);
    while ((int)depth.size() < used) {
        size_t pick = r.below((uint32_t)depth.size());
        for (int tries = 0; tries < 3; tries++) {                  // prefer a deeper leaf: long codes, sub-tables
            const size_t other = r.below((uint32_t)depth.size());
            if (depth[other] > depth[pick] && depth[other] < max_len) pick = other;
        }
        if (depth[pick] >= max_len) {
            bool any = false;
            for (size_t i = 0; i < depth.size(); i++) if (depth[i] < max_len) { pick = i; any = true; break; }
            if (!any) break;
        }
        depth[pick]++;
        depth.push_back(depth[pick]);
    }
    for (int i = 0; i < n; i++) lens[i] = 0;
    std::vector<int> sym(n);
    for (int i = 0; i < n; i++) sym[i] = i;
    for (int i = n - 1; i > 0; i--) std::swap(sym[i], sym[r.below((uint32_t)i + 1)]);
    for (size_t i = 0; i < depth.size(); i++) lens[sym[i]] = (uint8_t)depth[i];
}

// what the next symbol of bit pattern v is, by either table: kind (0 literal, 1 base + extra, 2 end of block, 3 invalid),
// value, extra bits, code bits consumed
struct dgz_sym { uint32_t kind, val, extra, bits; };
dgz_sym dgz_lookup32(const uint32_t *t, uint32_t v, uint32_t tb) {
    uint32_t e = t[v & ((1u << tb) - 1u)], bits = 0;
    if (SSI_KIND(e) == SSI_SUB) { bits = tb; e = t[SSI_VAL(e) + ((v >> tb) & ((1u << SSI_EXTRA(e)) - 1u))]; }
    bits += SSI_LEN(e);
    switch (SSI_KIND(e)) {
    case SSI_LIT: return {0, SSI_VAL(e), 0, bits};
    case SSI_BASE: return {1, SSI_VAL(e), SSI_EXTRA(e), bits};
    case SSI_EOB: return {2, 0, 0, bits};
    default: return {3, 0, 0, 0};
    }
}
dgz_sym dgz_lookup16(const uint16_t *t, uint32_t v, uint32_t tb, uint32_t base_off) {
    uint32_t e = t[v & ((1u << tb) - 1u)], bits = 0;
    if (DGC_KIND(e) == DGC_SUB) { bits = tb; e = t[DGC_VAL(e) + ((v >> tb) & ((1u << DGC_LEN(e)) - 1u))]; }
    bits += DGC_LEN(e);
    if (DGC_KIND(e) == DGC_LIT) return {0, DGC_VAL(e), 0, bits};
    if (DGC_KIND(e) == DGC_SYM) { const uint32_t info = dgc_base_entry(base_off + DGC_VAL(e)); return {1, info & 0xFFFFu, info >> 16, bits}; }
    if (DGC_TAG(e) == DGC_TAG_EOB) return {2, 0, 0, bits};
    return {3, 0, 0, 0};
}
}   // namespace

int ss_dgz_tables_selftest(uint64_t seed, uint32_t trials, uint64_t *n_checked, std::string &err) {
    dgz_rng r{seed * 0x9E3779B97F4A7C15ull + 1u};
    ssi_tables *a = new ssi_tables;
    dgz_ctables *b = new dgz_ctables;
    uint64_t checked = 0;
    int rc = 0;
    for (uint32_t trial = 0; trial < trials && !rc; trial++) {
        uint8_t lens[288];
        // literal / length code: 257..286 symbols, 2..all of them used, lengths up to 15
        const int hlit = 257 + (int)r.below(30), used = 2 + (int)r.below((uint32_t)hlit - 1);
        dgz_random_code(r, lens, hlit, used, 15);
        const int ra = ssi_build_table(a->lit, SSI_LIT_CAP, lens, hlit, SSI_LIT_BITS, SSI_CODE_LIT);
        const int rb = dgc_build_table(b->lit, DGC_LIT_CAP, lens, hlit, DGC_LIT_BITS, SSI_CODE_LIT);
        if (ra != rb) { err = "the two builders disagree on whether a literal code is valid"; rc = 1; break; }
        if (!ra)
            for (uint32_t v = 0; v < 32768u; v++) {
                const dgz_sym x = dgz_lookup32(a->lit, v, SSI_LIT_BITS), y = dgz_lookup16(b->lit, v, DGC_LIT_BITS, 0);
                if (x.kind != y.kind || x.val != y.val || x.extra != y.extra || x.bits != y.bits) { err = "literal / length tables decode a bit pattern differently"; rc = 1; break; }
                checked++;
            }
        // distance code: 1..30 symbols; also the single code of length 1 and the empty code
        const int hdist = 1 + (int)r.below(30);
        const uint32_t shape = r.below(8);
        if (shape == 0) { for (int i = 0; i < hdist; i++) lens[i] = 0; }
        else if (shape == 1 || hdist < 2) { for (int i = 0; i < hdist; i++) lens[i] = 0; lens[r.below((uint32_t)hdist)] = 1; }
        else dgz_random_code(r, lens, hdist, 2 + (int)r.below((uint32_t)hdist - 1), 15);
        const int da = ssi_build_table(a->dist, SSI_DIST_CAP, lens, hdist, SSI_DIST_BITS, SSI_CODE_DIST);
        const int db = dgc_build_table(b->dist, DGC_DIST_CAP, lens, hdist, DGC_DIST_BITS, SSI_CODE_DIST);
        if (da != db) { err = "the two builders disagree on whether a distance code is valid"; rc = 1; break; }
        if (!da && !rc)
            for (uint32_t v = 0; v < 32768u; v++) {
                const dgz_sym x = dgz_lookup32(a->dist, v, SSI_DIST_BITS), y = dgz_lookup16(b->dist, v, DGC_DIST_BITS, 32);
                if (x.kind != y.kind || x.val != y.val || x.extra != y.extra || x.bits != y.bits) { err = "distance tables decode a bit pattern differently"; rc = 1; break; }
                checked++;
            }
        // an over-subscribed and an incomplete code must be refused by both
        if (hlit > 10) {
            for (int i = 0; i < hlit; i++) lens[i] = (uint8_t)(i < 5 ? 2 : 0);
            if (!ssi_build_table(a->lit, SSI_LIT_CAP, lens, hlit, SSI_LIT_BITS, SSI_CODE_LIT) || !dgc_build_table(b->lit, DGC_LIT_CAP, lens, hlit, DGC_LIT_BITS, SSI_CODE_LIT)) { err = "an over-subscribed code was accepted"; rc = 1; }
            for (int i = 0; i < hlit; i++) lens[i] = (uint8_t)(i < 3 ? 2 : 0);
            if (!ssi_build_table(a->lit, SSI_LIT_CAP, lens, hlit, SSI_LIT_BITS, SSI_CODE_LIT) || !dgc_build_table(b->lit, DGC_LIT_CAP, lens, hlit, DGC_LIT_BITS, SSI_CODE_LIT)) { err = "an incomplete code was accepted"; rc = 1; }
        }
    }
    delete a; delete b;
    if (n_checked) *n_checked = checked;
    return rc;
}
