// ss_dgz_host.h -- host driver of the device inflate of ordinary gzip streams (ss_dgz.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

#include "ss_dgz.cuh"

// bytes behind a batch's limit that must have arrived before the batch is decoded (a decoder runs on to the end of the
// deflate block that straddles the limit); what lies further is out of bounds for that batch's kernels
#define SS_DGZ_GATE_SLACK (8u << 20)
#define SS_DGZ_MAX_PIECES 16384u          // pieces per batch at most
// K8's shape: decoders per warp x warps per CTA of ss_dgz2.cuh (SS_DGZ_LANES, SS_DGZ_WARPS), or {0, 0}: one decoder
// per one-warp CTA (ss_dgz.cuh)
#ifndef SS_DGZ_LANES_DEFAULT
#define SS_DGZ_LANES_DEFAULT 2
#endif
#ifndef SS_DGZ_RING_DEFAULT
#define SS_DGZ_RING_DEFAULT 0
#endif
#ifndef SS_DGZ_LOCKSTEP_DEFAULT
#define SS_DGZ_LOCKSTEP_DEFAULT 1
#endif
struct ss_dgz_shape { int lanes, warps, lockstep, ring; };

// How one file is cut up for the device: piece size, pieces per batch, symbol room per piece, text room per batch.
//   - a WAVE is one piece per resident decoder (n_sm x decoders_per_sm); a wave's text must fit the batch buffer, so the
//     piece size is bounded by scratch / (wave x expected text per compressed byte), and by 128 KiB;
//   - the file should be a whole number of waves (a last wave with a few pieces takes as long as a full one): the
//     smallest number of waves whose pieces respect the bound, pieces sized to fill them exactly;
//   - `ratio` = text bytes per compressed byte the stream's head inflated to (dgz_eligible measures it on the first
//     64 KiB of text; the empty window at the start makes it an under-estimate, hence the margins): it sizes the symbol
//     room of a piece -- a piece that runs out of it ends the batch early (status FULL), so a fixed room would turn
//     well-compressed inputs into one piece per batch.
// Environment overrides (tests, experiments): SS_DGZ_BATCH_MB, SS_DGZ_PIECE_BYTES, SS_DGZ_MAX_PIECES, SS_DGZ_SYM_PER_BYTE.
struct ss_dgz_plan {
    uint32_t piece;          // compressed bytes per piece
    uint32_t max_pieces;     // pieces per batch at most
    uint32_t expand;         // symbols a piece may produce per compressed byte of its nominal size
    size_t scratch;          // text bytes per batch
    size_t device_bytes;     // what the decoder's device buffers and the batch buffer take with this plan
    double ratio;            // the estimate the plan was made with
};
ss_dgz_plan ss_dgz_make_plan(size_t n_up, int n_sm, ss_dgz_shape shape, double ratio);

// One gzip file (or the members of one file part) inflated batch by batch on the device.  The compressed bytes
// [0, comp_size) sit in device memory (d_comp, padded by 16 readable bytes) and in host memory (h_comp: headers are
// parsed there).  Text comes out in stream order, cut anywhere (the caller carries partial records over).
class ss_dgz {
public:
    ss_dgz() = default;
    ~ss_dgz();
    ss_dgz(const ss_dgz &) = delete;
    ss_dgz &operator=(const ss_dgz &) = delete;
    // first_member: byte offset of a member header; members that start at or behind stop_member_at are not decoded
    int open(int n_sm, cudaStream_t st, const uint8_t *d_comp, const uint8_t *h_comp, size_t comp_size, size_t first_member,
             size_t stop_member_at, const ss_dgz_plan &plan);
    size_t held_bytes() const;                             // device memory the buffers take now
    // While the compressed bytes are still being uploaded: wait(need) blocks until bytes [0, need) of the file are on the
    // device (false: the upload failed); complete() tells whether everything has arrived.  A batch that fails while the
    // upload was still running is decoded again once it is complete (a block longer than the slack the gate allows).
    void set_input_gate(std::function<bool(size_t)> wait, std::function<bool()> complete) { gate_wait_ = wait; gate_complete_ = complete; }
    size_t batch_text_capacity() const;
    int next(uint8_t *d_out, size_t out_cap, size_t *n_out, bool *done);
    void close();
    const std::string &error() const { return err_; }
    size_t stopped_at() const { return (size_t)(cur_bit_ >> 3); }     // after `done`: the byte the decoder stands on
    uint64_t members() const { return members_; }
    uint64_t pieces_found() const { return pieces_found_; }
    uint64_t pieces_used() const { return pieces_used_; }
    uint64_t batches() const { return batches_; }
    double ms_decode() const { return ms_decode_; }
    double ms_resolve() const { return ms_resolve_; }
    double ms_windows() const { return ms_windows_; }     // part of ms_resolve, SS_DEBUG_TIMING only
    static ss_dgz_shape shape_from_env();
    static uint32_t decoders_per_sm(ss_dgz_shape sh);      // K8 decoders resident per SM

private:
    int n_sm_ = 0;
    ss_dgz_shape shape_ = {0, 0, 1, 0};
    uint32_t rounds_ = 256;
    cudaStream_t st_ = nullptr;
    const uint8_t *d_comp_ = nullptr, *h_comp_ = nullptr;
    size_t size_ = 0, stop_at_ = 0;
    uint32_t max_pieces_ = 0, cap_ = 0, piece_ = 0;
    size_t alloc_pieces_ = 0, alloc_sym_ = 0;     // what the device buffers were allocated for (they only grow)
    uint64_t cur_bit_ = 0;
    uint32_t win_len_ = 0;
    size_t gate_slack_ = SS_DGZ_GATE_SLACK;
    double ms_decode_ = 0, ms_resolve_ = 0, ms_windows_ = 0;
    double ratio_ = 4.0;              // text bytes per compressed byte seen so far
    bool done_ = false;
    uint64_t members_ = 0, pieces_found_ = 0, pieces_used_ = 0, batches_ = 0;
    dgz_piece *d_pieces_ = nullptr;
    uint16_t *d_sym_ = nullptr;
    uint8_t *d_windows_ = nullptr;
    uint32_t *d_order_ = nullptr;
    uint64_t *d_off_ = nullptr;
    unsigned int *d_ctr_ = nullptr;
    uint16_t *d_maps_ = nullptr, *d_gpre_ = nullptr;     // tail maps of the accepted pieces / composed group totals (K9)
    uint32_t *d_wlen_ = nullptr;                          // window bytes that exist in front of every accepted piece
    std::vector<dgz_piece> h_pieces_;
    std::string err_;
    std::function<bool(size_t)> gate_wait_;
    std::function<bool()> gate_complete_;
};

// the same pipeline on host threads (tests: the algorithms against zlib without a GPU); out is resized to the text
int ss_dgz_host_inflate(const uint8_t *comp, size_t comp_size, size_t first_member, size_t stop_member_at,
                        uint32_t max_pieces, uint32_t piece_bytes, uint32_t sym_per_byte, std::vector<uint8_t> &out,
                        size_t *stopped_at, uint64_t *stats4, std::string &err);

// host-only self-test of the 16-bit decode tables (ss_dgz2.cuh) against the 32-bit ones (ss_inflate.cuh) on random
// prefix codes: every 15-bit pattern must decode to the same symbol, value, extra bits and length.  0 = agree.
int ss_dgz_tables_selftest(uint64_t seed, uint32_t trials, uint64_t *n_checked, std::string &err);
