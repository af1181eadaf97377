// ss_pgz.cuh -- parallel decoding of ONE ordinary gzip member on the host ingest threads.
//
// A `gzip file.fq` output is a single deflate stream: no independent units, every back-reference may
// reach 32 KiB behind.  The reference pipes it through one `zcat` (library/identify.py:82,
// library/Vote_Strain_L2_Lasso_new_sp.py:359,367).  Here a stream is decoded by T threads per round:
//
//   1. cut the next T * span compressed bytes into T pieces; piece 0 starts at a known block boundary
//      with a known 32 KiB window; for every other piece a deflate block start is FOUND by trying bit
//      positions (strict dynamic-header validation, then a trial decode of the whole block and a check of
//      the header behind it);
//   2. piece 0 is decoded normally; the others are decoded with an UNKNOWN window into 16-bit symbols
//      (literal 0..255, or 256 + offset into the unknown window), switching to the ordinary 8-bit decoder
//      as soon as the last 32 KiB they produced contain no unknown symbol;
//   3. every decoder stops when it stands on a block boundary that is another piece's start.  Pieces are
//      then stitched in order: piece j is accepted only if the piece before it ENDED exactly on j's start,
//      which is a proof that the found start was a real block boundary (false positives cost work, never
//      correctness); its symbols are resolved against the window the accepted text before it provides.
//
// The idea of marker symbols for the unknown window follows pugz (Kerbiriou & Chikhi, 2019); the code is
// written from scratch on top of ss_inflate.cuh.  Host only.
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <vector>

#include "ss_inflate.cuh"

#define PGZ_WINDOW 32768u

// growable array of 16-bit symbols (no zero-fill on growth: the decoder writes every element it exposes)
struct pgz_symbuf {
    uint16_t *p = nullptr;
    size_t n = 0, cap = 0;
    pgz_symbuf() = default;
    pgz_symbuf(const pgz_symbuf &) = delete;
    pgz_symbuf &operator=(const pgz_symbuf &) = delete;
    ~pgz_symbuf() { free(p); }
    bool room(size_t extra) {
        if (cap - n >= extra) return true;
        size_t nc = std::max<size_t>(cap * 2, n + extra + (1u << 16));
        uint16_t *q = (uint16_t *)realloc(p, nc * sizeof(uint16_t));
        if (!q) return false;
        p = q; cap = nc;
        return true;
    }
    size_t size() const { return n; }
    void clear() { n = 0; }
};

// position a fresh stream on absolute bit `bit` of the file
inline void pgz_seek(ssi_stream &s, const uint8_t *base, size_t size, uint64_t bit) {
    ssi_stream_init(s, base + (bit >> 3), base + size);
    s.base = base;
    ssi_refill(s.bits);
    ssi_drop(s.bits, (uint32_t)(bit & 7u));
}

// read the 3 header bits of a block and prepare its tables; returns the block type or < 0
inline int pgz_block_header(ssi_stream &s, ssi_tables &t) {
    ssi_bits &b = s.bits;
    ssi_refill(b);
    s.last_block = (int)ssi_take(b, 1);
    uint32_t type = ssi_take(b, 2);
    if (type == 0) {
        ssi_drop(b, b.cnt & 7u);
        ssi_refill(b);
        uint32_t len = ssi_take(b, 16), nlen = ssi_take(b, 16);
        if ((len ^ nlen) != 0xFFFFu) return ssi_truncated(b) ? SSI_ERR_TRUNC : SSI_ERR_DATA;
        s.stored_left = len;
        return 0;
    }
    if (type == 1) return ssi_fixed_tables(t) ? SSI_ERR_DATA : 1;
    if (type == 2) { int rc = ssi_dynamic_tables(b, t); return rc ? rc : 2; }
    return ssi_truncated(b) ? SSI_ERR_TRUNC : SSI_ERR_DATA;
}

// One deflate block with an unknown window: appends 16-bit symbols to `sym`.  The stream stands behind the
// block afterwards.  Returns 0 or an SSI_ERR_* code.
inline int pgz_block_markers(ssi_stream &s, ssi_tables &t, pgz_symbuf &sym) {
    int type = pgz_block_header(s, t);
    if (type < 0) return type;
    ssi_bits &b = s.bits;
    if (type == 0) {
        if (!sym.room(s.stored_left)) return SSI_ERR_DATA;
        while (s.stored_left && (b.cnt >> 3) > b.overrun) { sym.p[sym.n++] = (uint16_t)ssi_take(b, 8); s.stored_left--; }
        if (s.stored_left) {
            if (b.overrun || (size_t)(b.in_end - b.in) < s.stored_left) return SSI_ERR_TRUNC;
            b.buf = 0; b.cnt = 0;
            for (uint32_t i = 0; i < s.stored_left; i++) sym.p[sym.n++] = b.in[i];
            b.in += s.stored_left;
            s.stored_left = 0;
        }
        return 0;
    }
    size_t n = sym.n;
    bool eob = false;
    // fast path (same shape as ssi_huff_fast): bit state in registers, up to three literals per refill, the
    // next entry loaded before a match is copied; needs input and output slack, the careful loop below
    // finishes the block
    while (!eob && !b.overrun && b.in_end - b.in >= 16) {
        if (sym.cap - n < (1u << 16)) { sym.n = n; if (!sym.room(1u << 18)) return SSI_ERR_DATA; }
        uint16_t *o = sym.p;
        const size_t n_safe = sym.cap - 600;
        const uint8_t *in = b.in, *const in_safe = b.in_end - 16;
        uint64_t buf = b.buf;
        uint32_t cnt = b.cnt;
        const uint32_t LM = (1u << SSI_LIT_BITS) - 1u, DM = (1u << SSI_DIST_BITS) - 1u;
        int bad = 0;
#define PGZ_REFILL() do { uint64_t w_; memcpy(&w_, in, 8); buf |= w_ << cnt; in += (63u - cnt) >> 3; cnt |= 56u; } while (0)
        PGZ_REFILL();
        uint32_t e = t.lit[buf & LM];
        while (in <= in_safe && n <= n_safe) {
            if (SSI_KIND(e) == SSI_SUB) { buf >>= SSI_LIT_BITS; cnt -= SSI_LIT_BITS; e = t.lit[SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u))]; }
            buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
            if (SSI_KIND(e) == SSI_LIT) {
                o[n++] = (uint16_t)SSI_VAL(e);
                e = t.lit[buf & LM];
                if (SSI_KIND(e) == SSI_LIT) {
                    buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
                    o[n++] = (uint16_t)SSI_VAL(e);
                    e = t.lit[buf & LM];
                    if (SSI_KIND(e) == SSI_LIT) {
                        buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
                        o[n++] = (uint16_t)SSI_VAL(e);
                        PGZ_REFILL();
                        e = t.lit[buf & LM];
                        continue;
                    }
                }
                PGZ_REFILL();
                continue;
            }
            if (SSI_KIND(e) != SSI_BASE) { if (SSI_KIND(e) == SSI_EOB) eob = true; else bad = 1; break; }
            uint32_t len = SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u));
            buf >>= SSI_EXTRA(e); cnt -= SSI_EXTRA(e);
            uint32_t d = t.dist[buf & DM];
            if (SSI_KIND(d) == SSI_SUB) { buf >>= SSI_DIST_BITS; cnt -= SSI_DIST_BITS; d = t.dist[SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u))]; }
            buf >>= SSI_LEN(d); cnt -= SSI_LEN(d);
            if (SSI_KIND(d) != SSI_BASE) { bad = 1; break; }
            uint32_t dist = SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u));
            buf >>= SSI_EXTRA(d); cnt -= SSI_EXTRA(d);
            if (dist > PGZ_WINDOW) { bad = 1; break; }
            PGZ_REFILL();
            e = t.lit[buf & LM];
            if (dist <= n) {
                const uint16_t *src = o + n - dist;
                uint16_t *dst = o + n, *end = dst + len;
                if (dist >= 4) { do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 4; dst += 4; } while (dst < end); }
                else { do { *dst++ = *src++; } while (dst < end); }
            } else {                                            // reaches into the unknown window
                for (uint32_t i = 0; i < len; i++) {
                    int64_t idx = (int64_t)(n + i) - (int64_t)dist;
                    o[n + i] = idx >= 0 ? o[idx] : (uint16_t)(256 + (int64_t)PGZ_WINDOW + idx);
                }
            }
            n += len;
        }
#undef PGZ_REFILL
        b.in = in; b.buf = buf; b.cnt = cnt;
        if (bad) { sym.n = n; return SSI_ERR_DATA; }
        if (in > in_safe) break;                                // input slack used up: careful loop
    }
    while (!eob) {
        if (sym.cap - n < 260) { sym.n = n; if (!sym.room(260)) return SSI_ERR_DATA; }
        uint16_t *o = sym.p;
        ssi_refill(b);
        uint32_t e = t.lit[ssi_peek(b, SSI_LIT_BITS)];
        if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_LIT_BITS); e = t.lit[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
        ssi_drop(b, SSI_LEN(e));
        uint32_t kind = SSI_KIND(e);
        if (kind == SSI_LIT) {
            o[n++] = (uint16_t)SSI_VAL(e);
            e = t.lit[ssi_peek(b, SSI_LIT_BITS)];                  // a second literal from the same refill
            if (SSI_KIND(e) == SSI_LIT) { ssi_drop(b, SSI_LEN(e)); o[n++] = (uint16_t)SSI_VAL(e); }
            continue;
        }
        if (kind == SSI_EOB) break;
        if (kind != SSI_BASE) { sym.n = n; return SSI_ERR_DATA; }
        uint32_t len = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
        e = t.dist[ssi_peek(b, SSI_DIST_BITS)];
        if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_DIST_BITS); e = t.dist[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
        ssi_drop(b, SSI_LEN(e));
        if (SSI_KIND(e) != SSI_BASE) { sym.n = n; return SSI_ERR_DATA; }
        uint32_t dist = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
        if (dist > PGZ_WINDOW) { sym.n = n; return SSI_ERR_DATA; }
        if (dist <= n) {
            const uint16_t *src = o + n - dist;
            for (uint32_t i = 0; i < len; i++) o[n + i] = src[i];
        } else {
            for (uint32_t i = 0; i < len; i++) {
                int64_t idx = (int64_t)(n + i) - (int64_t)dist;
                o[n + i] = idx >= 0 ? o[idx] : (uint16_t)(256 + (int64_t)PGZ_WINDOW + idx);
            }
        }
        n += len;
        if (ssi_truncated(b)) { sym.n = n; return SSI_ERR_TRUNC; }
    }
    sym.n = n;
    return ssi_truncated(b) ? SSI_ERR_TRUNC : 0;
}

// Does a deflate block start at bit `p`?  Cheap rejections first (7 of 8 positions fall at the 3 type bits),
// then the complete dynamic header, then a trial decode of the whole block and the header behind it.
inline bool pgz_is_block_start(const uint8_t *base, size_t size, uint64_t p, ssi_tables &t, pgz_symbuf &scratch) {
    const size_t byte = (size_t)(p >> 3);
    if (byte + 16 > size) return false;
    uint64_t w;
    memcpy(&w, base + byte, 8);
    w >>= (p & 7u);
    if ((w & 7u) != 4u) return false;                              // BFINAL = 0, BTYPE = 2 (dynamic)
    if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;   // HLIT, HDIST
    {   // the code-length code must be complete (zlib emits nothing else): sum 2^(7-len) == 128
        const uint32_t hclen = (uint32_t)((w >> 13) & 15u) + 4u;
        const uint32_t have = 47u - (uint32_t)(p & 7u);            // valid bits left in w behind the 17 header bits
        uint64_t w2;
        memcpy(&w2, base + byte + 8, 8);
        uint64_t v = (w >> 17) | (w2 << have);                     // >= 57 bits: up to 19 lengths of 3 bits
        uint32_t kraft = 0;
        for (uint32_t i = 0; i < hclen; i++, v >>= 3) {
            uint32_t l = (uint32_t)(v & 7u);
            if (l) kraft += 128u >> l;
        }
        if (kraft != 128u) return false;
    }
    ssi_stream s;
    pgz_seek(s, base, size, p);
    scratch.clear();
    if (pgz_block_markers(s, t, scratch) != 0 || s.last_block) return false;
    if (scratch.size() < 64) return false;                          // a real block of a large stream is not this small
    // the header behind it
    ssi_stream s2 = s;
    int type = pgz_block_header(s2, t);
    return type >= 0;
}

inline uint64_t pgz_find_block(const uint8_t *base, size_t size, uint64_t from_bit, uint64_t to_bit, ssi_tables &t,
                               pgz_symbuf &scratch) {
    for (uint64_t p = from_bit; p < to_bit; p++)
        if (pgz_is_block_start(base, size, p, t, scratch)) return p;
    return ~0ull;
}

// ---- one piece of a round -----------------------------------------------------------------------------
struct pgz_piece {
    uint64_t start_bit = ~0ull, end_bit = 0;
    bool valid = false;            // a start was found (piece 0: always)
    int status = SSI_ERR_DATA;     // SSI_STOP, SSI_OK (member end) or an error
    pgz_symbuf sym;                // unknown-window part
    std::vector<uint8_t> text;     // 8-bit part; text[0, hist) is history, not output
    size_t hist = 0, text_len = 0; // text_len: bytes in `text` including the history prefix
    std::vector<uint8_t> resolved; // the unknown-window part as bytes (filled while stitching)
    // buffers keep their memory from round to round (no page faults after the first rounds)
    void reset() { start_bit = ~0ull; end_bit = 0; valid = false; status = SSI_ERR_DATA; sym.clear(); hist = text_len = 0; }
};

// 8-bit decoding into a growing vector until a stop, the limit or the member end
inline int pgz_run_8bit(ssi_stream &s, ssi_tables &t, pgz_piece &pc) {
    while (true) {
        if (pc.text.size() - pc.text_len < (1u << 16)) pc.text.resize(std::max<size_t>(pc.text.size() * 3 / 2, pc.text_len + (4u << 20)));
        uint8_t *pos = pc.text.data() + pc.text_len;
        int rc = ssi_inflate(s, t, &pos, pc.text.data() + pc.text.size());
        pc.text_len = (size_t)(pos - pc.text.data());
        if (rc != SSI_MORE_OUTPUT) return rc;
    }
}

// piece 0: known window (the last <= 32 KiB before the piece) and known member offset
inline void pgz_decode_known(const uint8_t *base, size_t size, pgz_piece &pc, const std::vector<uint8_t> &window,
                             uint64_t member_out, const std::vector<uint64_t> &stops, uint64_t limit_bit) {
    ssi_tables *t = new ssi_tables;
    ssi_stream s;
    pgz_seek(s, base, size, pc.start_bit);
    s.stops = stops.data(); s.n_stops = (uint32_t)stops.size(); s.limit_bit = limit_bit;
    s.out_total = member_out;
    if (pc.text.size() < window.size() + (4u << 20)) pc.text.resize(window.size() + (8u << 20));
    memcpy(pc.text.data(), window.data(), window.size());
    pc.hist = pc.text_len = window.size();
    pc.status = pgz_run_8bit(s, *t, pc);
    pc.end_bit = ssi_bitpos(s);
    delete t;
}

// pieces 1..: find a block start in [from_bit, to_bit), then decode with an unknown window
inline void pgz_decode_unknown(const uint8_t *base, size_t size, pgz_piece &pc, const uint64_t *stops, uint32_t n_stops,
                               uint64_t limit_bit) {
    ssi_tables *t = new ssi_tables;
    ssi_stream s;
    pgz_seek(s, base, size, pc.start_bit);
    pc.status = SSI_STOP;
    bool clean = false;
    while (true) {                                                   // block by block with markers
        int rc = pgz_block_markers(s, *t, pc.sym);
        if (rc) { pc.status = rc; break; }
        const uint64_t at = ssi_bitpos(s);
        if (s.last_block) { pc.status = SSI_OK; break; }
        bool stop = limit_bit && at >= limit_bit;
        for (uint32_t i = 0; i < n_stops && !stop; i++) stop = stops[i] == at;
        if (stop) break;
        if (pc.sym.size() >= PGZ_WINDOW) {
            const uint16_t *w = pc.sym.p + pc.sym.size() - PGZ_WINDOW;
            uint16_t mx = 0;
            for (uint32_t i = 0; i < PGZ_WINDOW; i++) mx = std::max(mx, w[i]);
            if (mx < 256) { clean = true; break; }
        }
    }
    if (clean) {                                                     // no unknown symbol can be copied any more
        if (pc.text.size() < PGZ_WINDOW + (4u << 20)) pc.text.resize(PGZ_WINDOW + (8u << 20));
        const uint16_t *w = pc.sym.p + pc.sym.size() - PGZ_WINDOW;
        for (uint32_t i = 0; i < PGZ_WINDOW; i++) pc.text[i] = (uint8_t)w[i];
        pc.hist = pc.text_len = PGZ_WINDOW;
        s.phase = SSI_PH_BLOCK; s.last_block = 0; s.stored_left = 0;
        s.out_total = PGZ_WINDOW;
        s.stops = stops; s.n_stops = n_stops; s.limit_bit = limit_bit;
        pc.status = pgz_run_8bit(s, *t, pc);
    }
    pc.end_bit = ssi_bitpos(s);
    delete t;
}

// ---- the member decoder -------------------------------------------------------------------------------
// the pieces of one round: decoded by pgz_round_decode, handed out by pgz_round_emit (two of these let the
// next round decode while the previous one is resolved and copied out)
struct pgz_roundbuf {
    std::vector<pgz_piece> pieces;                     // reused from round to round
    std::vector<int> order;                            // accepted pieces, in stream order
    std::vector<std::vector<uint8_t>> win_before;      // the window in front of each accepted piece
    int T = 0;
    int rc = SSI_ERR_DATA;                             // 0 more of the member follows, 1 member end, < 0 error
};

struct pgz_member {
    const uint8_t *base = nullptr;
    size_t size = 0;
    uint64_t bit = 0;                 // a block boundary of the member (absolute bit position)
    std::vector<uint8_t> window;      // the last <= 32 KiB of text produced before `bit`
    uint64_t member_out = 0;
    uint64_t pieces_found = 0, pieces_used = 0, rounds = 0, marker_syms = 0;
    double t_find = 0, t_decode = 0, t_stitch = 0;    // seconds, summed over rounds
};

inline double pgz_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// unknown-window symbols -> bytes, given the window in front of the piece
inline bool pgz_resolve(const uint16_t *sym, size_t n, const std::vector<uint8_t> &win, uint8_t *out) {
    const size_t wn = win.size();
    bool ok = true;
    for (size_t i = 0; i < n; i++) {
        uint16_t v = sym[i];
        if (v < 256) out[i] = (uint8_t)v;
        else {
            size_t off = (size_t)v - 256u;                       // 0 = 32 KiB before the piece, 32767 = the byte before it
            if (off + wn < PGZ_WINDOW) { ok = false; out[i] = 0; }   // reaches before the start of the member
            else out[i] = win[off + wn - PGZ_WINDOW];
        }
    }
    return ok;                                                   // (copies of unknown symbols carry the same offsets)
}

// Decode one round with T threads of `span` compressed bytes each into `rb`, and advance the member state
// (bit position, window, bytes produced) past it.  rb.rc: 0 more follows, 1 the member's final block was
// decoded (m.bit stands behind it), < 0 an SSI_ERR_* code.
inline int pgz_round_decode(pgz_member &m, pgz_roundbuf &rb, int T, size_t span) {
    const uint64_t size_bits = (uint64_t)m.size * 8u;
    const uint64_t first_byte = m.bit >> 3;
    T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)T, (m.size - first_byte) / span + 1));
    if (rb.pieces.size() < (size_t)T) { std::vector<pgz_piece> np((size_t)T); rb.pieces.swap(np); }
    std::vector<pgz_piece> &pieces = rb.pieces;
    for (auto &pc : pieces) pc.reset();
    rb.T = T; rb.order.clear(); rb.rc = SSI_ERR_DATA;
    pieces[0].start_bit = m.bit; pieces[0].valid = true;
    const uint64_t limit_bit = std::min<uint64_t>(size_bits, (first_byte + (uint64_t)T * span) * 8u);
    double t0 = pgz_now();
    // phase A: find the block starts of pieces 1..T-1
    if (T > 1) {
        std::vector<std::thread> th;
        for (int j = 1; j < T; j++)
            th.emplace_back([&, j]() {
                ssi_tables *t = new ssi_tables;
                pgz_symbuf scratch;
                uint64_t from = (first_byte + (uint64_t)j * span) * 8u, to = std::min<uint64_t>(from + span * 8u, size_bits);
                uint64_t p = from < to ? pgz_find_block(m.base, m.size, from, to, *t, scratch) : ~0ull;
                if (p != ~0ull) { pieces[(size_t)j].start_bit = p; pieces[(size_t)j].valid = true; }
                delete t;
            });
        for (auto &x : th) x.join();
    }
    std::vector<uint64_t> stops;
    for (int j = 1; j < T; j++) if (pieces[(size_t)j].valid) stops.push_back(pieces[(size_t)j].start_bit);
    m.pieces_found += stops.size();
    double t1 = pgz_now();
    m.t_find += t1 - t0;
    // phase B: decode
    {
        std::vector<std::thread> th;
        for (int j = 1; j < T; j++) {
            if (!pieces[(size_t)j].valid) continue;
            th.emplace_back([&, j]() {
                // stop on any LATER start
                const uint64_t *first = std::upper_bound(stops.data(), stops.data() + stops.size(), pieces[(size_t)j].start_bit);
                pgz_decode_unknown(m.base, m.size, pieces[(size_t)j], first, (uint32_t)(stops.data() + stops.size() - first), limit_bit);
            });
        }
        pgz_decode_known(m.base, m.size, pieces[0], m.window, m.member_out, stops, limit_bit);
        for (auto &x : th) x.join();
    }
    m.rounds++;
    double t2 = pgz_now();
    m.t_decode += t2 - t1;
    // stitch: a piece is used only if the one before it ended exactly on its start
    for (int cur = 0;;) {
        pgz_piece &pc = pieces[(size_t)cur];
        if (pc.status < 0) return rb.rc = pc.status;
        rb.order.push_back(cur);
        if (pc.status == SSI_OK) break;                              // member end
        int next = -1;
        for (int j = cur + 1; j < T; j++)
            if (pieces[(size_t)j].valid && pieces[(size_t)j].start_bit == pc.end_bit) { next = j; break; }
        if (next < 0) break;                                         // stopped on the round's limit
        cur = next;
    }
    // the window in front of every accepted piece: sequential, but only the last 32 KiB of each piece matter
    rb.win_before.resize(rb.order.size());
    std::vector<uint8_t> w = m.window, tail, nw;
    for (size_t k = 0; k < rb.order.size(); k++) {
        pgz_piece &pc = pieces[(size_t)rb.order[k]];
        rb.win_before[k] = w;
        const size_t n8 = pc.text_len - pc.hist, n16 = rb.order[k] ? pc.sym.size() : 0;
        nw.clear();
        if (n8 < PGZ_WINDOW) {                                       // new window = last 32 KiB of (w | resolved sym | 8-bit text)
            size_t need = PGZ_WINDOW - n8, take16 = std::min(need, n16);
            tail.resize(take16);
            if (!pgz_resolve(pc.sym.p + n16 - take16, take16, w, tail.data())) return rb.rc = SSI_ERR_DATA;
            size_t from_w = std::min(w.size(), need - take16);
            nw.insert(nw.end(), w.end() - (long)from_w, w.end());
            nw.insert(nw.end(), tail.begin(), tail.end());
        }
        size_t t8 = std::min<size_t>(n8, PGZ_WINDOW);
        nw.insert(nw.end(), pc.text.data() + pc.text_len - t8, pc.text.data() + pc.text_len);
        w.swap(nw);
        m.member_out += n16 + n8;
        if (k > 0) { m.marker_syms += n16; m.pieces_used++; }
        m.bit = pc.end_bit;
    }
    m.window.swap(w);
    m.t_stitch += pgz_now() - t2;
    return rb.rc = pieces[(size_t)rb.order.back()].status == SSI_OK ? 1 : 0;
}

// Hand the text of a decoded round out in order.  `emit(n, fill)` receives each part as a length and a
// function fill(dst, off, len) that writes bytes [off, off + len) of the part to dst -- a memcpy for the
// 8-bit parts, the symbol resolution for the unknown-window parts -- so that the receiver can produce the
// bytes directly where they are needed, several slices at a time.  Touches only `rb`.
// Returns rb.rc, or SSI_ERR_DATA for a reference before the member start.
template <typename Emit>
int pgz_round_emit(pgz_roundbuf &rb, Emit &&emit) {
    if (rb.rc < 0) return rb.rc;
    std::atomic<int> bad{0};
    for (size_t k = 0; k < rb.order.size(); k++) {
        pgz_piece &pc = rb.pieces[(size_t)rb.order[k]];
        if (k > 0 && pc.sym.size()) {
            const uint16_t *sym = pc.sym.p;
            const std::vector<uint8_t> &win = rb.win_before[k];
            emit(pc.sym.size(), [&, sym](uint8_t *dst, size_t off, size_t len) { if (!pgz_resolve(sym + off, len, win, dst)) bad = 1; });
        }
        const size_t n8 = pc.text_len - pc.hist;
        if (n8) {
            const uint8_t *src = pc.text.data() + pc.hist;
            emit(n8, [src](uint8_t *dst, size_t off, size_t len) { memcpy(dst, src + off, len); });
        }
        if (bad) return SSI_ERR_DATA;
    }
    return rb.rc;
}

// run fill(dst, off, len) over [0, n) in slices on up to `threads` threads
template <typename Fill>
void pgz_parallel_fill(uint8_t *dst, size_t n, int threads, Fill &&fill) {
    const size_t min_slice = 1u << 20;
    int t = (int)std::min<size_t>((size_t)std::max(1, threads), n / min_slice);
    if (t <= 1) { fill(dst, 0, n); return; }
    std::vector<std::thread> th;
    size_t per = (n + (size_t)t - 1) / (size_t)t;
    for (int i = 1; i < t; i++) {
        size_t lo = per * (size_t)i, hi = std::min(n, lo + per);
        if (lo < hi) th.emplace_back([&, lo, hi]() { fill(dst + lo, lo, hi - lo); });
    }
    fill(dst, 0, std::min(n, per));
    for (auto &x : th) x.join();
}

// Whole member, two round buffers: round r+1 is decoded while round r is resolved and handed out.
// Returns 1 at the member end (m.bit stands behind the final block) or an SSI_ERR_* code; `keep_going()`
// is polled between rounds (false -> returns 0).
template <typename Emit, typename Keep>
int pgz_member_decode(pgz_member &m, int T, size_t span, Emit &&emit, Keep &&keep_going) {
    pgz_roundbuf *rb = new pgz_roundbuf[2];
    int cur = 0, rc = pgz_round_decode(m, rb[0], T, span);
    while (true) {
        if (rc < 0) break;
        const bool more = rc == 0;
        std::thread next;
        int rc_next = 0;
        if (more) next = std::thread([&]() { rc_next = pgz_round_decode(m, rb[cur ^ 1], T, span); });
        double t0 = pgz_now();
        int erc = pgz_round_emit(rb[cur], emit);
        double dt = pgz_now() - t0;
        if (more) next.join();
        m.t_stitch += dt;
        if (erc < 0) { rc = erc; break; }
        if (!more) { rc = 1; break; }
        if (!keep_going()) { rc = 0; break; }
        rc = rc_next;
        cur ^= 1;
    }
    delete[] rb;
    return rc;
}
