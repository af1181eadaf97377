// ss_dgz.cuh -- ordinary gzip (one deflate stream per member, no independent blocks) inflated by MANY decoders at
// once, on the device.  What `zcat a b |` does serially at library/identify.py:82 / Vote_Strain_L2_Lasso_new_sp.py:
// 359,367, and what ss_pgz.cuh does with a dozen host threads, done with thousands of GPU lanes:
//
//   K7 find     the compressed bytes of a batch are cut into pieces of SS_DGZ_PIECE bytes; for every piece but the
//               first a deflate block start is FOUND: lanes try consecutive bit positions (type bits, HLIT / HDIST,
//               complete code-length code), a survivor's whole dynamic header must build two valid Huffman codes;
//   K8 decode   one decoder per piece walks the stream from its start with an UNKNOWN 32 KiB window into 16-bit
//               symbols (literal, or 256 + offset into the unknown window -- the marker idea of pugz, Kerbiriou &
//               Chikhi 2019) until it stands exactly on a later piece's start (or the batch end);
//   chain       a piece is accepted only if the piece before it ENDED exactly on its start -- the proof that the
//               found start was a real block boundary; pieces that were stepped over are dropped (host, trivial);
//   K9 windows  the 32 KiB window in front of every accepted piece, piece after piece (one CTA; only the last
//               32 KiB of a piece matter);
//   K10 resolve symbols -> bytes at their final place in the FASTQ text (all pieces at once).
//
// The algorithms are SS_HD (host + device): tests run them on the CPU against zlib (ss_dgz_host_inflate), the
// kernels in ss_dgz.cu run the same code.  Exactness never rests on the search heuristics: a false start costs work.
#pragma once
#include "ss_inflate.cuh"

#define SS_DGZ_WINDOW 32768u
#define SS_DGZ_PIECE_DEFAULT (128u << 10)  // compressed bytes per piece at most (ss_dgz_make_plan; SS_DGZ_PIECE_BYTES)
#define SS_DGZ_EXPAND_DEFAULT 5u          // symbols a piece may produce per compressed byte of its nominal size, at least (ss_dgz_make_plan)

enum {
    SS_DGZ_LINKED = 0,      // stopped exactly on a later piece's start (dgz_piece::next) or on the batch limit
    SS_DGZ_END = 1,         // the stream ended: final block of a member and no further member behind it
    SS_DGZ_FULL = 2,        // out of symbol room at a block boundary (end_bit says where): the chain continues elsewhere
    SS_DGZ_ERROR = -1,      // invalid deflate data
    SS_DGZ_NOSTART = -2     // no block start was found for this piece
};

struct dgz_piece {
    uint64_t start_bit;     // absolute bit position in the batch's compressed bytes (~0 = none found)
    uint64_t end_bit;       // where the decoder stopped (a block boundary)
    uint32_t n_sym;         // symbols produced
    uint32_t next;          // LINKED: index of the piece whose start it stands on (n_pieces = the batch limit)
    int32_t status;
    uint32_t members;       // gzip members that ENDED inside this piece
    uint32_t fresh_from;    // symbols [fresh_from, n_sym) lie behind a member start seen by this piece (no unknown window)
    uint32_t pad_;
};

// position a fresh stream on absolute bit `bit` of `base`
SS_HD void dgz_seek(ssi_stream &s, const uint8_t *base, size_t size, uint64_t bit) {
    ssi_stream_init(s, base + (bit >> 3), base + size);
    s.base = base;
    ssi_refill(s.bits);
    ssi_drop(s.bits, (uint32_t)(bit & 7u));
}

// cheap test of bit position p: BFINAL = 0, BTYPE = 2, HLIT / HDIST in range, complete code-length code
SS_HD bool dgz_quick_test(const uint8_t *base, size_t size, uint64_t p) {
    const size_t byte = (size_t)(p >> 3);
    if (byte + 16 > size) return false;
    uint64_t w = ssi_load64(base + byte), w2 = ssi_load64(base + byte + 8);
    const uint32_t sh = (uint32_t)(p & 7u);
    w >>= sh;
    if ((w & 7u) != 4u) return false;
    if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;
    const uint32_t hclen = (uint32_t)((w >> 13) & 15u) + 4u;
    const uint32_t have = 47u - sh;                                // valid bits of w behind the 17 header bits
    uint64_t v = (w >> 17) | (w2 << have);
    uint32_t kraft = 0;
    for (uint32_t i = 0; i < hclen; i++, v >>= 3) {
        uint32_t l = (uint32_t)(v & 7u);
        if (l) kraft += 128u >> l;
    }
    return kraft == 128u;
}

// full test: the dynamic header at p builds both Huffman codes
SS_HD bool dgz_full_test(const uint8_t *base, size_t size, uint64_t p, ssi_tables &t) {
    ssi_stream s;
    dgz_seek(s, base, size, p);
    ssi_drop(s.bits, 3);
    return ssi_dynamic_tables(s.bits, t) == SSI_OK && !ssi_truncated(s.bits);
}

#ifdef __CUDA_ARCH__
// DEVICE fast path of the Huffman loop in marker mode (the shape of ssi_huff_fast's device loop: one lane walks the
// stream, so the loop is written for instruction count -- bit buffer, a 32-bit word index into the 4-byte aligned
// input and one prefetched word in registers; decode tables through 32-bit shared addresses; the literal run is its own
// inner loop and everything rarer is reached through its exits; bounds are checked once per burst of symbols).
// Returns 1 at the end of the block, 0 when the careful loop has to go on (input or room nearly used up), < 0 on bad data.
__device__ __forceinline__ int dgz_huff_fast_dev(ssi_stream &s, ssi_tables &t, uint16_t *sym, uint32_t &n_io, uint32_t cap,
                                                 uint32_t fresh, bool unknown) {
    ssi_bits &b = s.bits;
    if (b.overrun || b.in_end - b.in < 64 || cap - n_io < 2u * 260u) return 0;
    uint32_t cnt = b.cnt & 7u;
    const uint8_t *in = b.in - (b.cnt >> 3);
    uint64_t buf = b.buf & ((1ull << cnt) - 1ull);
    while ((uintptr_t)in & 3u) { buf |= (uint64_t)(*in++) << cnt; cnt += 8u; }
    const uint32_t *const words = reinterpret_cast<const uint32_t *>(in);
    const uint32_t n_words = (uint32_t)((b.in_end - in) >> 2);
    uint32_t wi = 0, n = n_io;
    const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(t.lit), dist_sa = (uint32_t)__cvta_generic_to_shared(t.dist);
    const uint32_t LM4 = ((1u << SSI_LIT_BITS) - 1u) << 2, DM4 = ((1u << SSI_DIST_BITS) - 1u) << 2;
    uint32_t nextw = words[0];
    int ret = 0;
#define DGZ_LDS(v, addr) asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr))
#define DGZ_REFILL() do { if (cnt < 32u) { buf |= (uint64_t)nextw << cnt; cnt += 32u; nextw = words[++wi]; } } while (0)
    uint32_t burst = 0;
    while (ret == 0) {
        if (burst == 0) {
            const uint32_t words_left = n_words > wi + 8u ? n_words - wi - 8u : 0u;
            burst = words_left >> 1;
            if (cap - n < 2u * 260u) break;
            const uint32_t by_out = (cap - n - 2u * 260u) / 258u + 1u;
            if (by_out < burst) burst = by_out;
            if (burst == 0) break;
            if (burst > 4096u) burst = 4096u;
        }
        DGZ_REFILL();
        uint32_t e;
        bool more = true;
        do {                                                          // literal run
            DGZ_LDS(e, lit_sa + (((uint32_t)buf << 2) & LM4));
            if (SSI_KIND(e) != SSI_LIT) break;
            buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
            sym[n++] = (uint16_t)SSI_VAL(e);
            more = --burst != 0 && cnt >= 32u;
        } while (more);
        if (!more) continue;
        burst--;
        if (SSI_KIND(e) == SSI_SUB) {
            buf >>= SSI_LIT_BITS; cnt -= SSI_LIT_BITS;
            DGZ_LDS(e, lit_sa + ((SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u))) << 2));
        }
        buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
        if (SSI_KIND(e) == SSI_LIT) { sym[n++] = (uint16_t)SSI_VAL(e); continue; }
        if (SSI_KIND(e) != SSI_BASE) { ret = SSI_KIND(e) == SSI_EOB ? 1 : SSI_ERR_DATA; break; }
        const uint32_t len = SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u));
        buf >>= SSI_EXTRA(e); cnt -= SSI_EXTRA(e);
        DGZ_REFILL();
        uint32_t d;
        DGZ_LDS(d, dist_sa + (((uint32_t)buf << 2) & DM4));
        if (SSI_KIND(d) == SSI_SUB) {
            buf >>= SSI_DIST_BITS; cnt -= SSI_DIST_BITS;
            DGZ_LDS(d, dist_sa + ((SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u))) << 2));
        }
        buf >>= SSI_LEN(d); cnt -= SSI_LEN(d);
        if (SSI_KIND(d) != SSI_BASE) { ret = SSI_ERR_DATA; break; }
        const uint32_t dist = SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u));
        buf >>= SSI_EXTRA(d); cnt -= SSI_EXTRA(d);
        if (dist > SS_DGZ_WINDOW) { ret = SSI_ERR_DATA; break; }
        uint16_t *dst = sym + n;
        if (dist <= n - fresh) {
            const uint16_t *src = dst - dist;
            if (dist >= len) {                                        // no overlap: the loads do not wait for the stores
                // four symbols per round, the last round predicated (DNA text at low compression levels is mostly
                // matches of 3-6 symbols: a separate tail loop cost more than the copies themselves)
                for (uint32_t i = 0; i < len; i += 4) {
                    const uint32_t left = len - i;
                    const uint16_t a0 = src[i];
                    const uint16_t a1 = left > 1 ? src[i + 1] : (uint16_t)0;
                    const uint16_t a2 = left > 2 ? src[i + 2] : (uint16_t)0;
                    const uint16_t a3 = left > 3 ? src[i + 3] : (uint16_t)0;
                    dst[i] = a0;
                    if (left > 1) dst[i + 1] = a1;
                    if (left > 2) dst[i + 2] = a2;
                    if (left > 3) dst[i + 3] = a3;
                }
            } else {
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
            }
        } else {
            if (fresh != 0 || !unknown) { ret = SSI_ERR_DATA; break; }
            for (uint32_t i = 0; i < len; i++) {
                const int32_t idx = (int32_t)(n + i) - (int32_t)dist;
                dst[i] = idx >= 0 ? sym[idx] : (uint16_t)(256 + (int32_t)SS_DGZ_WINDOW + idx);
            }
        }
        n += len;
    }
#undef DGZ_REFILL
#undef DGZ_LDS
    b.in = reinterpret_cast<const uint8_t *>(words + wi); b.buf = buf; b.cnt = cnt;
    n_io = n;
    return ret;
}
#endif

// One deflate block in marker mode.  sym[0, n) are the symbols this piece has produced so far, `fresh` the index of
// the first symbol behind a member start (references may not reach in front of it; 0 with unknown = true means the
// window in front of sym[0] is unknown and reaches 32 KiB back).  Returns 0 (block done), 1 (out of room: nothing
// of this block counts, the caller stops at the block's start) or an SSI_ERR_* code.
template <bool DGZ_TABLES_IN_SHARED = true>
SS_HD int dgz_block(ssi_stream &s, ssi_tables &t, uint16_t *sym, uint32_t &n_io, uint32_t cap, uint32_t fresh, bool unknown) {
    ssi_bits &b = s.bits;
    ssi_refill(b);
    s.last_block = (int)ssi_take(b, 1);
    const uint32_t type = ssi_take(b, 2);
    uint32_t n = n_io;
    if (type == 0) {
        ssi_drop(b, b.cnt & 7u);
        ssi_refill(b);
        uint32_t len = ssi_take(b, 16), nlen = ssi_take(b, 16);
        if ((len ^ nlen) != 0xFFFFu) return ssi_truncated(b) ? SSI_ERR_TRUNC : SSI_ERR_DATA;
        if (cap - n < len) return 1;
        for (uint32_t i = 0; i < len; i++) {
            if ((b.cnt >> 3) <= b.overrun) ssi_refill(b);
            sym[n++] = (uint16_t)ssi_take(b, 8);
        }
        if (ssi_truncated(b)) return SSI_ERR_TRUNC;
        n_io = n;
        return 0;
    }
    if (type == 1) { if (ssi_fixed_tables(t)) return SSI_ERR_DATA; }
    else if (type == 2) { int rc = ssi_dynamic_tables(b, t); if (rc) return rc; }
    else return ssi_truncated(b) ? SSI_ERR_TRUNC : SSI_ERR_DATA;
    bool eob = false;
#ifdef __CUDA_ARCH__
    if (DGZ_TABLES_IN_SHARED) {
        const int fr = dgz_huff_fast_dev(s, t, sym, n, cap, fresh, unknown);
        if (fr < 0) return fr;
        eob = fr == 1;
    }
#endif
    while (!eob) {
        if (cap - n < 260u) return 1;
        ssi_refill(b);
        uint32_t e = t.lit[ssi_peek(b, SSI_LIT_BITS)];
        if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_LIT_BITS); e = t.lit[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
        ssi_drop(b, SSI_LEN(e));
        uint32_t kind = SSI_KIND(e);
        if (kind == SSI_LIT) {
            sym[n++] = (uint16_t)SSI_VAL(e);
            e = t.lit[ssi_peek(b, SSI_LIT_BITS)];                  // a second and a third literal from the same refill
            if (SSI_KIND(e) == SSI_LIT) {
                ssi_drop(b, SSI_LEN(e)); sym[n++] = (uint16_t)SSI_VAL(e);
                e = t.lit[ssi_peek(b, SSI_LIT_BITS)];
                if (SSI_KIND(e) == SSI_LIT) { ssi_drop(b, SSI_LEN(e)); sym[n++] = (uint16_t)SSI_VAL(e); }
            }
            continue;
        }
        if (kind == SSI_EOB) break;
        if (kind != SSI_BASE) return SSI_ERR_DATA;
        const uint32_t len = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
        e = t.dist[ssi_peek(b, SSI_DIST_BITS)];
        if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_DIST_BITS); e = t.dist[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
        ssi_drop(b, SSI_LEN(e));
        if (SSI_KIND(e) != SSI_BASE) return SSI_ERR_DATA;
        const uint32_t dist = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
        if (dist > SS_DGZ_WINDOW) return SSI_ERR_DATA;
        if (dist <= n - fresh) {                                   // inside what this piece produced since the member start
            const uint16_t *src = sym + n - dist;
            uint16_t *dst = sym + n;
            for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
        } else {
            if (fresh != 0 || !unknown) return SSI_ERR_DATA;       // reaches in front of a member start
            for (uint32_t i = 0; i < len; i++) {                   // (partly) in the unknown window in front of sym[0]
                const int32_t idx = (int32_t)(n + i) - (int32_t)dist;
                sym[n + i] = idx >= 0 ? sym[idx] : (uint16_t)(256 + (int32_t)SS_DGZ_WINDOW + idx);
            }
        }
        n += len;
        if (ssi_truncated(b)) return SSI_ERR_TRUNC;
    }
    if (ssi_truncated(b)) return SSI_ERR_TRUNC;
    n_io = n;
    return 0;
}

// Decode piece `me` of a batch: from its start, block after block, until the decoder stands on the start of a later
// piece, at or behind `limit_bit`, at the end of the stream, or out of room.  A member that ends with the next one
// starting at or behind byte `stop_byte` ends the stream for this decoder (member-split files: the next member
// belongs to another rank).  The window in front of the piece is unknown to the decoder (markers); a batch that
// opens a member has an EMPTY known window, against which K9 rejects any marker.
// `comp_size` bytes are readable now; the stream really ends at `true_size` (>= comp_size: the rest is still being
// uploaded, and whatever would need it is reported as an error for the caller to retry, never guessed).
SS_HD void dgz_decode_piece(const uint8_t *comp, size_t comp_size, size_t true_size, dgz_piece *pieces, uint32_t n_pieces, uint32_t me,
                            uint64_t limit_bit, uint64_t stop_byte, uint16_t *sym, uint32_t cap, ssi_tables &t) {
    dgz_piece &pc = pieces[me];
    pc.n_sym = 0; pc.next = n_pieces; pc.members = 0; pc.fresh_from = 0xFFFFFFFFu; pc.end_bit = pc.start_bit;
    if (pc.start_bit == ~0ull) { pc.status = SS_DGZ_NOSTART; return; }
    ssi_stream s;
    dgz_seek(s, comp, comp_size, pc.start_bit);
    uint32_t n = 0, nx = me + 1, fresh = 0;
    bool unknown = true;
    while (true) {
        const uint64_t at = ssi_bitpos(s);
        pc.end_bit = at; pc.n_sym = n;
        // standing on a block boundary: somebody else's start, or the limit?
        while (nx < n_pieces && (pieces[nx].start_bit == ~0ull || pieces[nx].start_bit < at)) nx++;
        if (at != pc.start_bit && nx < n_pieces && pieces[nx].start_bit == at) { pc.next = nx; pc.status = SS_DGZ_LINKED; return; }
        if (at != pc.start_bit && at >= limit_bit) { pc.next = n_pieces; pc.status = SS_DGZ_LINKED; return; }
        int rc = dgz_block(s, t, sym, n, cap, fresh, unknown);
        if (rc == 1) { pc.status = SS_DGZ_FULL; return; }
        if (rc < 0) { pc.status = SS_DGZ_ERROR; return; }
        if (s.last_block) {
            // member end: trailer (CRC32, ISIZE), then another member or the end of the stream
            ssi_drop(s.bits, s.bits.cnt & 7u);
            const uint8_t *q = ssi_in_pos(s.bits);
            pc.members++;
            pc.end_bit = (uint64_t)(q - comp) * 8u; pc.n_sym = n;
            if (comp + comp_size - q < 8) { pc.status = SS_DGZ_ERROR; return; }
            q += 8;
            ssi_gz_header h;
            const bool at_end = q >= comp + true_size || (uint64_t)(q - comp) >= stop_byte;
            // (a header that is not all there yet is nobody's "trailing garbage")
            if (!at_end && comp_size < true_size && (size_t)(comp + comp_size - q) < 4096) { pc.status = SS_DGZ_ERROR; return; }
            if (at_end || ssi_gz_parse_header(q, comp + comp_size, &h) != SSI_OK) {
                pc.end_bit = (uint64_t)(q - comp) * 8u;
                pc.status = SS_DGZ_END;
                return;
            }
            dgz_seek(s, comp, comp_size, (uint64_t)(q + h.header_len - comp) * 8u);
            fresh = n; unknown = false;
            if (pc.fresh_from == 0xFFFFFFFFu) pc.fresh_from = n;
        }
    }
}

// the window in front of the NEXT accepted piece, given the window in front of this one: the last 32 KiB of
// (win | resolved symbols).  One thread's share: positions [lo, hi) of the new window.  win_len <= 32 KiB bytes are
// right-aligned in `win` (win[32768 - win_len .. 32768)); returns false for a reference in front of the stream start.
SS_HD bool dgz_next_window_part(const uint8_t *win, uint32_t win_len, const uint16_t *sym, uint32_t n_sym,
                                uint8_t *win_out, uint32_t lo, uint32_t hi) {
    bool ok = true;
    for (uint32_t i = lo; i < hi; i++) {
        // position i of the new window = element (win_len32k + n_sym - 32768 + i) of (win | sym), win counted as 32 KiB
        const int64_t e = (int64_t)n_sym - (int64_t)SS_DGZ_WINDOW + (int64_t)i;    // index into sym, negative = old window
        uint8_t v = 0;
        if (e >= 0) {
            const uint16_t x = sym[e];
            if (x < 256) v = (uint8_t)x;
            else {
                const uint32_t off = (uint32_t)x - 256u;
                if (off + win_len < SS_DGZ_WINDOW) ok = false; else v = win[off];
            }
        } else {
            const uint32_t off = (uint32_t)((int64_t)SS_DGZ_WINDOW + e);
            v = off + win_len < SS_DGZ_WINDOW ? 0 : win[off];         // in front of the stream start: never referenced
        }
        win_out[i] = v;
    }
    return ok;
}

SS_HD uint8_t dgz_resolve1(uint16_t x, const uint8_t *win) { return x < 256 ? (uint8_t)x : win[(uint32_t)x - 256u]; }
