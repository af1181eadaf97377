// ss_dgz2.cuh -- K8 of the device gzip inflate (ss_dgz.cuh) for SEVERAL decoders per warp; the form that ships.
// Replaces the `zcat` of library/identify.py:82 / Vote_Strain_L2_Lasso_new_sp.py:359,367 together with K7, K9, K10.
//
// A Huffman decoder is one serial bit chain, so K8's throughput is decoders in flight x the rate of one.  With one
// decoder per warp (ss_dgz.cuh: dgz_decode_piece) the register file allows 24 per SM and every instruction issued
// works for one lane.  Here a few lanes of a warp decode one piece each (measured best: 2 lanes x 24 warps per SM,
// DESIGN.md section 3b has the table):
//   - 16-bit decode-table entries and a 9-bit literal index: 2.5 KB per decoder instead of 6.9 KB (up to 92 decoders
//     per SM fit the shared memory);
//   - the decoder is a state machine (wants a piece / on a block boundary / inside a Huffman block), and the part that
//     is the work -- the symbols of a Huffman block -- runs in ROUNDS of a fixed number of iterations that the lanes
//     of a warp enter together: inside a round the lanes follow the same instruction stream (an instruction issued
//     serves all of them), everything rare (block headers, table building, stored blocks, the last bytes of the
//     input, piece ends) happens between rounds, lane by lane;
//   - an iteration takes up to two tokens: a literal, then a literal or a match;
//   - optionally (measured slower, off by default) a ring of the decoder's last 1024 symbols in shared memory that
//     match copies read instead of the symbol buffer in HBM / L2.
// Same contract as dgz_decode_piece: the pieces' status / end_bit / n_sym / next / members / fresh_from come out
// identical (tests/test_dgz.py runs both on the CPU against zlib).  SS_HD like the rest: the CPU runs the lanes one
// after the other.
#pragma once
#include "ss_dgz.cuh"

#define DGC_LIT_BITS 9
#define DGC_LIT_CAP 852                  // zlib's ENOUGH for 286 symbols, 9 index bits, 15-bit codes
#define DGC_DIST_BITS 8
#define DGC_DIST_CAP 402                 // 32 symbols, 8 index bits (the code-length code, 7 index bits, borrows it)
#define DGC_PRE_BITS 7

// entry = val << 6 | kind << 4 | len
//   LIT      val = the byte (code-length code: the symbol), len = code bits this step consumes
//   SYM      val = length code - 257 / distance code; base and extra bits come from dgc_base_entry's table
//   SUB      val = offset of the sub-table, len = its index bits (the step itself consumes the main index bits)
//   SPECIAL  val 0 = invalid code, 1 = end of block (len = code bits), 2 = (while building) longest code under a prefix
enum { DGC_LIT = 0, DGC_SYM = 1, DGC_SUB = 2, DGC_SPECIAL = 3 };
#define DGC_ENTRY(val, kind, len) ((uint16_t)(((uint32_t)(val) << 6) | ((uint32_t)(kind) << 4) | (uint32_t)(len)))
#define DGC_LEN(e) ((uint32_t)(e) & 15u)
#define DGC_KIND(e) (((uint32_t)(e) >> 4) & 3u)
#define DGC_VAL(e) ((uint32_t)(e) >> 6)
#define DGC_TAG(e) ((uint32_t)(e) >> 4)                      // val and kind together
#define DGC_TAG_BAD 3u
#define DGC_TAG_EOB 7u
#define DGC_TAG_TMP 11u

struct dgz_ctables {
    uint16_t lit[DGC_LIT_CAP];
    uint16_t dist[DGC_DIST_CAP];
};

// base value | extra bits << 16 of length code 257 + i (i < 29) and of distance code i - 32 (32 <= i < 62)
SS_HD uint32_t dgc_base_entry(uint32_t i) {
    if (i < 29u) {
        if (i < 8u) return 3u + i;
        if (i == 28u) return 258u;
        const uint32_t e = (i - 4u) >> 2;
        return (3u + ((4u + (i & 3u)) << e)) | (e << 16);
    }
    if (i < 32u || i >= 62u) return 0;
    const uint32_t s = i - 32u;
    if (s < 4u) return 1u + s;
    const uint32_t e = (s >> 1) - 1u;
    return (1u + ((2u + (s & 1u)) << e)) | (e << 16);
}

SS_HD uint16_t dgc_symbol_entry(int which, int sym, int len) {
    if (which == SSI_CODE_PRE) return DGC_ENTRY(sym, DGC_LIT, len);
    if (which == SSI_CODE_LIT) {
        if (sym < 256) return DGC_ENTRY(sym, DGC_LIT, len);
        if (sym == 256) return DGC_ENTRY(1, DGC_SPECIAL, len);
        if (sym < 286) return DGC_ENTRY(sym - 257, DGC_SYM, len);
        return DGC_ENTRY(0, DGC_SPECIAL, len);                  // 286, 287: only in the fixed code, never valid
    }
    if (sym < 30) return DGC_ENTRY(sym, DGC_SYM, len);
    return DGC_ENTRY(0, DGC_SPECIAL, len);                      // 30, 31
}

// ssi_build_table for 16-bit entries (same rules: over-subscribed and incomplete codes are refused, a single code of
// length 1 and an empty distance code are legal).
SS_HD int dgc_build_table(uint16_t *table, int cap, const uint8_t *lens, int n, int tb, int which) {
    int count[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i] & 15]++;
    const int main_size = 1 << tb;
    for (int i = 0; i < main_size; i++) table[i] = DGC_ENTRY(0, DGC_SPECIAL, 1);
    if (count[0] == n) return which == SSI_CODE_DIST ? 0 : -1;
    int left = 1, maxlen = 0;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
        if (count[l]) maxlen = l;
    }
    if (left > 0 && (which == SSI_CODE_PRE || maxlen != 1)) return -1;
    uint32_t first[16], nx[16];
    {
        uint32_t code = 0;
        first[0] = 0;
        for (int l = 1; l < 16; l++) { code = (code + (uint32_t)(l > 1 ? count[l - 1] : 0)) << 1; first[l] = code; }
    }
    const uint32_t mmask = (uint32_t)main_size - 1u;
    for (int l = 0; l < 16; l++) nx[l] = first[l];
    for (int s = 0; s < n; s++) {                                   // codes that fit the main index; longer ones leave
        const int l = lens[s] & 15;                                 // their longest length under their prefix
        if (!l) continue;
        const uint32_t rev = ssi_bitrev(nx[l]++, l);
        if (l <= tb) {
            const uint16_t e = dgc_symbol_entry(which, s, l);
            for (uint32_t j = rev; j < (uint32_t)main_size; j += 1u << l) table[j] = e;
        } else {
            const uint32_t p = rev & mmask, cur = table[p];
            const int m = (DGC_TAG(cur) == DGC_TAG_TMP && (int)DGC_LEN(cur) > l) ? (int)DGC_LEN(cur) : l;
            table[p] = DGC_ENTRY(2, DGC_SPECIAL, m);
        }
    }
    if (maxlen <= tb) return 0;
    int alloc = main_size;
    for (int l = 0; l < 16; l++) nx[l] = first[l];
    for (int s = 0; s < n; s++) {                                   // one sub-table per long prefix
        const int l = lens[s] & 15;
        if (!l) continue;
        const uint32_t rev = ssi_bitrev(nx[l]++, l);
        if (l <= tb) continue;
        const uint32_t p = rev & mmask;
        uint32_t cur = table[p];
        if (DGC_TAG(cur) == DGC_TAG_TMP) {
            const int sb = (int)DGC_LEN(cur) - tb;
            if (alloc + (1 << sb) > cap) return -1;
            cur = DGC_ENTRY(alloc, DGC_SUB, sb);
            table[p] = (uint16_t)cur;
            for (int j = 0; j < (1 << sb); j++) table[alloc + j] = DGC_ENTRY(0, DGC_SPECIAL, 1);
            alloc += 1 << sb;
        }
        const uint32_t sub = DGC_VAL(cur), sb = DGC_LEN(cur);
        const uint16_t e = dgc_symbol_entry(which, s, l - tb);
        for (uint32_t j = rev >> tb; j < (1u << sb); j += 1u << (l - tb)) table[sub + j] = e;
    }
    return 0;
}

SS_HD int dgc_fixed_tables(dgz_ctables &t) {
    uint8_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    if (dgc_build_table(t.lit, DGC_LIT_CAP, lens, 288, DGC_LIT_BITS, SSI_CODE_LIT)) return -1;
    for (int i = 0; i < 32; i++) lens[i] = 5;
    return dgc_build_table(t.dist, DGC_DIST_CAP, lens, 32, DGC_DIST_BITS, SSI_CODE_DIST);
}

// dynamic block header (RFC 1951 3.2.7): HLIT / HDIST / HCLEN, the code-length code, the two codes
SS_HD int dgc_dynamic_tables(ssi_bits &b, dgz_ctables &t) {
    ssi_refill(b);
    const int hlit = (int)ssi_take(b, 5) + 257, hdist = (int)ssi_take(b, 5) + 1, hclen = (int)ssi_take(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return SSI_ERR_DATA;
    uint8_t lens[320];
    {
        uint8_t pre[19];
        for (int i = 0; i < 19; i++) pre[i] = 0;
        for (int i = 0; i < hclen; i++) {
            // order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, five bits each
            const uint64_t lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                                (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
            const uint64_t hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
            const int sym = (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
            if ((i & 7) == 0) ssi_refill(b);
            pre[sym] = (uint8_t)ssi_take(b, 3);
        }
        if (dgc_build_table(t.dist, DGC_DIST_CAP, pre, 19, DGC_PRE_BITS, SSI_CODE_PRE)) return SSI_ERR_DATA;
    }
    int n = 0;
    const int total = hlit + hdist;
    while (n < total) {
        ssi_refill(b);
        const uint32_t e = t.dist[ssi_peek(b, DGC_PRE_BITS)];
        if (DGC_KIND(e) != DGC_LIT) return SSI_ERR_DATA;
        ssi_drop(b, DGC_LEN(e));
        const int sym = (int)DGC_VAL(e);
        if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (n == 0) return SSI_ERR_DATA;
            val = lens[n - 1];
            rep = 3 + (int)ssi_take(b, 2);
        } else if (sym == 17) rep = 3 + (int)ssi_take(b, 3);
        else rep = 11 + (int)ssi_take(b, 7);
        if (n + rep > total) return SSI_ERR_DATA;
        while (rep--) lens[n++] = (uint8_t)val;
    }
    if (ssi_truncated(b)) return SSI_ERR_TRUNC;
    if (lens[256] == 0) return SSI_ERR_DATA;
    if (dgc_build_table(t.lit, DGC_LIT_CAP, lens, hlit, DGC_LIT_BITS, SSI_CODE_LIT)) return SSI_ERR_DATA;
    if (dgc_build_table(t.dist, DGC_DIST_CAP, lens + hlit, hdist, DGC_DIST_BITS, SSI_CODE_DIST)) return SSI_ERR_DATA;
    return SSI_OK;
}

// ---------------------------------------------------------------------------------------------
// one decoder
// ---------------------------------------------------------------------------------------------
enum { DGZ2_IDLE = 0,        // wants a piece
       DGZ2_BOUNDARY = 1,    // stands on a block boundary of its piece
       DGZ2_HUFF = 2,        // inside a Huffman block, tables built
       DGZ2_DONE = 3 };      // no pieces left

struct dgz2_lane {
    ssi_stream s;
    uint16_t *sym;
    uint32_t me, n, nx, fresh;
    uint32_t ring_from;      // symbols [ring_from, n) are in the decoder's ring (dgz2_round)
    int mode;
    bool unknown;
};

struct dgz2_job {            // what all decoders of a batch share
    const uint8_t *comp;
    size_t comp_size, true_size;
    dgz_piece *pieces;
    uint32_t n_pieces;
    uint64_t limit_bit, stop_byte;
    uint16_t *sym_pool;
    uint32_t cap;
    uint32_t rounds;         // iterations per round
};

#define DGZ2_ITER_SYMBOLS 259u           // an iteration writes at most a literal and a 258-symbol match
#define DGZ2_ITER_WORDS 3u               // ... and loads at most three 32-bit words of input

SS_HD void dgz2_begin_piece(dgz2_lane &L, const dgz2_job &J, uint32_t me) {
    dgz_piece &pc = J.pieces[me];
    pc.n_sym = 0; pc.next = J.n_pieces; pc.members = 0; pc.fresh_from = 0xFFFFFFFFu; pc.end_bit = pc.start_bit;
    L.me = me;
    if (pc.start_bit == ~0ull) { pc.status = SS_DGZ_NOSTART; L.mode = DGZ2_IDLE; return; }
    dgz_seek(L.s, J.comp, J.comp_size, pc.start_bit);
    L.sym = J.sym_pool + (uint64_t)me * J.cap;
    L.n = 0; L.nx = me + 1; L.fresh = 0; L.unknown = true; L.ring_from = 0;
    L.mode = DGZ2_BOUNDARY;
}

SS_HD void dgz2_finish(dgz2_lane &L, const dgz2_job &J, int status) { J.pieces[L.me].status = status; L.mode = DGZ2_IDLE; }

// The end-of-block code has been consumed.
SS_HD void dgz2_after_block(dgz2_lane &L, const dgz2_job &J) {
    dgz_piece &pc = J.pieces[L.me];
    ssi_stream &s = L.s;
    if (ssi_truncated(s.bits)) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
    L.mode = DGZ2_BOUNDARY;
    if (!s.last_block) return;
    // member end: trailer (CRC32, ISIZE), then another member or the end of the stream (as dgz_decode_piece)
    ssi_drop(s.bits, s.bits.cnt & 7u);
    const uint8_t *q = ssi_in_pos(s.bits);
    pc.members++;
    pc.end_bit = (uint64_t)(q - J.comp) * 8u; pc.n_sym = L.n;
    if (J.comp + J.comp_size - q < 8) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
    q += 8;
    ssi_gz_header h;
    const bool at_end = q >= J.comp + J.true_size || (uint64_t)(q - J.comp) >= J.stop_byte;
    if (!at_end && J.comp_size < J.true_size && (size_t)(J.comp + J.comp_size - q) < 4096) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
    if (at_end || ssi_gz_parse_header(q, J.comp + J.comp_size, &h) != SSI_OK) {
        pc.end_bit = (uint64_t)(q - J.comp) * 8u;
        dgz2_finish(L, J, SS_DGZ_END);
        return;
    }
    dgz_seek(s, J.comp, J.comp_size, (uint64_t)(q + h.header_len - J.comp) * 8u);
    L.fresh = L.n; L.unknown = false;
    if (pc.fresh_from == 0xFFFFFFFFu) pc.fresh_from = L.n;
}

// copy of a match whose source may lie (partly) in the unknown window in front of the piece: markers
SS_HD void dgz2_copy_marked(uint16_t *sym, uint32_t n, uint32_t dist, uint32_t len) {
    for (uint32_t i = 0; i < len; i++) {
        const int32_t idx = (int32_t)(n + i) - (int32_t)dist;
        sym[n + i] = idx >= 0 ? sym[idx] : (uint16_t)(256 + (int32_t)SS_DGZ_WINDOW + idx);
    }
}

// On a block boundary: linked to a later piece / at the limit?  Otherwise the block header (a stored block is copied
// here and now; a Huffman block gets its tables).
SS_HD void dgz2_boundary(dgz2_lane &L, const dgz2_job &J, dgz_ctables &t) {
    dgz_piece &pc = J.pieces[L.me];
    ssi_stream &s = L.s;
    const uint64_t at = ssi_bitpos(s);
    pc.end_bit = at; pc.n_sym = L.n;
    while (L.nx < J.n_pieces && (J.pieces[L.nx].start_bit == ~0ull || J.pieces[L.nx].start_bit < at)) L.nx++;
    if (at != pc.start_bit && L.nx < J.n_pieces && J.pieces[L.nx].start_bit == at) { pc.next = L.nx; dgz2_finish(L, J, SS_DGZ_LINKED); return; }
    if (at != pc.start_bit && at >= J.limit_bit) { pc.next = J.n_pieces; dgz2_finish(L, J, SS_DGZ_LINKED); return; }
    ssi_bits &b = s.bits;
    ssi_refill(b);
    s.last_block = (int)ssi_take(b, 1);
    const uint32_t type = ssi_take(b, 2);
    if (type == 0) {
        ssi_drop(b, b.cnt & 7u);
        ssi_refill(b);
        const uint32_t len = ssi_take(b, 16), nlen = ssi_take(b, 16);
        if ((len ^ nlen) != 0xFFFFu) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
        if (J.cap - L.n < len) { dgz2_finish(L, J, SS_DGZ_FULL); return; }
        uint32_t n = L.n;
        for (uint32_t i = 0; i < len; i++) {
            if ((b.cnt >> 3) <= b.overrun) ssi_refill(b);
            L.sym[n++] = (uint16_t)ssi_take(b, 8);
        }
        L.n = n;
        dgz2_after_block(L, J);
        return;
    }
    int rc = SSI_ERR_DATA;
    if (type == 1) rc = dgc_fixed_tables(t) ? SSI_ERR_DATA : SSI_OK;
    else if (type == 2) rc = dgc_dynamic_tables(b, t);
    if (rc) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
    L.mode = DGZ2_HUFF;
}

// iterations of the round loop this decoder may run without looking at its bounds (0: the careful loop has to finish
// the block -- the input or the symbol room is nearly used up)
SS_HD uint32_t dgz2_budget(const dgz2_lane &L, uint32_t cap) {
    const ssi_bits &b = L.s.bits;
    if (b.overrun || b.in_end - b.in < 64 || cap - L.n < 2u * 260u) return 0;
    const uint32_t n_words = (uint32_t)((b.in_end - b.in) >> 2) - 1u;          // (the word pointer may move up by 3 bytes)
    const uint32_t by_in = (n_words - 4u) / DGZ2_ITER_WORDS, by_out = (cap - L.n - 260u) / DGZ2_ITER_SYMBOLS;
    return by_in < by_out ? by_in : by_out;
}

// the block's symbols one at a time with every bound checked: to the end of the block
SS_HD void dgz2_careful(dgz2_lane &L, const dgz2_job &J, const dgz_ctables &t) {
    ssi_bits &b = L.s.bits;
    uint16_t *sym = L.sym;
    uint32_t n = L.n;
    while (true) {
        if (J.cap - n < 260u) { dgz2_finish(L, J, SS_DGZ_FULL); return; }
        ssi_refill(b);
        uint32_t e = t.lit[ssi_peek(b, DGC_LIT_BITS)];
        if (DGC_KIND(e) == DGC_SUB) { ssi_drop(b, DGC_LIT_BITS); e = t.lit[DGC_VAL(e) + ssi_peek(b, DGC_LEN(e))]; }
        ssi_drop(b, DGC_LEN(e));
        if (DGC_KIND(e) == DGC_LIT) { sym[n++] = (uint16_t)DGC_VAL(e); continue; }
        if (DGC_KIND(e) != DGC_SYM) {
            if (DGC_TAG(e) != DGC_TAG_EOB) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
            break;
        }
        uint32_t info = dgc_base_entry(DGC_VAL(e));
        const uint32_t len = (info & 0xFFFFu) + ssi_take(b, info >> 16);
        e = t.dist[ssi_peek(b, DGC_DIST_BITS)];
        if (DGC_KIND(e) == DGC_SUB) { ssi_drop(b, DGC_DIST_BITS); e = t.dist[DGC_VAL(e) + ssi_peek(b, DGC_LEN(e))]; }
        if (DGC_KIND(e) != DGC_SYM) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
        ssi_drop(b, DGC_LEN(e));
        info = dgc_base_entry(32u + DGC_VAL(e));
        const uint32_t dist = (info & 0xFFFFu) + ssi_take(b, info >> 16);
        if (dist <= n - L.fresh) {
            const uint16_t *src = sym + n - dist;
            for (uint32_t i = 0; i < len; i++) sym[n + i] = src[i];
        } else {
            if (L.fresh != 0 || !L.unknown) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
            dgz2_copy_marked(sym, n, dist, len);
        }
        n += len;
        if (ssi_truncated(b)) { dgz2_finish(L, J, SS_DGZ_ERROR); return; }
    }
    L.n = n;
    dgz2_after_block(L, J);
}

// Everything between rounds: returns the budget (> 0) of a decoder that is ready for a round, or 0 with
// mode == DGZ2_IDLE (the caller hands it a piece: dgz2_begin_piece) or DGZ2_DONE.
SS_HD uint32_t dgz2_advance(dgz2_lane &L, const dgz2_job &J, dgz_ctables &t) {
    while (true) {
        if (L.mode == DGZ2_BOUNDARY) {
            const uint32_t n0 = L.n;
            dgz2_boundary(L, J, t);
            if (L.n != n0) L.ring_from = L.n;                          // a stored block: its symbols are not in the ring
            continue;
        }
        if (L.mode == DGZ2_HUFF) {
            const uint32_t budget = dgz2_budget(L, J.cap);
            if (budget) return budget;
            dgz2_careful(L, J, t);
            L.ring_from = L.n;
            continue;
        }
        return 0;
    }
}

// table reads: 32-bit shared-memory addresses on the device (the generic form costs an address conversion per load).
// The asm statements carry no memory clobber on purpose (a clobber would pin the symbol stores around them): every
// address is computed from the bit buffer, which has moved on since the tables were last written, so no two
// executions that straddle a table build share their operands, and nothing in the round loop writes the tables.
#ifdef __CUDA_ARCH__
#define DGZ2_TABLE(name, ptr) const uint32_t name = (uint32_t)__cvta_generic_to_shared(ptr)
#define DGZ2_LD16(v, tab, idx) asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"((tab) + ((idx) << 1)))
#define DGZ2_LD32(v, tab, idx) asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((tab) + ((idx) << 2)))
#else
#define DGZ2_TABLE(name, ptr) const auto *name = (ptr)
#define DGZ2_LD16(v, tab, idx) ((v) = (tab)[idx])
#define DGZ2_LD32(v, tab, idx) ((v) = (tab)[idx])
#endif

// One round: up to `rounds` iterations of a decoder with `budget` (0: it sits the round out).  All decoders of a warp
// enter together and run the same loop; `any_active(bool)` tells whether any of them still works (a warp vote on the
// device).  Returns 0 (inside the block), 1 (the end-of-block code was consumed) or SSI_ERR_DATA.
//
// RING: the decoder also keeps its last DGZ2_RING symbols in `ring` (shared memory on the device), and a match whose
// source lies in there is copied out of it.  The symbol buffer of a piece lives in HBM / L2 (the windows of all
// decoders in flight are several times the L2), so a copy out of it waits hundreds of cycles for its source -- the
// largest single stall of the kernel -- while deflate's matches mostly reach back a few hundred bytes.  Positions
// [L.ring_from, n) of the ring are valid (symbols produced outside a round do not go through it).
#define DGZ2_RING 1024u
template <bool RING, typename Vote>
SS_HD int dgz2_round(dgz2_lane &L, const dgz_ctables &t, const uint32_t *base_tab, uint16_t *ring, uint32_t budget, uint32_t rounds, Vote any_active) {
    ssi_bits &b = L.s.bits;
    uint32_t cnt = 0, wi = 0, nextw = 0, n = L.n;
    uint64_t buf = 0;
    const uint32_t *words = nullptr;
    if (budget) {
        cnt = b.cnt & 7u;
        const uint8_t *in = b.in - (b.cnt >> 3);
        buf = b.buf & ((1ull << cnt) - 1ull);
        while ((uintptr_t)in & 3u) { buf |= (uint64_t)(*in++) << cnt; cnt += 8u; }
        words = reinterpret_cast<const uint32_t *>(in);
        nextw = words[0];
    }
    DGZ2_TABLE(lit, t.lit);
    DGZ2_TABLE(dst_tab, t.dist);
    DGZ2_TABLE(base, base_tab);
    uint16_t *const sym = L.sym;
    const uint32_t fresh = L.fresh;
    const uint32_t ring_lo = RING ? (fresh > L.ring_from ? fresh : L.ring_from) : 0u;    // the ring serves sources at or behind it
    const uint32_t LM = (1u << DGC_LIT_BITS) - 1u, DM = (1u << DGC_DIST_BITS) - 1u, RM = DGZ2_RING - 1u;
    uint32_t left = budget < rounds ? budget : rounds;
    int ret = 0;
#define DGZ2_REFILL() do { if (cnt < 32u) { buf |= (uint64_t)nextw << cnt; cnt += 32u; nextw = words[++wi]; } } while (0)
#define DGZ2_EMIT(pos, v) do { sym[pos] = (uint16_t)(v); if (RING) ring[(pos) & RM] = (uint16_t)(v); } while (0)
    while (any_active(left != 0)) {
        if (left == 0) continue;
        left--;
        DGZ2_REFILL();
        uint32_t e;
        DGZ2_LD16(e, lit, (uint32_t)buf & LM);
        if (DGC_KIND(e) == DGC_LIT) {                                  // first token a literal: a second one may follow
            buf >>= DGC_LEN(e); cnt -= DGC_LEN(e);
            DGZ2_EMIT(n, DGC_VAL(e)); n++;
            DGZ2_LD16(e, lit, (uint32_t)buf & LM);
            if (DGC_KIND(e) == DGC_LIT) {
                buf >>= DGC_LEN(e); cnt -= DGC_LEN(e);
                DGZ2_EMIT(n, DGC_VAL(e)); n++;
                continue;
            }
            DGZ2_REFILL();
        }
        if (DGC_KIND(e) == DGC_SUB) {
            buf >>= DGC_LIT_BITS; cnt -= DGC_LIT_BITS;
            DGZ2_LD16(e, lit, DGC_VAL(e) + ((uint32_t)buf & ((1u << DGC_LEN(e)) - 1u)));
        }
        buf >>= DGC_LEN(e); cnt -= DGC_LEN(e);
        if (DGC_KIND(e) == DGC_LIT) { DGZ2_EMIT(n, DGC_VAL(e)); n++; continue; }
        if (DGC_KIND(e) != DGC_SYM) { ret = DGC_TAG(e) == DGC_TAG_EOB ? 1 : SSI_ERR_DATA; left = 0; continue; }
        uint32_t info;
        DGZ2_LD32(info, base, DGC_VAL(e));
        const uint32_t len = (info & 0xFFFFu) + ((uint32_t)buf & ((1u << (info >> 16)) - 1u));
        buf >>= (info >> 16); cnt -= (info >> 16);
        DGZ2_REFILL();
        uint32_t d;
        DGZ2_LD16(d, dst_tab, (uint32_t)buf & DM);
        if (DGC_KIND(d) == DGC_SUB) {
            buf >>= DGC_DIST_BITS; cnt -= DGC_DIST_BITS;
            DGZ2_LD16(d, dst_tab, DGC_VAL(d) + ((uint32_t)buf & ((1u << DGC_LEN(d)) - 1u)));
        }
        if (DGC_KIND(d) != DGC_SYM) { ret = SSI_ERR_DATA; left = 0; continue; }
        buf >>= DGC_LEN(d); cnt -= DGC_LEN(d);
        DGZ2_LD32(info, base, 32u + DGC_VAL(d));
        const uint32_t dist = (info & 0xFFFFu) + ((uint32_t)buf & ((1u << (info >> 16)) - 1u));
        buf >>= (info >> 16); cnt -= (info >> 16);
        if (RING && dist <= DGZ2_RING && dist <= n - ring_lo) {
            // out of the ring.  Four symbols per step while neither the source nor the destination run wraps and the
            // step's loads do not need its own stores; one at a time otherwise.
            uint32_t i = 0;
            if (dist >= 4u) {
                for (; i < len; i += 4) {
                    const uint32_t s0 = (n + i - dist) & RM, d0 = (n + i) & RM;
                    if (s0 > RM - 3u || d0 > RM - 3u) break;
                    const uint32_t rest = len - i;
                    const uint16_t a0 = ring[s0];
                    const uint16_t a1 = rest > 1 ? ring[s0 + 1] : (uint16_t)0;
                    const uint16_t a2 = rest > 2 ? ring[s0 + 2] : (uint16_t)0;
                    const uint16_t a3 = rest > 3 ? ring[s0 + 3] : (uint16_t)0;
                    sym[n + i] = a0; ring[d0] = a0;
                    if (rest > 1) { sym[n + i + 1] = a1; ring[d0 + 1] = a1; }
                    if (rest > 2) { sym[n + i + 2] = a2; ring[d0 + 2] = a2; }
                    if (rest > 3) { sym[n + i + 3] = a3; ring[d0 + 3] = a3; }
                }
            }
            for (; i < len; i++) {
                const uint16_t a = ring[(n + i - dist) & RM];
                sym[n + i] = a; ring[(n + i) & RM] = a;
            }
        } else if (dist <= n - fresh) {
            uint16_t *dst = sym + n;
            const uint16_t *src = dst - dist;
            if (dist >= 4u) {
                // four symbols per step, the last step predicated (a step's loads never need its own stores)
                for (uint32_t i = 0; i < len; i += 4) {
                    const uint32_t rest = len - i;
                    const uint16_t a0 = src[i];
                    const uint16_t a1 = rest > 1 ? src[i + 1] : (uint16_t)0;
                    const uint16_t a2 = rest > 2 ? src[i + 2] : (uint16_t)0;
                    const uint16_t a3 = rest > 3 ? src[i + 3] : (uint16_t)0;
                    dst[i] = a0;
                    if (rest > 1) dst[i + 1] = a1;
                    if (rest > 2) dst[i + 2] = a2;
                    if (rest > 3) dst[i + 3] = a3;
                    if (RING) {
                        ring[(n + i) & RM] = a0;
                        if (rest > 1) ring[(n + i + 1) & RM] = a1;
                        if (rest > 2) ring[(n + i + 2) & RM] = a2;
                        if (rest > 3) ring[(n + i + 3) & RM] = a3;
                    }
                }
            } else {
                for (uint32_t i = 0; i < len; i++) { const uint16_t a = src[i]; DGZ2_EMIT(n + i, a); }
            }
        } else {
            if (fresh != 0 || !L.unknown) { ret = SSI_ERR_DATA; left = 0; continue; }
            for (uint32_t i = 0; i < len; i++) {                       // (partly) in the unknown window in front of the piece: markers
                const int32_t idx = (int32_t)(n + i) - (int32_t)dist;
                const uint16_t a = idx >= 0 ? sym[idx] : (uint16_t)(256 + (int32_t)SS_DGZ_WINDOW + idx);
                DGZ2_EMIT(n + i, a);
            }
        }
        n += len;
    }
#undef DGZ2_REFILL
#undef DGZ2_EMIT
    if (budget) {
        b.in = reinterpret_cast<const uint8_t *>(words + wi); b.buf = buf; b.cnt = cnt;
        L.n = n;
    }
    return ret;
}

// what follows a round
SS_HD void dgz2_after_round(dgz2_lane &L, const dgz2_job &J, int rc) {
    if (rc == 1) dgz2_after_block(L, J);
    else if (rc < 0) dgz2_finish(L, J, SS_DGZ_ERROR);
}

struct dgz2_vote_alone { SS_HD bool operator()(bool a) const { return a; } };     // a decoder that shares its warp with nobody

// the CPU form of K8: one decoder takes the pieces one after the other
inline void dgz2_decode_pieces_host(const dgz2_job &J, bool with_ring) {
    dgz_ctables *t = new dgz_ctables;
    uint32_t base_tab[64];
    for (uint32_t i = 0; i < 64; i++) base_tab[i] = dgc_base_entry(i);
    uint16_t *ring = new uint16_t[DGZ2_RING];
    dgz2_lane L;
    L.mode = DGZ2_IDLE;
    uint32_t next = 0;
    while (true) {
        const uint32_t budget = dgz2_advance(L, J, *t);
        if (L.mode == DGZ2_IDLE) {
            if (next >= J.n_pieces) break;
            dgz2_begin_piece(L, J, next++);
            continue;
        }
        const int rc = with_ring ? dgz2_round<true>(L, *t, base_tab, ring, budget, J.rounds, dgz2_vote_alone())
                                 : dgz2_round<false>(L, *t, base_tab, ring, budget, J.rounds, dgz2_vote_alone());
        dgz2_after_round(L, J, rc);
    }
    delete[] ring;
    delete t;
}
