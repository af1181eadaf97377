// ss_inflate.cuh -- DEFLATE (RFC 1951) / gzip (RFC 1952) decoder, ONE code base for two places:
//
//   host   : the ingest threads of ss_ingest (one resumable stream per .gz read file; replaces the
//            `zcat a b |` pipe of library/identify.py:82 and Vote_Strain_L2_Lasso_new_sp.py:359,367)
//   device : ss_inflate_bgzf_kernel (one warp per independent gzip member / BGZF block), so that
//            compressed bytes cross PCIe and the FASTQ text is produced directly in HBM
//
// Table-driven canonical-Huffman decoding with a 64-bit bit buffer: one main-table lookup per
// symbol (sub-tables for codes longer than the main table's index width).  All functions are
// header-only and SS_HD (__host__ __device__); the host-only gzip stream wrapper sits at the end.
//
// zlib is NOT used on the product path; tests/ use Python's zlib/gzip as the checker.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SS_HD __host__ __device__ __forceinline__
#else
#define SS_HD inline
#endif

// ---- decode-table entries ----------------------------------------------------------------------
// entry = val << 16 | extra << 8 | kind << 5 | len
//   len   : code bits this step consumes (sub-table entries: the bits after the main index)
//   kind  : LIT literal byte / BASE length-or-distance base with `extra` extra bits / EOB / SUB pointer / BAD
//   val   : literal, base value, or offset of the sub-table; extra (SUB) = sub-table index bits
enum { SSI_LIT = 0, SSI_BASE = 1, SSI_EOB = 2, SSI_SUB = 3, SSI_BAD = 4, SSI_TMP = 5 };
#define SSI_ENTRY(val, extra, kind, len) (((uint32_t)(val) << 16) | ((uint32_t)(extra) << 8) | ((uint32_t)(kind) << 5) | (uint32_t)(len))
#define SSI_LEN(e) ((e) & 31u)
#define SSI_KIND(e) (((e) >> 5) & 7u)
#define SSI_EXTRA(e) (((e) >> 8) & 31u)
#define SSI_VAL(e) ((e) >> 16)

// table geometry: main index bits and worst-case total entries (zlib's `enough` / libdeflate's
// ENOUGH figures for 288 / 32 / 19 symbols, max code length 15 / 15 / 7)
#ifndef SSI_LIT_BITS
#define SSI_LIT_BITS 10
#endif
#if SSI_LIT_BITS == 11
#define SSI_LIT_CAP 2342
#elif SSI_LIT_BITS == 10
#define SSI_LIT_CAP 1334
#elif SSI_LIT_BITS == 9
#define SSI_LIT_CAP 852
#else
#error "SSI_LIT_BITS must be 9, 10 or 11"
#endif
#define SSI_DIST_BITS 8
#define SSI_DIST_CAP 402
#define SSI_PRE_BITS 7
#define SSI_PRE_CAP 128

enum { SSI_CODE_PRE = 0, SSI_CODE_LIT = 1, SSI_CODE_DIST = 2 };

struct ssi_tables {
    uint32_t lit[SSI_LIT_CAP];
    uint32_t dist[SSI_DIST_CAP];
};

SS_HD uint32_t ssi_symbol_entry(int which, int sym, int len) {
    if (which == SSI_CODE_PRE) return SSI_ENTRY(sym, 0, SSI_LIT, len);
    if (which == SSI_CODE_LIT) {
        if (sym < 256) return SSI_ENTRY(sym, 0, SSI_LIT, len);
        if (sym == 256) return SSI_ENTRY(0, 0, SSI_EOB, len);
        if (sym < 265) return SSI_ENTRY(3 + (sym - 257), 0, SSI_BASE, len);
        if (sym < 285) { int e = (sym - 261) >> 2; return SSI_ENTRY(3 + ((4 + ((sym - 265) & 3)) << e), e, SSI_BASE, len); }
        if (sym == 285) return SSI_ENTRY(258, 0, SSI_BASE, len);
        return SSI_ENTRY(0, 0, SSI_BAD, len);           // 286, 287: only in the fixed code, never valid
    }
    if (sym < 4) return SSI_ENTRY(1 + sym, 0, SSI_BASE, len);
    if (sym < 30) { int e = (sym >> 1) - 1; return SSI_ENTRY(1 + ((2 + (sym & 1)) << e), e, SSI_BASE, len); }
    return SSI_ENTRY(0, 0, SSI_BAD, len);               // 30, 31
}

SS_HD uint32_t ssi_bitrev(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// Build the decode table of one canonical Huffman code.  Returns 0, or -1 for an over-subscribed /
// incomplete code (zlib's rule: incomplete is accepted only for a single code of length 1) or a
// table overflow.  An all-zero distance code is legal (a block of literals only).
SS_HD int ssi_build_table(uint32_t *table, int cap, const uint8_t *lens, int n, int tb, int which) {
    int count[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i] & 15]++;
    const int main_size = 1 << tb;
    for (int i = 0; i < main_size; i++) table[i] = SSI_ENTRY(0, 0, SSI_BAD, 1);
    if (count[0] == n) return which == SSI_CODE_DIST ? 0 : -1;
    int left = 1, maxlen = 0;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
        if (count[l]) maxlen = l;
    }
    if (left > 0 && (which == SSI_CODE_PRE || maxlen != 1)) return -1;
    uint32_t next[16];
    {
        uint32_t code = 0;
        next[0] = 0;
        for (int l = 1; l < 16; l++) { code = (code + (uint32_t)(l > 1 ? count[l - 1] : 0)) << 1; next[l] = code; }
    }
    const uint32_t mmask = (uint32_t)main_size - 1u;
    // pass 1: codes that fit the main index; longer ones leave their maximum length under their prefix
    uint32_t nx[16];
    for (int l = 0; l < 16; l++) nx[l] = next[l];
    for (int s = 0; s < n; s++) {
        int l = lens[s] & 15;
        if (!l) continue;
        uint32_t rev = ssi_bitrev(nx[l]++, l);
        if (l <= tb) {
            uint32_t e = ssi_symbol_entry(which, s, l);
            for (uint32_t j = rev; j < (uint32_t)main_size; j += 1u << l) table[j] = e;
        } else {
            uint32_t p = rev & mmask, cur = table[p];
            int m = (SSI_KIND(cur) == SSI_TMP && (int)SSI_VAL(cur) > l) ? (int)SSI_VAL(cur) : l;
            table[p] = SSI_ENTRY(m, 0, SSI_TMP, 0);
        }
    }
    if (maxlen <= tb) return 0;
    // pass 2: allocate one sub-table per long prefix (sized by its longest code) and fill it
    int alloc = main_size;
    for (int l = 0; l < 16; l++) nx[l] = next[l];
    for (int s = 0; s < n; s++) {
        int l = lens[s] & 15;
        if (!l) continue;
        uint32_t rev = ssi_bitrev(nx[l]++, l);
        if (l <= tb) continue;
        uint32_t p = rev & mmask, cur = table[p];
        if (SSI_KIND(cur) == SSI_TMP) {
            int sb = (int)SSI_VAL(cur) - tb;
            if (alloc + (1 << sb) > cap) return -1;
            cur = SSI_ENTRY(alloc, sb, SSI_SUB, tb);
            table[p] = cur;
            for (int j = 0; j < (1 << sb); j++) table[alloc + j] = SSI_ENTRY(0, 0, SSI_BAD, 1);
            alloc += 1 << sb;
        }
        uint32_t sub = SSI_VAL(cur), sb = SSI_EXTRA(cur);
        uint32_t e = ssi_symbol_entry(which, s, l - tb);
        for (uint32_t j = rev >> tb; j < (1u << sb); j += 1u << (l - tb)) table[sub + j] = e;
    }
    return 0;
}

// ---- bit reader --------------------------------------------------------------------------------
// `in` may be read up to 8 bytes at a time while at least 8 bytes remain; the tail goes bytewise.
// Reading past `in_end` yields zero bits and counts `overrun` (a truncated stream is an error, but
// never an out-of-bounds access).  DEVICE: `in` buffers must be padded by 16 readable bytes.
struct ssi_bits {
    const uint8_t *in, *in_end;
    uint64_t buf;
    uint32_t cnt, overrun;
};

SS_HD uint64_t ssi_load64(const uint8_t *p) {
#ifdef __CUDA_ARCH__
    const uint64_t *a = reinterpret_cast<const uint64_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7u) * 8u;
    uint64_t w0 = a[0];
    if (sh == 0) return w0;
    return (w0 >> sh) | (a[1] << (64u - sh));
#else
    uint64_t w;
    memcpy(&w, p, 8);
    return w;
#endif
}

SS_HD void ssi_refill(ssi_bits &b) {   // afterwards cnt >= 56 (bits past the end are zeros)
    if (b.in_end - b.in >= 8) {
        b.buf |= ssi_load64(b.in) << b.cnt;
        b.in += (63u - b.cnt) >> 3;
        b.cnt |= 56u;
    } else {
        while (b.cnt <= 56u) {
            if (b.in < b.in_end) b.buf |= (uint64_t)(*b.in++) << b.cnt;
            else b.overrun++;
            b.cnt += 8u;
        }
    }
}
// true once bits that were never in the input have been consumed (phantom zero bits past in_end)
SS_HD bool ssi_truncated(const ssi_bits &b) { return b.overrun * 8u > b.cnt; }
SS_HD uint32_t ssi_peek(const ssi_bits &b, uint32_t n) { return (uint32_t)b.buf & ((1u << n) - 1u); }
SS_HD void ssi_drop(ssi_bits &b, uint32_t n) { b.buf >>= n; b.cnt -= n; }
SS_HD uint32_t ssi_take(ssi_bits &b, uint32_t n) { uint32_t v = ssi_peek(b, n); ssi_drop(b, n); return v; }
// bytes of input really consumed (whole bytes still sitting in the bit buffer are given back)
SS_HD const uint8_t *ssi_in_pos(const ssi_bits &b) {
    uint32_t back = b.cnt >> 3;
    return (back > b.overrun) ? b.in - (back - b.overrun) : b.in;
}

// ---- block decoding -----------------------------------------------------------------------------
enum {
    SSI_OK = 0,            // stream / member finished
    SSI_MORE_OUTPUT = 1,   // out of output room: call again with a fresh window (history kept by the caller)
    SSI_STOP = 2,          // stopped at a requested block boundary (ssi_stream::stops / limit_bit)
    SSI_ERR_DATA = -1,     // invalid deflate data
    SSI_ERR_TRUNC = -2,    // input ended inside the stream
    SSI_ERR_HEADER = -3,   // not a gzip member
    SSI_ERR_SIZE = -4      // ISIZE of the trailer disagrees with the bytes produced
};

enum { SSI_PH_BLOCK = 0, SSI_PH_STORED = 1, SSI_PH_HUFF = 2, SSI_PH_DONE = 3 };

struct ssi_stream {
    ssi_bits bits;
    int phase, last_block;
    uint32_t stored_left;
    uint64_t out_total;     // bytes produced since the start of this deflate stream (= member)
    // optional stop points (parallel gzip decoding on the host, ss_pgz.cuh): standing at a block boundary whose
    // bit position (relative to `base`) is one of `stops`, or at the first boundary >= `limit_bit`, the
    // decoder returns SSI_STOP instead of reading the next block header
    const uint8_t *base;
    const uint64_t *stops;
    uint32_t n_stops;
    uint64_t limit_bit;
};

SS_HD void ssi_stream_init(ssi_stream &s, const uint8_t *in, const uint8_t *in_end) {
    s.bits.in = in; s.bits.in_end = in_end; s.bits.buf = 0; s.bits.cnt = 0; s.bits.overrun = 0;
    s.phase = SSI_PH_BLOCK; s.last_block = 0; s.stored_left = 0; s.out_total = 0;
    s.base = nullptr; s.stops = nullptr; s.n_stops = 0; s.limit_bit = 0;
}

// bit position of the next unread bit, relative to s.base (exact while no phantom bits were loaded)
SS_HD uint64_t ssi_bitpos(const ssi_stream &s) {
    return (uint64_t)(s.bits.in - s.base) * 8u - s.bits.cnt + (uint64_t)s.bits.overrun * 8u;
}

SS_HD int ssi_fixed_tables(ssi_tables &t) {
    uint8_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    if (ssi_build_table(t.lit, SSI_LIT_CAP, lens, 288, SSI_LIT_BITS, SSI_CODE_LIT)) return -1;
    for (int i = 0; i < 32; i++) lens[i] = 5;
    return ssi_build_table(t.dist, SSI_DIST_CAP, lens, 32, SSI_DIST_BITS, SSI_CODE_DIST);
}

// dynamic block header: HLIT / HDIST / HCLEN, the code-length code, then the two codes
SS_HD int ssi_dynamic_tables(ssi_bits &b, ssi_tables &t) {
    ssi_refill(b);
    int hlit = (int)ssi_take(b, 5) + 257, hdist = (int)ssi_take(b, 5) + 1, hclen = (int)ssi_take(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return SSI_ERR_DATA;
    uint8_t lens[320];
    {
        uint8_t pre[19];
        for (int i = 0; i < 19; i++) pre[i] = 0;
        for (int i = 0; i < hclen; i++) {
            // permutation 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, packed 5 bits per entry
            const uint64_t lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                                (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
            const uint64_t hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
            int sym = (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
            if ((i & 7) == 0) ssi_refill(b);
            pre[sym] = (uint8_t)ssi_take(b, 3);
        }
        // the code-length code lives in the (still unused) distance table
        if (ssi_build_table(t.dist, SSI_PRE_CAP, pre, 19, SSI_PRE_BITS, SSI_CODE_PRE)) return SSI_ERR_DATA;
    }
    int n = 0;
    const int total = hlit + hdist;
    while (n < total) {
        ssi_refill(b);
        uint32_t e = t.dist[ssi_peek(b, SSI_PRE_BITS)];
        if (SSI_KIND(e) != SSI_LIT) return SSI_ERR_DATA;
        ssi_drop(b, SSI_LEN(e));
        int sym = (int)SSI_VAL(e);
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
    if (lens[256] == 0) return SSI_ERR_DATA;   // no end-of-block code
    // the distance lengths are read out of `lens` before its table (which held the precode) is rebuilt
    if (ssi_build_table(t.lit, SSI_LIT_CAP, lens, hlit, SSI_LIT_BITS, SSI_CODE_LIT)) return SSI_ERR_DATA;
    if (ssi_build_table(t.dist, SSI_DIST_CAP, lens + hlit, hdist, SSI_DIST_BITS, SSI_CODE_DIST)) return SSI_ERR_DATA;
    return SSI_OK;
}

SS_HD void ssi_copy_match(uint8_t *out, uint32_t dist, uint32_t len) {
    const uint8_t *src = out - dist;
#ifndef __CUDA_ARCH__
    if (dist >= 8) {            // 8 bytes at a time; may write up to 7 bytes past out + len (slack reserved)
        uint8_t *end = out + len;
        do { uint64_t w; memcpy(&w, src, 8); memcpy(out, &w, 8); src += 8; out += 8; } while (out < end);
        return;
    }
    if (dist == 1) { memset(out, *src, len); return; }
#endif
    for (uint32_t i = 0; i < len; i++) out[i] = src[i];
}

// output room the Huffman loop wants before it decodes another symbol: the longest match + copy slack
#define SSI_OUT_SLACK (258 + 8)

// Host fast path of the Huffman loop: bit state in registers, branch-free refill, up to three literals
// per refill, the next table entry loaded before the current symbol is retired.  Runs while >= 16 input
// bytes and a full slack of output room remain; the careful loop below finishes the block.
// Returns 1 when the end-of-block code was consumed, 0 to continue in the careful loop, < 0 on bad data.
SS_HD int ssi_huff_fast(ssi_stream &s, const ssi_tables &t, uint8_t **out_io, const uint8_t *call_start, uint8_t *out_end) {
#ifndef __CUDA_ARCH__
    ssi_bits &b = s.bits;
    if (b.overrun || b.in_end - b.in < 16 || out_end - *out_io < 2 * SSI_OUT_SLACK) return 0;
    const uint8_t *in = b.in, *const in_safe = b.in_end - 16;
    uint8_t *out = *out_io, *const out_safe = out_end - 2 * SSI_OUT_SLACK;
    uint64_t buf = b.buf;
    uint32_t cnt = b.cnt;
    const uint32_t LM = (1u << SSI_LIT_BITS) - 1u, DM = (1u << SSI_DIST_BITS) - 1u;
    int ret = 0;
#define SSI_FAST_REFILL() do { uint64_t w_; memcpy(&w_, in, 8); buf |= w_ << cnt; in += (63u - cnt) >> 3; cnt |= 56u; } while (0)
    SSI_FAST_REFILL();
    uint32_t e = t.lit[buf & LM];
    while (in <= in_safe && out <= out_safe) {
        if (SSI_KIND(e) == SSI_SUB) { buf >>= SSI_LIT_BITS; cnt -= SSI_LIT_BITS; e = t.lit[SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u))]; }
        buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
        if (SSI_KIND(e) == SSI_LIT) {
            *out++ = (uint8_t)SSI_VAL(e);
            e = t.lit[buf & LM];
            if (SSI_KIND(e) == SSI_LIT) {
                buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
                *out++ = (uint8_t)SSI_VAL(e);
                e = t.lit[buf & LM];
                if (SSI_KIND(e) == SSI_LIT) {
                    buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
                    *out++ = (uint8_t)SSI_VAL(e);
                    SSI_FAST_REFILL();
                    e = t.lit[buf & LM];
                    continue;
                }
            }
            SSI_FAST_REFILL();       // only adds high bits: the entry already loaded stays valid
            continue;
        }
        if (SSI_KIND(e) != SSI_BASE) {
            if (SSI_KIND(e) == SSI_EOB) { ret = 1; break; }
            ret = SSI_ERR_DATA; break;
        }
        uint32_t len = SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u));
        buf >>= SSI_EXTRA(e); cnt -= SSI_EXTRA(e);
        uint32_t d = t.dist[buf & DM];
        if (SSI_KIND(d) == SSI_SUB) { buf >>= SSI_DIST_BITS; cnt -= SSI_DIST_BITS; d = t.dist[SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u))]; }
        buf >>= SSI_LEN(d); cnt -= SSI_LEN(d);
        if (SSI_KIND(d) != SSI_BASE) { ret = SSI_ERR_DATA; break; }
        uint32_t dist = SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u));
        buf >>= SSI_EXTRA(d); cnt -= SSI_EXTRA(d);
        if (dist > s.out_total + (uint64_t)(out - call_start) || dist > 32768u) { ret = SSI_ERR_DATA; break; }
        SSI_FAST_REFILL();
        e = t.lit[buf & LM];         // in flight while the match is copied
        const uint8_t *src = out - dist;
        uint8_t *end = out + len;
        if (dist >= 8) {
            do { uint64_t w; memcpy(&w, src, 8); memcpy(out, &w, 8); src += 8; out += 8; } while (out < end);
        } else if (dist == 1) {
            memset(out, *src, len);
        } else {
            do { *out++ = *src++; } while (out < end);
        }
        out = end;
    }
#undef SSI_FAST_REFILL
    b.in = in; b.buf = buf; b.cnt = cnt;
    *out_io = out;
    return ret;
#else
    // DEVICE fast path (one lane of a warp walks the stream, ss_gunzip.cu).  A lone lane issues one instruction
    // every ~5 cycles, so the loop is written for instruction count: the bit buffer, a 32-bit word index into
    // the (4-byte aligned) input and one prefetched input word live in registers; the refill and the
    // sub-table step are real, rarely taken branches; the loop bounds are checked once per burst of
    // symbols, not per symbol.  Needs 16 readable bytes behind in_end (the staging buffers are padded).
    ssi_bits &b = s.bits;
    if (b.overrun || b.in_end - b.in < 64 || out_end - *out_io < 2 * SSI_OUT_SLACK) return 0;
    // hand the whole bytes of the bit buffer back to the input, then take single bytes up to a word boundary
    uint32_t cnt = b.cnt & 7u;
    const uint8_t *in = b.in - (b.cnt >> 3);
    uint64_t buf = b.buf & ((1ull << cnt) - 1ull);
    while ((uintptr_t)in & 3u) { buf |= (uint64_t)(*in++) << cnt; cnt += 8u; }
    const uint32_t *const words = reinterpret_cast<const uint32_t *>(in);
    const uint32_t n_words = (uint32_t)((b.in_end - in) >> 2);        // whole words before in_end
    uint32_t wi = 0;                                                  // words[wi] is `nextw`, not yet in the bit buffer
    uint8_t *const out0 = *out_io;
    uint32_t o = 0;                                                   // bytes written by this call: out0[o] is next
    const uint64_t total0 = s.out_total + (uint64_t)(out0 - call_start);   // member bytes produced before out0
    const uint32_t room0 = (uint32_t)((uint64_t)(out_end - out0) > 0x7FFFFFFFull ? 0x7FFFFFFFull : (uint64_t)(out_end - out0));
    // the decode tables live in shared memory (ss_gunzip.cu): 32-bit shared addresses and ld.shared keep the
    // per-symbol address arithmetic at one instruction
    const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(t.lit), dist_sa = (uint32_t)__cvta_generic_to_shared(t.dist);
    const uint32_t LM4 = ((1u << SSI_LIT_BITS) - 1u) << 2, DM4 = ((1u << SSI_DIST_BITS) - 1u) << 2;
    uint32_t nextw = words[0];
    int ret = 0;
#define SSI_LDS(v, addr) asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr))
#define SSI_DEV_REFILL() do { if (cnt < 32u) { buf |= (uint64_t)nextw << cnt; cnt += 32u; nextw = words[++wi]; } } while (0)
    // Loop shape: ptxas predicates every small conditional block, and a lone lane pays for each predicated-off
    // instruction.  So the common case -- a run of literals decoded from the main table while >= 32 bits remain --
    // is its own inner loop, and everything rarer (refill, sub-table, match, end of block, burst bookkeeping)
    // is reached only through that loop's EXITS, which are real branches.
    uint32_t burst = 0;
    while (ret == 0) {
        if (burst == 0) {
            // symbols that can be decoded before a bound could be crossed: each takes <= 2 words of input and
            // writes <= 258 bytes
            uint32_t words_left = n_words > wi + 8u ? n_words - wi - 8u : 0u;
            burst = words_left >> 1;
            if (room0 - o < 2 * SSI_OUT_SLACK) break;
            uint32_t by_out = (room0 - o - 2 * SSI_OUT_SLACK) / 258u + 1u;
            if (by_out < burst) burst = by_out;
            if (burst == 0) break;
            if (burst > 4096u) burst = 4096u;
        }
        SSI_DEV_REFILL();                                             // >= 32 bits: a length code with its extra bits
        uint32_t e;
        bool more = true;
        do {                                                          // literal run
            SSI_LDS(e, lit_sa + (((uint32_t)buf << 2) & LM4));
            if (SSI_KIND(e) != SSI_LIT) break;
            buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
            out0[o++] = (uint8_t)SSI_VAL(e);
            more = --burst != 0 && cnt >= 32u;
        } while (more);
        if (!more) continue;                                          // refill / new burst, then on with the run
        burst--;
        if (SSI_KIND(e) == SSI_SUB) {
            buf >>= SSI_LIT_BITS; cnt -= SSI_LIT_BITS;
            SSI_LDS(e, lit_sa + ((SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u))) << 2));
        }
        buf >>= SSI_LEN(e); cnt -= SSI_LEN(e);
        if (SSI_KIND(e) == SSI_LIT) { out0[o++] = (uint8_t)SSI_VAL(e); continue; }
        if (SSI_KIND(e) != SSI_BASE) { ret = SSI_KIND(e) == SSI_EOB ? 1 : SSI_ERR_DATA; break; }
        uint32_t len = SSI_VAL(e) + ((uint32_t)buf & ((1u << SSI_EXTRA(e)) - 1u));
        buf >>= SSI_EXTRA(e); cnt -= SSI_EXTRA(e);
        SSI_DEV_REFILL();                                             // >= 32 bits: a distance code with its extra bits
        uint32_t d;
        SSI_LDS(d, dist_sa + (((uint32_t)buf << 2) & DM4));
        if (SSI_KIND(d) == SSI_SUB) {
            buf >>= SSI_DIST_BITS; cnt -= SSI_DIST_BITS;
            SSI_LDS(d, dist_sa + ((SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u))) << 2));
        }
        buf >>= SSI_LEN(d); cnt -= SSI_LEN(d);
        if (SSI_KIND(d) != SSI_BASE) { ret = SSI_ERR_DATA; break; }
        uint32_t dist = SSI_VAL(d) + ((uint32_t)buf & ((1u << SSI_EXTRA(d)) - 1u));
        buf >>= SSI_EXTRA(d); cnt -= SSI_EXTRA(d);
        if (dist > total0 + o || dist > 32768u) { ret = SSI_ERR_DATA; break; }
        uint8_t *dst = out0 + o;
        const uint8_t *src = dst - dist;
        if (dist >= len) {                                            // no overlap: the loads do not wait for the stores
            uint32_t i = 0;
            for (; i + 4 <= len; i += 4) {
                uint8_t a0 = src[i], a1 = src[i + 1], a2 = src[i + 2], a3 = src[i + 3];
                dst[i] = a0; dst[i + 1] = a1; dst[i + 2] = a2; dst[i + 3] = a3;
            }
            for (; i < len; i++) dst[i] = src[i];
        } else {
            for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
        }
        o += len;
    }
#undef SSI_DEV_REFILL
#undef SSI_LDS
    uint8_t *out = out0 + o;
    // hand the state back: the prefetched word was not consumed
    b.in = reinterpret_cast<const uint8_t *>(words + wi); b.buf = buf; b.cnt = cnt;
    *out_io = out;
    return ret;
#endif
}

// Decode until the stream ends (SSI_OK), the output window [out, out_end) is (nearly) full
// (SSI_MORE_OUTPUT) or an error.  `*out_pos` advances.  Back-references reach at most 32 KiB behind
// the write position: the caller keeps that much history directly in front of `out` when it hands
// in a new window.
SS_HD int ssi_inflate(ssi_stream &s, ssi_tables &t, uint8_t **out_pos, uint8_t *out_end) {
    uint8_t *out = *out_pos;
    ssi_bits &b = s.bits;
    int rc = SSI_OK;
    while (true) {
        if (s.phase == SSI_PH_BLOCK) {
            if (s.last_block) { s.phase = SSI_PH_DONE; break; }
            if (s.base) {
                const uint64_t at = ssi_bitpos(s);
                bool stop = s.limit_bit && at >= s.limit_bit;
                for (uint32_t i = 0; i < s.n_stops && !stop; i++) stop = s.stops[i] == at;
                if (stop) { rc = SSI_STOP; break; }
            }
            ssi_refill(b);
            s.last_block = (int)ssi_take(b, 1);
            uint32_t type = ssi_take(b, 2);
            if (type == 0) {
                ssi_drop(b, b.cnt & 7u);                   // to the byte boundary
                ssi_refill(b);
                uint32_t len = ssi_take(b, 16), nlen = ssi_take(b, 16);
                if ((len ^ nlen) != 0xFFFFu) { rc = b.overrun ? SSI_ERR_TRUNC : SSI_ERR_DATA; break; }
                s.stored_left = len;
                s.phase = SSI_PH_STORED;
            } else if (type == 1) {
                if (ssi_fixed_tables(t)) { rc = SSI_ERR_DATA; break; }
                s.phase = SSI_PH_HUFF;
            } else if (type == 2) {
                rc = ssi_dynamic_tables(b, t);
                if (rc) break;
                s.phase = SSI_PH_HUFF;
            } else { rc = b.overrun ? SSI_ERR_TRUNC : SSI_ERR_DATA; break; }
        }
        if (s.phase == SSI_PH_STORED) {
            // the bit buffer holds whole bytes here: drain it, then copy straight from the input
            while (s.stored_left && (b.cnt >> 3) > b.overrun && out < out_end) { *out++ = (uint8_t)ssi_take(b, 8); s.stored_left--; }
            if (s.stored_left && (b.cnt >> 3) <= b.overrun) {
                if (b.overrun) { rc = SSI_ERR_TRUNC; break; }
                b.buf = 0; b.cnt = 0;
                size_t room = (size_t)(out_end - out), avail = (size_t)(b.in_end - b.in), n = s.stored_left;
                if (n > room) n = room;
                if (n > avail) n = avail;
                for (size_t i = 0; i < n; i++) out[i] = b.in[i];
                out += n; b.in += n; s.stored_left -= (uint32_t)n;
                if (s.stored_left && b.in >= b.in_end) { rc = SSI_ERR_TRUNC; break; }
            }
            if (s.stored_left) { rc = SSI_MORE_OUTPUT; break; }
            s.phase = SSI_PH_BLOCK;
            continue;
        }
        if (s.phase == SSI_PH_HUFF) {
            bool eob = false;
            {
                int fr = ssi_huff_fast(s, t, &out, *out_pos, out_end);
                if (fr < 0) { rc = fr; break; }
                eob = fr == 1;
            }
            while (!eob && out_end - out >= SSI_OUT_SLACK) {
                ssi_refill(b);
                uint32_t e = t.lit[ssi_peek(b, SSI_LIT_BITS)];
                if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_LIT_BITS); e = t.lit[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
                ssi_drop(b, SSI_LEN(e));
                uint32_t kind = SSI_KIND(e);
                if (kind == SSI_LIT) {
                    *out++ = (uint8_t)SSI_VAL(e);
                    // a second literal from the same refill (>= 56 - 15 bits are left)
                    e = t.lit[ssi_peek(b, SSI_LIT_BITS)];
                    if (SSI_KIND(e) == SSI_LIT) { ssi_drop(b, SSI_LEN(e)); *out++ = (uint8_t)SSI_VAL(e); }
                    continue;
                }
                if (kind == SSI_EOB) { eob = true; break; }
                if (kind != SSI_BASE) { rc = SSI_ERR_DATA; break; }
                uint32_t len = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
                // <= 20 bits used so far of >= 56: the distance code (<= 15 + 13) still fits
                e = t.dist[ssi_peek(b, SSI_DIST_BITS)];
                if (SSI_KIND(e) == SSI_SUB) { ssi_drop(b, SSI_DIST_BITS); e = t.dist[SSI_VAL(e) + ssi_peek(b, SSI_EXTRA(e))]; }
                ssi_drop(b, SSI_LEN(e));
                if (SSI_KIND(e) != SSI_BASE) { rc = SSI_ERR_DATA; break; }
                uint32_t dist = SSI_VAL(e) + ssi_take(b, SSI_EXTRA(e));
                uint64_t produced = s.out_total + (uint64_t)(out - *out_pos);
                if (dist > produced || dist > 32768u) { rc = SSI_ERR_DATA; break; }
                ssi_copy_match(out, dist, len);
                out += len;
            }
            if (rc) break;
            if (ssi_truncated(b)) { rc = SSI_ERR_TRUNC; break; }
            if (!eob) { rc = SSI_MORE_OUTPUT; break; }
            s.phase = SSI_PH_BLOCK;
            continue;
        }
        break;
    }
    s.out_total += (uint64_t)(out - *out_pos);
    *out_pos = out;
    if (rc == SSI_OK && ssi_truncated(b)) rc = SSI_ERR_TRUNC;
    return rc;
}

// ---- gzip member framing (RFC 1952) ---------------------------------------------------------------
struct ssi_gz_header {
    uint32_t header_len;     // bytes before the deflate stream
    uint32_t bgzf_bsize;     // total member size from a BGZF "BC" extra subfield, 0 if absent
};

// Parse one member header at p.  Returns 0, SSI_ERR_HEADER (no gzip magic / unknown method) or SSI_ERR_TRUNC.
SS_HD int ssi_gz_parse_header(const uint8_t *p, const uint8_t *end, ssi_gz_header *h) {
    h->header_len = 0; h->bgzf_bsize = 0;
    if (end - p < 10) return (end - p >= 2 && (p[0] != 0x1f || p[1] != 0x8b)) ? SSI_ERR_HEADER : SSI_ERR_TRUNC;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return SSI_ERR_HEADER;
    const uint8_t flg = p[3];
    const uint8_t *q = p + 10;
    if (flg & 4) {                                     // FEXTRA
        if (end - q < 2) return SSI_ERR_TRUNC;
        uint32_t xlen = (uint32_t)q[0] | ((uint32_t)q[1] << 8);
        q += 2;
        if ((uint32_t)(end - q) < xlen) return SSI_ERR_TRUNC;
        const uint8_t *x = q, *xe = q + xlen;
        while (xe - x >= 4) {
            uint32_t sl = (uint32_t)x[2] | ((uint32_t)x[3] << 8);
            if ((uint32_t)(xe - x - 4) < sl) break;
            if (x[0] == 'B' && x[1] == 'C' && sl == 2) h->bgzf_bsize = ((uint32_t)x[4] | ((uint32_t)x[5] << 8)) + 1u;
            x += 4 + sl;
        }
        q = xe;
    }
    if (flg & 8) { while (q < end && *q) q++; if (q >= end) return SSI_ERR_TRUNC; q++; }    // FNAME
    if (flg & 16) { while (q < end && *q) q++; if (q >= end) return SSI_ERR_TRUNC; q++; }   // FCOMMENT
    if (flg & 2) { if (end - q < 2) return SSI_ERR_TRUNC; q += 2; }                         // FHCRC
    h->header_len = (uint32_t)(q - p);
    return SSI_OK;
}

// ---- resumable multi-member gzip stream over an in-memory (mmap'ed) file: host side -----------------
// Same behaviour as `zcat`: members are concatenated; bytes after a member that do not start another
// member (zero padding included) end the stream and are ignored (zcat: "trailing garbage ignored").  ISIZE is checked, CRC32 is not
// (zcat would only print a warning after the data is already in the pipe).
struct ssi_gz_stream {
    const uint8_t *p, *end;     // next member header / end of the file
    const uint8_t *stop_p;      // optional: standing between two members at or behind this byte, return SSI_OK
    ssi_stream s;
    ssi_tables t;
    int in_member;
    uint64_t n_members;
    uint64_t total_out;
};

inline void ssi_gz_init(ssi_gz_stream &g, const uint8_t *data, size_t len) {
    g.p = data; g.end = data + len; g.in_member = 0; g.n_members = 0; g.total_out = 0; g.stop_p = nullptr;
}

// Fill [*out_pos, out_end).  SSI_OK = the whole file is done, SSI_MORE_OUTPUT = window full (call again
// with a new window whose preceding 32 KiB hold the previous output), < 0 = error.
inline int ssi_gz_read(ssi_gz_stream &g, uint8_t **out_pos, uint8_t *out_end) {
    while (true) {
        if (!g.in_member) {
            if (g.p >= g.end) return g.n_members ? SSI_OK : SSI_ERR_HEADER;
            if (g.stop_p && g.n_members && g.p >= g.stop_p) return SSI_OK;      // the next member belongs to someone else
            ssi_gz_header h;
            int rc = ssi_gz_parse_header(g.p, g.end, &h);
            if (rc == SSI_ERR_HEADER && g.n_members) return SSI_OK;   // trailing garbage
            if (rc) return rc;
            ssi_stream_init(g.s, g.p + h.header_len, g.end);
            g.in_member = 1;
        }
        uint8_t *before = *out_pos;
        int rc = ssi_inflate(g.s, g.t, out_pos, out_end);
        g.total_out += (uint64_t)(*out_pos - before);
        if (rc != SSI_OK) return rc;
        ssi_drop(g.s.bits, g.s.bits.cnt & 7u);
        const uint8_t *q = ssi_in_pos(g.s.bits);
        if (g.end - q < 8) return SSI_ERR_TRUNC;
        uint32_t isize = (uint32_t)q[4] | ((uint32_t)q[5] << 8) | ((uint32_t)q[6] << 16) | ((uint32_t)q[7] << 24);
        if (isize != (uint32_t)g.s.out_total) return SSI_ERR_SIZE;
        g.p = q + 8;
        g.in_member = 0;
        g.n_members++;
    }
}
