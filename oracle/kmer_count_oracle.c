/*
 * oracle/kmer_count_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the match+count step that StrainScan delegates to
 * the bundled `library/jellyfish-linux` binary (Jellyfish 2.3.0; gmarcais/Jellyfish,
 * source NOT vendored under /root/reference, version known from `--version`):
 *
 *     jellyfish count -m K -s 100M -t T --if <kmers.fa> -o X.jf <reads...>
 *     jellyfish dump -c X.jf
 *
 * Reference call sites that this file stands in for:
 *     library/identify.py:82-87            (L1, k fixed to 31, no -C)
 *     library/identify_low_mem.py:74-75
 *     library/identify_low_depth.py:54-59
 *     library/Vote_Strain_L2_Lasso_new_sp.py:357-372   (L2, -m ksize, no -C)
 * and for the record-string view the Python adapters take of the same FASTA:
 *     library/identify.py:90-95            (record i = rstrip(line[2i+1]), .upper())
 *     library/Vote_Strain_L2_Lasso_new_sp.py:386-389 (raw strings, kid order)
 *
 * The algorithm restated here is Jellyfish's published one: sequence files are
 * parsed record by record (FASTA '>' / FASTQ '@', multi-line sequence joined, FASTQ
 * quality consumed by length), every window of K consecutive A/C/G/T bases (case
 * folded) on the FORWARD strand is looked up in the set seeded from --if, and a hit
 * increments that k-mer's counter.  There is no canonicalisation (no -C).  Every
 * k-window of every --if record is seeded (so a record longer than K seeds all of
 * its windows), duplicates collapse, zero-count k-mers are still dumped.
 *
 * Parity pin: tests/golden/ holds count vectors produced by running the real
 * jellyfish-linux binary in the build container (script: tests/golden/make_golden.py);
 * tests/test_oracle.py checks this restatement against every one of them, and
 * against the live binary in oracle/_ref/ when that is present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product (strainscan_b200/) never does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- 2-bit code table: A/a=0 C/c=1 G/g=2 T/t=3, everything else = 4 (breaks the window) */
static uint8_t CODE[256];
static int code_ready = 0;
static void init_code(void) {
    if (code_ready) return;
    memset(CODE, 4, sizeof CODE);
    CODE['A'] = CODE['a'] = 0;
    CODE['C'] = CODE['c'] = 1;
    CODE['G'] = CODE['g'] = 2;
    CODE['T'] = CODE['t'] = 3;
    code_ready = 1;
}

/* ---- seeded set: open addressing over the packed k-mer (first base most significant) */
typedef struct {
    uint64_t *keys;
    uint64_t *cnt;
    uint8_t  *used;   /* every 64-bit value is a legal key at k = 32 (poly-T = all ones), so occupancy is separate */
    uint64_t  cap;    /* power of two */
    uint64_t  n;
} kset;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

static int kset_init(kset *s, uint64_t expect) {
    uint64_t cap = 64;
    while (cap < expect * 2 + 16) cap <<= 1;
    s->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    s->cnt = (uint64_t *)calloc(cap, sizeof(uint64_t));
    s->used = (uint8_t *)calloc(cap, 1);
    if (!s->keys || !s->cnt || !s->used) return -1;
    s->cap = cap; s->n = 0;
    return 0;
}
static void kset_free(kset *s) { free(s->keys); free(s->cnt); free(s->used); }

static int kset_grow(kset *s);

/* insert-if-absent (seeding: count stays 0) */
static int kset_seed(kset *s, uint64_t key) {
    if ((s->n + 1) * 2 > s->cap) { if (kset_grow(s)) return -1; }
    uint64_t m = s->cap - 1, h = mix64(key) & m;
    while (s->used[h]) {
        if (s->keys[h] == key) return 0;
        h = (h + 1) & m;
    }
    s->keys[h] = key; s->used[h] = 1; s->n++;
    return 0;
}
static int kset_grow(kset *s) {
    kset t;
    if (kset_init(&t, s->cap)) return -1;       /* doubles */
    for (uint64_t i = 0; i < s->cap; i++)
        if (s->used[i]) {
            uint64_t m = t.cap - 1, h = mix64(s->keys[i]) & m;
            while (t.used[h]) h = (h + 1) & m;
            t.keys[h] = s->keys[i]; t.cnt[h] = s->cnt[i]; t.used[h] = 1; t.n++;
        }
    kset_free(s);
    *s = t;
    return 0;
}
/* returns slot or -1 */
static int64_t kset_find(const kset *s, uint64_t key) {
    uint64_t m = s->cap - 1, h = mix64(key) & m;
    while (s->used[h]) {
        if (s->keys[h] == key) return (int64_t)h;
        h = (h + 1) & m;
    }
    return -1;
}

/* ---- line cursor over an in-memory file */
typedef struct { const char *p, *end; } cur;
static int cur_eof(const cur *c) { return c->p >= c->end; }
static int cur_peek(const cur *c) { return c->p < c->end ? (unsigned char)*c->p : -1; }
/* returns the line without its '\n' (a trailing '\r' is kept, as std::getline would) */
static void cur_line(cur *c, const char **s, size_t *n) {
    const char *nl = (const char *)memchr(c->p, '\n', (size_t)(c->end - c->p));
    *s = c->p;
    if (nl) { *n = (size_t)(nl - c->p); c->p = nl + 1; }
    else    { *n = (size_t)(c->end - c->p); c->p = c->end; }
}

/* rolling forward-strand k-mer over a stream of chars; state survives line joins */
typedef struct { uint64_t mer, mask; int filled, k; } roll;
static void roll_init(roll *r, int k) {
    r->k = k; r->mer = 0; r->filled = 0;
    r->mask = (k == 32) ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) - 1);
}
static void roll_reset(roll *r) { r->mer = 0; r->filled = 0; }

typedef int (*mer_fn)(void *ctx, uint64_t mer);

static int roll_feed(roll *r, const char *s, size_t n, mer_fn fn, void *ctx) {
    for (size_t i = 0; i < n; i++) {
        uint8_t c = CODE[(unsigned char)s[i]];
        if (c > 3) { r->filled = 0; r->mer = 0; continue; }
        r->mer = ((r->mer << 2) | c) & r->mask;
        if (r->filled < r->k) r->filled++;
        if (r->filled >= r->k) { int e = fn(ctx, r->mer); if (e) return e; }
    }
    return 0;
}

/*
 * Jellyfish-style sequence-file walk.  File type from the first byte: '>' FASTA,
 * '@' FASTQ.  FASTA: header line, then sequence lines until the next line that starts
 * with '>' (joined).  FASTQ: header line, sequence lines until a line starting with '+'
 * (joined), the '+' line is skipped, then quality lines are consumed until at least as
 * many quality chars as sequence chars were read.  Returns 0, or -2 on a malformed file.
 */
static int walk_sequences(const char *buf, size_t len, int k, mer_fn fn, void *ctx) {
    cur c = { buf, buf + len };
    roll r; roll_init(&r, k);
    const char *s; size_t n;
    while (!cur_eof(&c) && (cur_peek(&c) == '\n' || cur_peek(&c) == '\r')) c.p++;  /* leading blanks */
    if (cur_eof(&c)) return 0;
    int type = cur_peek(&c);
    if (type != '>' && type != '@') return -2;
    while (!cur_eof(&c)) {
        if (cur_peek(&c) != type) {
            /* tolerate trailing blank lines only */
            const char *q = c.p; while (q < c.end && (*q == '\n' || *q == '\r')) q++;
            if (q == c.end) return 0;
            return -2;
        }
        cur_line(&c, &s, &n);                     /* header */
        roll_reset(&r);
        size_t seqlen = 0;
        if (type == '>') {
            while (!cur_eof(&c) && cur_peek(&c) != '>') {
                cur_line(&c, &s, &n);
                int e = roll_feed(&r, s, n, fn, ctx); if (e) return e;
            }
        } else {
            while (!cur_eof(&c) && cur_peek(&c) != '+') {
                cur_line(&c, &s, &n); seqlen += n;
                int e = roll_feed(&r, s, n, fn, ctx); if (e) return e;
            }
            if (cur_eof(&c)) return -2;           /* truncated */
            cur_line(&c, &s, &n);                 /* '+' line */
            size_t q = 0;
            while (q < seqlen && !cur_eof(&c)) { cur_line(&c, &s, &n); q += n; }
            if (q != seqlen) return -2;
        }
    }
    return 0;
}

static int seed_cb(void *ctx, uint64_t mer) { return kset_seed((kset *)ctx, mer); }
static int count_cb(void *ctx, uint64_t mer) {
    kset *s = (kset *)ctx;
    int64_t h = kset_find(s, mer);
    if (h >= 0) s->cnt[h]++;
    return 0;
}

/* pack a record string if it is exactly k chars of ACGT after case folding */
static int pack_record(const char *s, size_t n, int k, uint64_t *out, int *raw_upper) {
    if ((int)n != k) return 0;
    uint64_t m = 0; int up = 1;
    for (int i = 0; i < k; i++) {
        uint8_t c = CODE[(unsigned char)s[i]];
        if (c > 3) return 0;
        if (s[i] >= 'a') up = 0;
        m = (m << 2) | c;
    }
    *out = m; *raw_upper = up;
    return 1;
}

static size_t rstrip_len(const char *s, size_t n) {
    while (n > 0 && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n' ||
                     s[n - 1] == '\v' || s[n - 1] == '\f')) n--;
    return n;
}

/*
 * Number of records the Python adapters see in a k-mer FASTA: int(len(lines)/2)
 * (identify.py:92-93).
 */
uint64_t orc_fasta_records(const char *fa, size_t fa_len) {
    uint64_t lines = 0;
    cur c = { fa, fa + fa_len };
    const char *s; size_t n;
    while (!cur_eof(&c)) { cur_line(&c, &s, &n); lines++; }
    return lines / 2;
}

/*
 * The whole step.
 *   fa, fa_len          the --if k-mer FASTA (kmer.fa / all_kmer.fasta), in memory
 *   k                   -m
 *   reads, read_lens    n_files in-memory read files, processed in argv order
 * Outputs, one entry per adapter-view record i (caller allocates n_records each):
 *   cnt[i]        jellyfish's dumped count of upper(r_i), 0 if upper(r_i) is not a dumped key
 *   in_set[i]     1 if upper(r_i) is a dumped key (|r_i| == k, all ACGT after folding)
 *   is_last[i]    1 if no later record has the same upper-cased string (identify.py:94 "last wins")
 *   raw_upper[i]  1 if the raw string r_i itself is a dumped key (pure uppercase ACGT; the L2
 *                 adapter looks raw strings up, Vote_Strain_L2_Lasso_new_sp.py:315)
 *   header_id[i]  integer after '>' on the header line (all_kmer.fasta: kid), 0 if not numeric
 * n_distinct: |S|, the number of lines `jellyfish dump -c` would print.
 * Returns 0, -1 out of memory, -2 malformed read file, -3 bad k.
 */
int orc_count(const char *fa, size_t fa_len, int k,
              const char *const *reads, const size_t *read_lens, int n_files,
              uint64_t n_records, uint64_t *cnt, uint8_t *in_set, uint8_t *is_last,
              uint8_t *raw_upper, uint64_t *header_id, uint64_t *n_distinct) {
    init_code();
    if (k < 1 || k > 32) return -3;
    kset S;
    if (kset_init(&S, n_records + 16)) return -1;
    int rc = 0;

    /* 1. seeding, Jellyfish's view of the FASTA (all k-windows of every record) */
    if (fa_len) {
        rc = walk_sequences(fa, fa_len, k, seed_cb, &S);
        if (rc) { kset_free(&S); return rc == -2 ? -4 : rc; }
    }

    /* 2. read scan */
    for (int f = 0; f < n_files; f++) {
        rc = walk_sequences(reads[f], read_lens[f], k, count_cb, &S);
        if (rc) { kset_free(&S); return rc; }
    }

    /* 3. adapter view: record i = rstrip(line[2i+1]) */
    {
        /* "last wins": walk records backwards, mark the first time each key is seen */
        kset seen;
        if (kset_init(&seen, n_records + 16)) { kset_free(&S); return -1; }
        /* collect line starts of the 2-line records */
        const char **ls = (const char **)malloc(sizeof(char *) * (n_records ? n_records : 1));
        size_t *ln = (size_t *)malloc(sizeof(size_t) * (n_records ? n_records : 1));
        if (!ls || !ln) { kset_free(&S); kset_free(&seen); free(ls); free(ln); return -1; }
        cur c = { fa, fa + fa_len };
        const char *s; size_t n;
        for (uint64_t i = 0; i < n_records; i++) {
            cur_line(&c, &s, &n);                 /* header */
            uint64_t id = 0; int ok = (n > 1);
            for (size_t j = 1; j < rstrip_len(s, n); j++) {
                if (s[j] < '0' || s[j] > '9') { ok = 0; break; }
                id = id * 10 + (uint64_t)(s[j] - '0');
            }
            if (header_id) header_id[i] = ok ? id : 0;
            cur_line(&c, &s, &n);
            ls[i] = s; ln[i] = rstrip_len(s, n);
        }
        for (uint64_t ii = n_records; ii-- > 0;) {
            uint64_t key; int up = 0;
            cnt[ii] = 0; in_set[ii] = 0; is_last[ii] = 0; raw_upper[ii] = 0;
            if (!pack_record(ls[ii], ln[ii], k, &key, &up)) continue;
            int64_t h = kset_find(&S, key);
            if (h < 0) continue;                  /* cannot happen: it was seeded */
            cnt[ii] = S.cnt[h]; in_set[ii] = 1; raw_upper[ii] = (uint8_t)up;
            if (kset_find(&seen, key) < 0) { is_last[ii] = 1; kset_seed(&seen, key); }
        }
        free(ls); free(ln); kset_free(&seen);
    }
    if (n_distinct) *n_distinct = S.n;
    kset_free(&S);
    return 0;
}

/* number of forward-strand k-windows (all ACGT) in the read files: the bench's unit of work */
static int tally_cb(void *ctx, uint64_t mer) { (void)mer; (*(uint64_t *)ctx)++; return 0; }
int orc_count_windows(const char *const *reads, const size_t *read_lens, int n_files, int k,
                      uint64_t *n_windows) {
    init_code();
    if (k < 1 || k > 32) return -3;
    uint64_t t = 0;
    for (int f = 0; f < n_files; f++) {
        int rc = walk_sequences(reads[f], read_lens[f], k, tally_cb, &t);
        if (rc) return rc;
    }
    *n_windows = t;
    return 0;
}
