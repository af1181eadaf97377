// ss_api.cu -- host side of the C ABI (include/strainscan_b200.h): context, k-mer FASTA loader and
// table build, FASTQ ingest (plain / gz, sharding, chunked streaming), count drivers, reducers.
//
// Replaces the process/file protocol around library/jellyfish-linux at
//   library/identify.py:73-103, library/identify_low_mem.py:67-90, library/identify_low_depth.py:46-74,
//   library/Vote_Strain_L2_Lasso_new_sp.py:354-403.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ss_common.cuh"
#include "ss_ingest.cuh"
#include "ss_kernels.cuh"
#include "ss_synth.cuh"
#include "ss_inflate.cuh"
#include "ss_fastx.h"
#include "ss_dgz_host.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void ss_set_error(const std::string &msg) { g_err = msg; }
int ss_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    return SS_ERR_CUDA;
}
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
extern "C" const char *ss_last_error(void) { return g_err.c_str(); }

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
#define SS_CHUNK_DEFAULT (32ull << 20)   // streaming chunk of FASTQ text (host -> device); env SS_CHUNK_BYTES
#define SS_SEG_DEFAULT (1ull << 30)      // read-cache segment when the total size is unknown (gzip); env SS_SEG_BYTES
#define SS_NPEND 4                       // host chunks whose H2D copy may be in flight
#define SS_NSLOT 4                       // device text slots of the streaming drivers (copy / inflate / scan overlap)
#define SS_NGZ 8                         // BGZF batches in flight on the device

struct ss_ctx {
    // Per-pass state (statistics, scratch vector, slot counters of the sets, streaming slots) is per context: the
    // count / load / reduce entry points take this lock, so two threads sharing a context serialise instead of
    // mixing their counters (one context per thread or per stream is the way to run passes concurrently).
    std::recursive_mutex mu;
    int device = 0;
    int n_sm = 0;
    cudaDeviceProp prop;
    cudaStream_t stream = nullptr, copy_stream = nullptr, own_stream = nullptr;
    unsigned long long *d_stats = nullptr;   // [0..3] kmers, hits, second, reads; [4] first bad byte; [5] table probes; [6] sink
    unsigned long long *h_stats = nullptr;   // pinned mirror
    uint32_t *d_dense = nullptr;             // scratch dense vector for host-output counts
    uint64_t dense_cap = 0;
    // streaming (ss_count_host / ss_count_files)
    uint8_t *d_chunk[SS_NSLOT] = {};
    uint32_t *d_chunk_line[SS_NSLOT] = {};
    cudaEvent_t ev_copied[SS_NSLOT] = {}, ev_done[SS_NSLOT] = {};
    uint8_t *h_pinned[SS_NSLOT] = {};            // staging for pageable sources
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_i = nullptr;
    size_t chunk_bytes = SS_CHUNK_DEFAULT, seg_bytes = SS_SEG_DEFAULT;
    // device inflate of BGZF batches (ss_gunzip.cu): compressed staging + member tables, double buffered
    bool device_bgzf = true;
    size_t bgzf_out_cap = 0;                     // text bytes one batch may inflate to
    uint8_t *d_comp[SS_NGZ] = {};
    ss_member *d_members[SS_NGZ] = {};
    cudaStream_t gz_stream[SS_NGZ] = {};         // batches inflate concurrently: one batch alone cannot fill the GPU
    cudaEvent_t ev_gz_copied[SS_NGZ] = {}, ev_gz_done[SS_NGZ] = {};
    unsigned int *d_gz = nullptr;                // [slot] work counters, [SS_NGZ] smallest failing member (0xFFFFFFFF = none)
    int gz_slot = 0;
    bool gz_used[SS_NGZ] = {};
    // binned probing (ss_count on resident reads against a large table with many table probes)
    unsigned long long *d_bin_keys = nullptr;    // pool of bin_pool_chunks chunks of SS_BIN_CHUNK keys
    uint32_t *d_bin_fill = nullptr;              // keys per chunk
    uint32_t *d_bin_n = nullptr;                 // [SS_BIN_PMAX] chunks per bin, [SS_BIN_PMAX] overflow flag
    unsigned long long *h_bin = nullptr;         // pinned: [0..5] the round's scan statistics, then the uint32 d_bin_n mirror
    uint64_t bin_pool_chunks = 0, bin_pool_want = 0;
    ss_text_source *src = nullptr;               // file ingest: producer threads + pinned chunk pool (lazy)
    // device inflate of ordinary gzip (ss_dgz.cu): decoder buffers and the text batch buffer, kept from file to file and
    // call to call (allocating ~11 GB per file cost more than a tenth of an 8-GPU config-3 pass); freed by ss_shutdown
    ss_dgz *dgz = nullptr;
    uint8_t *d_dgz_scratch = nullptr;
    size_t dgz_scratch_cap = 0;
    uint32_t *d_dgz_line = nullptr;          // line index of the batch being scanned where it was inflated (grows only)
    size_t dgz_line_words = 0;
    uint8_t *d_dgz_up[2] = {nullptr, nullptr};   // compressed bytes of the file being inflated / of the next one (grow only)
    size_t dgz_up_cap[2] = {0, 0};
    uint8_t *d_dgz_carry = nullptr;          // the partial record behind a batch's last whole one (grows only)
    size_t dgz_carry_cap = 0;
    cudaEvent_t ev_pend[SS_NPEND] = {nullptr, nullptr, nullptr, nullptr};
};

struct ss_kmerset {
    ss_ctx *ctx = nullptr;
    int k = 0;
    uint64_t n_records = 0, n_distinct = 0, n_buckets = 0;
    ss_bucket *d_buckets = nullptr;
    uint32_t *d_slot_cnt = nullptr;   // 4*n_buckets + 1
    uint32_t *d_slot_of = nullptr;    // n_records
    uint32_t *d_ord_of_slot = nullptr;   // 4*n_buckets + 1: the last record ordinal that holds the slot's k-mer
    uint32_t *d_dup = nullptr;           // records that are not their slot's representative (duplicate k-mers)
    uint64_t n_dup = 0;
    mutable bool counters_clean = false; // the last pass left every slot counter at zero (K3b clears what it reads)
    uint8_t *d_flags = nullptr;       // n_records
    uint32_t *d_row_of = nullptr;     // n_records or null (kid order == ordinal order)
    ss_fword *d_filter = nullptr;     // L2-resident prefilter or null
    uint32_t n_filter_words = 0;
    std::vector<uint8_t> flags;
    std::vector<uint64_t> header_ids;
    int has_ones = 0;
    // what the last resident pass of this set learned from its sample (direct or binned probing, ss_count_device):
    // a repeated pass over the same read cache (-b 1's second call, a re-run) does not sample again
    mutable uint64_t seen_reads_id = 0;
    mutable double seen_rate = 0.0, seen_per_tile = 0.0, seen_kmers_per_tile = 0.0;
    ss_table_view view() const {
        ss_table_view v;
        v.buckets = d_buckets; v.slot_cnt = d_slot_cnt; v.n_buckets = n_buckets;
        v.kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
        v.vmask = (k == 32) ? ~0u : ((1u << k) - 1u);
        v.k = k; v.has_ones = has_ones;
        v.filter = d_filter; v.n_filter_words = n_filter_words;
        return v;
    }
};

// The read cache: FASTQ text resident in HBM as one or more segments, each whole records ending in '\n'
// (files of unknown inflated size are appended segment by segment; one probe launch per segment).
struct ss_segment {
    uint8_t *d_text = nullptr;
    uint64_t len = 0, cap = 0;
    uint32_t n_tiles = 0;
    uint32_t *d_tile_line = nullptr;
    bool owns_text = false;
    bool indexed = false;             // d_tile_line holds the line index (built by the first pass over the segment)
};
static std::atomic<uint64_t> g_reads_id{0};
struct ss_reads {
    ss_ctx *ctx = nullptr;
    std::vector<ss_segment> seg;
    uint64_t len = 0;
    uint64_t id = ++g_reads_id;     // identity of this cache (never reused)
};

static inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int ss_init(int device, ss_ctx **out) {
    if (!out) return fail(SS_ERR_ARG, "ss_init: ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SS_ERR_NO_DEVICE, std::string("ss_init: no CUDA device (") + cudaGetErrorString(e) +
                                          "); this engine has no CPU fallback");
    if (device < 0 || device >= n) return fail(SS_ERR_ARG, "ss_init: device index out of range");
    ss_ctx *c = new ss_ctx();
    c->device = device;
    SS_CUDA(cudaSetDevice(device));
    SS_CUDA(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major != 10) {
        std::string m = std::string("ss_init: device '") + c->prop.name + "' is sm_" + std::to_string(c->prop.major) +
                        std::to_string(c->prop.minor) + "; kernels are built for sm_100a only (no fallback)";
        delete c;
        return fail(SS_ERR_NO_DEVICE, m);
    }
    c->n_sm = c->prop.multiProcessorCount;
    SS_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    SS_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SS_CUDA(cudaMalloc(&c->d_stats, 16 * sizeof(unsigned long long)));   // [8..13]: one binned round's scan statistics
    SS_CUDA(cudaMallocHost(&c->h_stats, 8 * sizeof(unsigned long long)));
    SS_CUDA(cudaEventCreate(&c->ev_a)); SS_CUDA(cudaEventCreate(&c->ev_b));
    SS_CUDA(cudaEventCreate(&c->ev_c)); SS_CUDA(cudaEventCreate(&c->ev_d)); SS_CUDA(cudaEventCreate(&c->ev_i));
    for (int i = 0; i < SS_NSLOT; i++) {
        SS_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        SS_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < SS_NPEND; i++) SS_CUDA(cudaEventCreateWithFlags(&c->ev_pend[i], cudaEventDisableTiming));
    if (const char *e = getenv("SS_CHUNK_BYTES")) { long long v = atoll(e); if (v >= (256 << 10) && v <= (1ll << 30)) c->chunk_bytes = (size_t)v; }
    if (const char *e = getenv("SS_SEG_BYTES")) { long long v = atoll(e); if (v >= (256 << 10)) c->seg_bytes = (size_t)v; }
    c->bgzf_out_cap = 6 * c->chunk_bytes;
    if (const char *e = getenv("SS_BGZF_OUT_CAP")) { long long v = atoll(e); if (v >= (1 << 20) && v <= (4ll << 30)) c->bgzf_out_cap = (size_t)v; }
    if (const char *e = getenv("SS_BGZF_GPU")) c->device_bgzf = atoi(e) != 0;
    for (int i = 0; i < SS_NGZ; i++) {
        SS_CUDA(cudaEventCreateWithFlags(&c->ev_gz_done[i], cudaEventDisableTiming));
        SS_CUDA(cudaEventCreateWithFlags(&c->ev_gz_copied[i], cudaEventDisableTiming));
        SS_CUDA(cudaStreamCreateWithFlags(&c->gz_stream[i], cudaStreamNonBlocking));
    }
    SS_CUDA(cudaMalloc(&c->d_gz, (SS_NGZ + 1) * sizeof(unsigned int)));
    *out = c;
    return SS_OK;
}

extern "C" int ss_shutdown(ss_ctx *c) {
    if (!c) return SS_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < SS_NSLOT; i++) {
        cudaFree(c->d_chunk[i]); cudaFree(c->d_chunk_line[i]);
        if (c->h_pinned[i]) cudaFreeHost(c->h_pinned[i]);
        cudaEventDestroy(c->ev_copied[i]); cudaEventDestroy(c->ev_done[i]);
    }
    delete c->src;
    delete c->dgz;
    cudaFree(c->d_dgz_scratch);
    cudaFree(c->d_dgz_line);
    cudaFree(c->d_dgz_up[0]); cudaFree(c->d_dgz_up[1]); cudaFree(c->d_dgz_carry);
    for (int i = 0; i < SS_NGZ; i++) {
        cudaFree(c->d_comp[i]); cudaFree(c->d_members[i]);
        cudaEventDestroy(c->ev_gz_done[i]); cudaEventDestroy(c->ev_gz_copied[i]);
        cudaStreamDestroy(c->gz_stream[i]);
    }
    cudaFree(c->d_gz);
    for (int i = 0; i < SS_NPEND; i++) cudaEventDestroy(c->ev_pend[i]);
    cudaFree(c->d_dense); cudaFree(c->d_stats); cudaFreeHost(c->h_stats);
    cudaFree(c->d_bin_keys); cudaFree(c->d_bin_fill); cudaFree(c->d_bin_n); cudaFreeHost(c->h_bin);
    cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b); cudaEventDestroy(c->ev_c); cudaEventDestroy(c->ev_d); cudaEventDestroy(c->ev_i);
    cudaStreamDestroy(c->own_stream); cudaStreamDestroy(c->copy_stream);
    delete c;
    return SS_OK;
}

extern "C" int ss_set_stream(ss_ctx *c, void *stream) {
    if (!c) return fail(SS_ERR_ARG, "ss_set_stream: ctx is NULL");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return SS_OK;
}

extern "C" int ss_device_info(const ss_ctx *c, char *name, size_t cap, int *n_sm, uint64_t *mem) {
    if (!c) return fail(SS_ERR_ARG, "ss_device_info: ctx is NULL");
    if (name && cap) { strncpy(name, c->prop.name, cap - 1); name[cap - 1] = 0; }
    if (n_sm) *n_sm = c->n_sm;
    if (mem) *mem = c->prop.totalGlobalMem;
    return SS_OK;
}

static int ensure_dense(ss_ctx *c, uint64_t n) {
    if (n <= c->dense_cap) return SS_OK;
    cudaFree(c->d_dense); c->d_dense = nullptr; c->dense_cap = 0;
    SS_CUDA(cudaMalloc(&c->d_dense, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
    c->dense_cap = n;
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// file helpers
// ---------------------------------------------------------------------------------------------
// whole file into memory (k-mer FASTA databases; gzip'ed ones are inflated by ss_inflate.cuh)
static int read_file(const char *path, std::vector<char> &out) {
    std::string err;
    int rc = ss_read_whole_file(path, out, err);
    return rc ? fail(rc, err) : SS_OK;
}

static size_t trim_tail(const char *buf, size_t len) { return ss_trim_tail(buf, len); }
static size_t find_record_start(const char *buf, size_t len, size_t from) { return ss_find_record_start(buf, len, from); }

static int check_fastq_head(const char *buf, size_t len, const char *what) {
    if (len == 0) return SS_OK;
    if (buf[0] != '@') return fail(SS_ERR_FORMAT, std::string(what) + ": not a FASTQ file (first byte is not '@')");
    return SS_OK;
}

// in-memory read text of any dialect Jellyfish accepts -> 4-line FASTQ (`store` keeps a rewritten copy alive)
static int normalize_host_text(const char *&buf, size_t &len, std::vector<char> &store, const char *what) {
    int kind = ss_fastx_kind(buf, len, true);
    if (kind < 0) return fail(SS_ERR_FORMAT, std::string(what) + ": not a FASTQ / FASTA file (first byte is neither '@' nor '>')");
    if (kind == 0) return SS_OK;
    store.clear();
    if (ss_fastx_normalize(buf, len, store))
        return fail(SS_ERR_FORMAT, std::string(what) + ": malformed FASTA / multi-line FASTQ (record framing breaks)");
    buf = store.data(); len = store.size();
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// k-mer FASTA -> packed keys (host, threaded) -> device table
// ---------------------------------------------------------------------------------------------
static inline int base_code(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

static void parse_fasta_range(const char *text, size_t len, size_t lo, size_t hi, uint64_t line_at_lo, int k,
                              uint64_t n_records, uint64_t *keys, uint8_t *ok, uint8_t *raw_upper,
                              uint64_t *header_ids) {
    size_t p = lo;
    uint64_t line = line_at_lo;
    if (p > 0 && text[p - 1] != '\n') {   // partial first line belongs to the previous range
        const char *nl = (const char *)memchr(text + p, '\n', len - p);
        if (!nl) return;
        p = (size_t)(nl - text) + 1;
        line++;
    }
    while (p < hi && p < len) {
        const char *nl = (const char *)memchr(text + p, '\n', len - p);
        size_t e = nl ? (size_t)(nl - text) : len;
        uint64_t rec = line >> 1;
        if (rec < n_records) {
            size_t n = e - p;   // rstrip() as Python does on the record line
            while (n > 0 && (text[p + n - 1] == ' ' || text[p + n - 1] == '\t' || text[p + n - 1] == '\r' ||
                             text[p + n - 1] == '\v' || text[p + n - 1] == '\f'))
                n--;
            if ((line & 1) == 0) {
                uint64_t id = 0;
                bool good = n > 1 && text[p] == '>';
                for (size_t j = 1; j < n && good; j++) {
                    if (text[p + j] < '0' || text[p + j] > '9') good = false;
                    else id = id * 10 + (uint64_t)(text[p + j] - '0');
                }
                header_ids[rec] = good ? id : 0;
            } else {
                uint64_t key = 0;
                bool good = ((int)n == k), up = true;
                for (int j = 0; j < k && good; j++) {
                    unsigned char ch = (unsigned char)text[p + j];
                    int c = base_code(ch);
                    if (c < 0) { good = false; break; }
                    if (ch >= 'a') up = false;
                    key |= (uint64_t)c << (2 * j);   // first base in the low bits, as the kernel extracts it
                }
                keys[rec] = good ? key : 0;
                ok[rec] = good ? 1 : 0;
                raw_upper[rec] = (good && up) ? 1 : 0;
            }
        }
        line++;
        p = e + 1;
    }
}

// the parsed form of a k-mer FASTA: what the table build consumes and what the binary cache stores
struct parsed_set {
    uint64_t n = 0;
    std::vector<uint64_t> keys;          // packed k-mer of every record (0 when not ok)
    std::vector<uint8_t> ok, raw_upper;  // record is k ACGT chars after case folding / already uppercase
    std::vector<uint64_t> header_ids;    // integer after '>' (0 if none)
};

static int build_from_parsed(ss_ctx *c, int k, parsed_set &ps, ss_kmerset **out);

static int parse_set(const char *text, size_t len, int k, parsed_set &ps) {
    // line structure (Python readlines(): a last line without '\n' still counts)
    unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (len < (1u << 20)) T = 1;
    std::vector<size_t> cut(T + 1);
    for (unsigned t = 0; t <= T; t++) cut[t] = len / T * t;
    cut[T] = len;
    std::vector<uint64_t> nl(T, 0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                uint64_t cnt = 0;
                const char *p = text + cut[t], *e = text + cut[t + 1];
                while (p < e) {
                    const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
                    if (!q) break;
                    cnt++; p = q + 1;
                }
                nl[t] = cnt;
            });
        for (auto &x : th) x.join();
    }
    uint64_t n_lines = 0;
    std::vector<uint64_t> line_at(T);
    for (unsigned t = 0; t < T; t++) { line_at[t] = n_lines; n_lines += nl[t]; }
    if (len > 0 && text[len - 1] != '\n') n_lines++;
    uint64_t n = n_lines / 2;
    if (n >= 0xFFFFFFF0ull) return fail(SS_ERR_UNSUPPORTED, "kmerset: more than 2^32 records");
    ps.n = n;
    ps.keys.assign(n, 0); ps.ok.assign(n, 0); ps.raw_upper.assign(n, 0); ps.header_ids.assign(n, 0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                parse_fasta_range(text, len, cut[t], cut[t + 1], line_at[t], k, n, ps.keys.data(), ps.ok.data(),
                                  ps.raw_upper.data(), ps.header_ids.data());
            });
        for (auto &x : th) x.join();
    }
    return SS_OK;
}

static int build_set(ss_ctx *c, const char *text, size_t len, int k, ss_kmerset **out, parsed_set *keep = nullptr) {
    if (!c || !out) return fail(SS_ERR_ARG, "kmerset: NULL argument");
    *out = nullptr;
    if (k < 1 || k > 32) return fail(SS_ERR_UNSUPPORTED, "kmerset: k must be in 1..32 (one 64-bit word per k-mer)");
    parsed_set local;
    parsed_set &ps = keep ? *keep : local;
    int rc = parse_set(text, len, k, ps);
    if (rc) return rc;
    return build_from_parsed(c, k, ps, out);
}

static int build_from_parsed(ss_ctx *c, int k, parsed_set &ps, ss_kmerset **out) {
    *out = nullptr;
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    SS_CUDA(cudaSetDevice(c->device));
    const uint64_t n = ps.n;
    std::vector<uint64_t> &keys = ps.keys;
    std::vector<uint8_t> &ok = ps.ok, &raw_upper = ps.raw_upper;
    ss_kmerset *s = new ss_kmerset();
    s->ctx = c; s->k = k; s->n_records = n;
    s->header_ids = ps.header_ids;
    uint64_t n_ok = 0;
    for (uint64_t i = 0; i < n; i++) {
        n_ok += ok[i];
        if (k == 32 && ok[i] && keys[i] == SS_EMPTY) s->has_ones = 1;
    }

    // table geometry: mean SS_BUCKET_LOAD keys per 4-slot bucket (default 1.0 -> P(bucket full) ~ 2 %)
    double load = 1.0;
    if (const char *e = getenv("SS_BUCKET_LOAD")) { double v = atof(e); if (v > 0.05 && v <= 3.0) load = v; }
    s->n_buckets = std::max<uint64_t>(64, (uint64_t)((double)n_ok / load) + 1);
    if (4 * s->n_buckets + 1 >= 0xFFFFFFFFull) { delete s; return fail(SS_ERR_UNSUPPORTED, "kmerset: table too large"); }

    uint64_t *d_keys = nullptr; uint8_t *d_ok = nullptr; uint32_t *d_last = nullptr;
    unsigned long long *d_nd = nullptr;
    const uint64_t n_slots = 4 * s->n_buckets + 1;
    auto cleanup = [&]() { cudaFree(d_keys); cudaFree(d_ok); cudaFree(d_nd); };
#define SS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ss_kmerset_free(s); return ss_cuda_fail(e_, #x, __FILE__, __LINE__); } } while (0)
    SS_TRY(cudaMalloc(&s->d_buckets, s->n_buckets * sizeof(ss_bucket)));
    SS_TRY(cudaMalloc(&s->d_slot_cnt, n_slots * sizeof(uint32_t)));
    SS_TRY(cudaMalloc(&s->d_slot_of, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
    SS_TRY(cudaMalloc(&s->d_flags, std::max<uint64_t>(n, 1)));
    SS_TRY(cudaMalloc(&d_keys, std::max<uint64_t>(n, 1) * sizeof(uint64_t)));
    SS_TRY(cudaMalloc(&d_ok, std::max<uint64_t>(n, 1)));
    SS_TRY(cudaMalloc(&s->d_ord_of_slot, n_slots * sizeof(uint32_t)));
    d_last = s->d_ord_of_slot;
    SS_TRY(cudaMalloc(&d_nd, sizeof(unsigned long long)));
    SS_TRY(cudaMemsetAsync(s->d_buckets, 0xFF, s->n_buckets * sizeof(ss_bucket), c->stream));
    SS_TRY(cudaMemsetAsync(s->d_slot_cnt, 0, n_slots * sizeof(uint32_t), c->stream));
    SS_TRY(cudaMemsetAsync(d_last, 0, n_slots * sizeof(uint32_t), c->stream));
    SS_TRY(cudaMemsetAsync(d_nd, 0, sizeof(unsigned long long), c->stream));
    if (n) {
        SS_TRY(cudaMemcpyAsync(d_keys, keys.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        SS_TRY(cudaMemcpyAsync(d_ok, ok.data(), n, cudaMemcpyHostToDevice, c->stream));
        // flags start as RAW_UPPER bit only; the flags kernel adds IN_SET / IS_LAST
        std::vector<uint8_t> f0(n);
        for (uint64_t i = 0; i < n; i++) f0[i] = raw_upper[i] ? SS_REC_RAW_UPPER : 0;
        SS_TRY(cudaMemcpyAsync(s->d_flags, f0.data(), n, cudaMemcpyHostToDevice, c->stream));
        SS_TRY(cudaStreamSynchronize(c->stream));   // f0 goes out of scope
    }
    SS_TRY(ss_launch_insert(d_keys, d_ok, n, (unsigned long long *)s->d_buckets, s->n_buckets, s->d_slot_of, d_last,
                            d_nd, c->stream));
    {   // L2-resident prefilter.  First built only for tables that cannot live in L2 themselves; but a k-mer that is
        // not in the set is also cheaper to turn away with one 4-byte filter load than with a 32-byte bucket load and
        // four 64-bit compares when the table does sit in L2 (11 MB cluster set, 1.6 % hits: 7.6 ms without, DESIGN.md
        // section 4), so every set of more than a few thousand k-mers gets one
        double bits_per_key = 16.0, max_mb = 48.0, min_table_mb = 0.25;
        if (const char *e = getenv("SS_FILTER_MIN_TABLE_MB")) { double v = atof(e); if (v >= 0) min_table_mb = v; }
        int force = -1;
        if (const char *e = getenv("SS_FILTER")) force = atoi(e);
        if (const char *e = getenv("SS_FILTER_BITS")) { double v = atof(e); if (v >= 4 && v <= 64) bits_per_key = v; }
        if (const char *e = getenv("SS_FILTER_MAX_MB")) { double v = atof(e); if (v >= 1 && v <= 1024) max_mb = v; }
        bool want = force >= 0 ? force != 0 : (double)s->n_buckets * sizeof(ss_bucket) > min_table_mb * 1e6;
        if (want && n_ok > 0) {
            // SS_FILTER_PAIRS: two (k-1)-mers per key go in
            const double wbits = 8.0 * sizeof(ss_fword);
            double words = std::min((double)n_ok * (SS_FILTER_PAIRS ? 2.0 : 1.0) * bits_per_key / wbits, max_mb * 1e6 * 8.0 / wbits);
            s->n_filter_words = (uint32_t)std::max(1024.0, words);
            SS_TRY(cudaMalloc(&s->d_filter, (uint64_t)s->n_filter_words * sizeof(ss_fword)));
            SS_TRY(cudaMemsetAsync(s->d_filter, 0, (uint64_t)s->n_filter_words * sizeof(ss_fword), c->stream));
            SS_TRY(ss_launch_filter_build(d_keys, d_ok, n, s->d_filter, s->n_filter_words, s->view().kmask, c->stream));
        }
    }
    SS_TRY(ss_launch_flags(s->d_slot_of, d_last, n, s->d_flags, c->stream));
    s->flags.assign(n, 0);
    unsigned long long nd = 0;
    if (n) SS_TRY(cudaMemcpyAsync(s->flags.data(), s->d_flags, n, cudaMemcpyDeviceToHost, c->stream));
    SS_TRY(cudaMemcpyAsync(&nd, d_nd, sizeof nd, cudaMemcpyDeviceToHost, c->stream));
    SS_TRY(cudaStreamSynchronize(c->stream));
    s->n_distinct = nd + (s->has_ones ? 1 : 0);
    {   // duplicate records: in the set, but not the last record of their k-mer
        std::vector<uint32_t> dup;
        for (uint64_t i = 0; i < n; i++)
            if ((s->flags[i] & SS_REC_IN_SET) && !(s->flags[i] & SS_REC_IS_LAST)) dup.push_back((uint32_t)i);
        s->n_dup = dup.size();
        if (s->n_dup) {
            SS_TRY(cudaMalloc(&s->d_dup, s->n_dup * sizeof(uint32_t)));
            SS_TRY(cudaMemcpy(s->d_dup, dup.data(), s->n_dup * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    s->counters_clean = true;       // zeroed above, nothing counted yet

    // L2 row order: sorted by kid (Vote_...:386); identity when headers are 1..n in order
    bool perm = n > 0;
    bool ident = true;
    for (uint64_t i = 0; i < n && perm; i++) {
        uint64_t h = s->header_ids[i];
        if (h < 1 || h > n) perm = false;
        if (h != i + 1) ident = false;
    }
    if (perm && !ident) {
        std::vector<uint8_t> seen(n, 0);
        std::vector<uint32_t> row(n);
        for (uint64_t i = 0; i < n && perm; i++) {
            uint64_t r = s->header_ids[i] - 1;
            if (seen[r]) perm = false;
            seen[r] = 1; row[i] = (uint32_t)r;
        }
        if (perm) {
            SS_TRY(cudaMalloc(&s->d_row_of, n * sizeof(uint32_t)));
            SS_TRY(cudaMemcpy(s->d_row_of, row.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
#undef SS_TRY
    cleanup();
    *out = s;
    return SS_OK;
}

extern "C" int ss_kmerset_from_text(ss_ctx *c, const char *text, size_t len, int k, ss_kmerset **out) {
    if (!text && len) return fail(SS_ERR_ARG, "ss_kmerset_from_text: text is NULL");
    return build_set(c, text, len, k, out);
}

extern "C" int ss_kmerset_from_fasta(ss_ctx *c, const char *path, int k, ss_kmerset **out) {
    if (!path) return fail(SS_ERR_ARG, "ss_kmerset_from_fasta: path is NULL");
    std::vector<char> buf;
    int rc = read_file(path, buf);
    if (rc) return rc;
    return build_set(c, buf.data(), buf.size(), k, out);
}

// ---- binary database cache: the parsed form of a k-mer FASTA next to / instead of re-parsing the
// text.  The FASTA stays the source of truth: the cache records its size and mtime and is ignored when they differ.
struct cache_header {
    char magic[8];              // "SSB200K1"
    uint32_t k, ids_mode;       // ids_mode: 0 all zero, 1 identity (1..n), 2 constant, 3 explicit array
    uint64_t n, ids_const, src_size, src_mtime_ns;
};

static bool stat_file(const char *path, uint64_t *size, uint64_t *mtime_ns) {
    struct stat st;
    if (stat(path, &st) != 0) return false;
    *size = (uint64_t)st.st_size;
    *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    return true;
}

static bool cache_write(const char *cache_path, int k, const parsed_set &ps, uint64_t src_size, uint64_t src_mtime_ns) {
    std::string tmp = std::string(cache_path) + ".tmp" + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    cache_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "SSB200K1", 8);
    h.k = (uint32_t)k; h.n = ps.n; h.src_size = src_size; h.src_mtime_ns = src_mtime_ns;
    bool zero = true, ident = true, cst = true;
    for (uint64_t i = 0; i < ps.n; i++) {
        uint64_t v = ps.header_ids[i];
        zero &= v == 0; ident &= v == i + 1; cst &= v == ps.header_ids[0];
    }
    h.ids_mode = zero ? 0 : ident ? 1 : cst ? 2 : 3;
    h.ids_const = (cst && ps.n) ? ps.header_ids[0] : 0;
    std::vector<uint8_t> fl(ps.n);
    for (uint64_t i = 0; i < ps.n; i++) fl[i] = (uint8_t)((ps.ok[i] ? 1 : 0) | (ps.raw_upper[i] ? 2 : 0));
    bool good = fwrite(&h, sizeof h, 1, f) == 1 && (ps.n == 0 || (fwrite(ps.keys.data(), 8, ps.n, f) == ps.n && fwrite(fl.data(), 1, ps.n, f) == ps.n));
    if (good && h.ids_mode == 3) good = fwrite(ps.header_ids.data(), 8, ps.n, f) == ps.n;
    good = (fclose(f) == 0) && good;
    if (!good || rename(tmp.c_str(), cache_path) != 0) { remove(tmp.c_str()); return false; }
    return true;
}

static bool cache_read(const char *cache_path, int k, uint64_t src_size, uint64_t src_mtime_ns, bool check_src, parsed_set &ps) {
    FILE *f = fopen(cache_path, "rb");
    if (!f) return false;
    cache_header h;
    bool good = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "SSB200K1", 8) == 0 && h.k == (uint32_t)k && h.ids_mode <= 3 &&
                h.n < 0xFFFFFFF0ull && (!check_src || (h.src_size == src_size && h.src_mtime_ns == src_mtime_ns));
    if (good) {
        ps.n = h.n;
        ps.keys.resize(h.n); ps.ok.resize(h.n); ps.raw_upper.resize(h.n); ps.header_ids.resize(h.n);
        std::vector<uint8_t> fl(h.n);
        good = h.n == 0 || (fread(ps.keys.data(), 8, h.n, f) == h.n && fread(fl.data(), 1, h.n, f) == h.n);
        for (uint64_t i = 0; good && i < h.n; i++) { ps.ok[i] = fl[i] & 1; ps.raw_upper[i] = (fl[i] >> 1) & 1; }
        if (good && h.ids_mode == 3) good = fread(ps.header_ids.data(), 8, h.n, f) == h.n;
        else if (good) for (uint64_t i = 0; i < h.n; i++) ps.header_ids[i] = h.ids_mode == 1 ? i + 1 : h.ids_mode == 2 ? h.ids_const : 0;
        if (good) { char extra; good = fread(&extra, 1, 1, f) == 0; }          // nothing behind the arrays
        if (good && k < 32) for (uint64_t i = 0; good && i < h.n; i += 4097) good = (ps.keys[i] >> (2 * k)) == 0;   // spot check
    }
    fclose(f);
    return good;
}

extern "C" int ss_kmerset_from_fasta_cached(ss_ctx *c, const char *path, int k, const char *cache_path, int *cache_hit,
                                            ss_kmerset **out) {
    if (!c || !path || !cache_path || !out) return fail(SS_ERR_ARG, "ss_kmerset_from_fasta_cached: NULL argument");
    *out = nullptr;
    if (cache_hit) *cache_hit = 0;
    if (k < 1 || k > 32) return fail(SS_ERR_UNSUPPORTED, "kmerset: k must be in 1..32 (one 64-bit word per k-mer)");
    uint64_t sz = 0, mt = 0;
    if (!stat_file(path, &sz, &mt)) return fail(SS_ERR_IO, std::string("cannot open ") + path);
    parsed_set ps;
    if (cache_read(cache_path, k, sz, mt, true, ps)) {
        if (cache_hit) *cache_hit = 1;
        return build_from_parsed(c, k, ps, out);
    }
    std::vector<char> buf;
    int rc = read_file(path, buf);
    if (rc) return rc;
    rc = build_set(c, buf.data(), buf.size(), k, out, &ps);
    if (rc) return rc;
    cache_write(cache_path, k, ps, sz, mt);          // best effort: a read-only cache directory is not an error
    return SS_OK;
}

extern "C" int ss_kmerset_free(ss_kmerset *s) {
    if (!s) return SS_OK;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_buckets); cudaFree(s->d_slot_cnt); cudaFree(s->d_slot_of); cudaFree(s->d_flags);
    cudaFree(s->d_ord_of_slot); cudaFree(s->d_dup);
    cudaFree(s->d_row_of); cudaFree(s->d_filter);
    delete s;
    return SS_OK;
}
extern "C" uint64_t ss_kmerset_records(const ss_kmerset *s) { return s ? s->n_records : 0; }
extern "C" uint64_t ss_kmerset_distinct(const ss_kmerset *s) { return s ? s->n_distinct : 0; }
extern "C" int ss_kmerset_k(const ss_kmerset *s) { return s ? s->k : 0; }
extern "C" uint64_t ss_kmerset_table_bytes(const ss_kmerset *s) {
    return s ? s->n_buckets * sizeof(ss_bucket) + (4 * s->n_buckets + 1) * 4 + s->n_records * 5 +
                   (uint64_t)s->n_filter_words * sizeof(ss_fword) : 0;
}
extern "C" int ss_kmerset_flags(const ss_kmerset *s, uint8_t *flags) {
    if (!s || !flags) return fail(SS_ERR_ARG, "ss_kmerset_flags: NULL argument");
    if (s->n_records) memcpy(flags, s->flags.data(), s->n_records);
    return SS_OK;
}
extern "C" int ss_kmerset_header_ids(const ss_kmerset *s, uint64_t *ids) {
    if (!s || !ids) return fail(SS_ERR_ARG, "ss_kmerset_header_ids: NULL argument");
    if (s->n_records) memcpy(ids, s->header_ids.data(), s->n_records * sizeof(uint64_t));
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// reads
// ---------------------------------------------------------------------------------------------
extern "C" size_t ss_reads_device_capacity(size_t len) { return round_up(len, SS_TILE) + SS_TEXT_PAD; }
extern "C" uint64_t ss_reads_bytes(const ss_reads *r) { return r ? r->len : 0; }

static void free_segment(ss_segment &g) {
    if (g.owns_text) cudaFree(g.d_text);
    cudaFree(g.d_tile_line);
    g = ss_segment();
}

extern "C" int ss_reads_free(ss_reads *r) {
    if (!r) return SS_OK;
    cudaSetDevice(r->ctx->device);
    for (auto &g : r->seg) free_segment(g);
    delete r;
    return SS_OK;
}

// pad one segment whose text (whole records) is already on the device.  Its line index (K1) is built by the first
// pass that scans it (ensure_index, inside that pass's timed region) and kept for the later passes.
static int finish_segment(ss_ctx *c, ss_segment &g) {
    g.n_tiles = (uint32_t)((g.len + SS_TILE - 1) / SS_TILE);
    size_t need = ss_reads_device_capacity(g.len);
    if (g.cap < need) return fail(SS_ERR_ARG, "reads: device buffer capacity too small (see ss_reads_device_capacity)");
    SS_CUDA(cudaMemsetAsync(g.d_text + g.len, '\n', need - g.len, c->stream));
    SS_CUDA(cudaMalloc(&g.d_tile_line, ss_index_words(g.n_tiles) * sizeof(uint32_t)));
    g.indexed = false;
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}

static int ensure_index(ss_ctx *c, const ss_reads *r, uint32_t *launches) {
    for (const ss_segment &cg : r->seg) {
        ss_segment &g = const_cast<ss_segment &>(cg);
        if (g.indexed || !g.n_tiles) continue;
        SS_CUDA(ss_launch_index(g.d_text, g.n_tiles, g.d_tile_line, 0, c->n_sm, c->stream));
        g.indexed = true;
        if (launches) (*launches)++;
    }
    return SS_OK;
}

extern "C" int ss_reads_drop_index(ss_reads *r) {
    if (!r) return fail(SS_ERR_ARG, "ss_reads_drop_index: reads is NULL");
    for (ss_segment &g : r->seg) g.indexed = false;
    return SS_OK;
}

static int alloc_segment(ss_segment &g, uint64_t text_bytes) {
    g = ss_segment();
    g.cap = ss_reads_device_capacity(text_bytes);
    g.owns_text = true;
    cudaError_t e = cudaMalloc(&g.d_text, g.cap);
    if (e != cudaSuccess) { g = ss_segment(); return ss_cuda_fail(e, "cudaMalloc(read cache segment)", __FILE__, __LINE__); }
    return SS_OK;
}

// concatenate normalised files: each ends with exactly one '\n', no trailing blank lines
static int gather_host_text(const char *const *bufs, const size_t *lens, int n, std::vector<char> &all) {
    size_t total = 0;
    for (int i = 0; i < n; i++) total += lens[i] + 1;
    all.reserve(total);
    std::vector<char> store;
    for (int i = 0; i < n; i++) {
        const char *b = bufs[i];
        size_t bl = lens[i];
        int rc = normalize_host_text(b, bl, store, "reads");
        if (rc) return rc;
        size_t l = trim_tail(b, bl);
        rc = check_fastq_head(b, l, "reads");
        if (rc) return rc;
        if (l == 0) continue;
        all.insert(all.end(), b, b + l);
        all.push_back('\n');
    }
    return SS_OK;
}

static int reads_from_vector(ss_ctx *c, const char *data, size_t len, ss_reads **out) {
    ss_reads *r = new ss_reads();
    r->ctx = c; r->len = len;
    r->seg.emplace_back();
    ss_segment &g = r->seg.back();
    int rc = alloc_segment(g, len);
    if (rc) { delete r; return rc; }
    g.len = len;
    if (len) {
        cudaError_t e = cudaMemcpyAsync(g.d_text, data, len, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { ss_reads_free(r); return ss_cuda_fail(e, "H2D reads", __FILE__, __LINE__); }
    }
    rc = finish_segment(c, g);
    if (rc) { ss_reads_free(r); return rc; }
    *out = r;
    return SS_OK;
}

extern "C" int ss_reads_from_host(ss_ctx *c, const char *const *bufs, const size_t *lens, int n, ss_reads **out) {
    if (!c || !out || (n > 0 && (!bufs || !lens))) return fail(SS_ERR_ARG, "ss_reads_from_host: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    *out = nullptr;
    SS_CUDA(cudaSetDevice(c->device));
    std::vector<char> all;
    int rc = gather_host_text(bufs, lens, n, all);
    if (rc) return rc;
    return reads_from_vector(c, all.data(), all.size(), out);
}

extern "C" int ss_fastq_shard_range(const char *buf, size_t len, int shard, int n_shards, size_t *lo, size_t *hi) {
    if (!lo || !hi || (!buf && len)) return fail(SS_ERR_ARG, "ss_fastq_shard_range: NULL argument");
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(SS_ERR_ARG, "ss_fastq_shard_range: bad shard / n_shards");
    *lo = find_record_start(buf, len, (size_t)((unsigned __int128)len * shard / n_shards));
    *hi = (shard + 1 == n_shards) ? len : find_record_start(buf, len, (size_t)((unsigned __int128)len * (shard + 1) / n_shards));
    return SS_OK;
}

// the context's file ingest (producer threads + pinned chunk pool), created at first use
static int ensure_source(ss_ctx *c) {
    if (c->src && c->src->ready()) return SS_OK;
    if (!c->src) c->src = new ss_text_source();
    unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    {   // the cores this process may actually run on (a rank bound to its GPU's cores, a container quota)
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0) { int n = CPU_COUNT(&set); if (n >= 1) hw = std::max(2u, std::min(hw, (unsigned)n)); }
    }
    int threads = (int)(hw > 8 ? std::min(12u, hw - 4) : hw);   // measured: 8 -> 21-30, 12 -> 35 GB/s of plain text; few cores: all of them
    if (const char *e = getenv("SS_INGEST_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) threads = v; }
    // one buffer per producer, the copies the consumer may have in flight (released as they complete), and slack
    int rc = c->src->init(c->chunk_bytes, threads + 3 + SS_NPEND, threads, true, c->device_bgzf, c->bgzf_out_cap);
    if (rc) return fail(rc, c->src->error());
    return SS_OK;
}

// host chunks whose H2D copy is still in flight: released back to the producers once their event fires
struct pending_ring {
    ss_ctx *c;
    ss_chunk *chunk[SS_NPEND] = {nullptr, nullptr, nullptr, nullptr};
    int at = 0;
    std::function<void(ss_chunk *)> on_done;          // called when a chunk's copy has completed, before it is handed back
    explicit pending_ring(ss_ctx *ctx) : c(ctx) {}
    void give_back(int i) { if (on_done) on_done(chunk[i]); c->src->release(chunk[i]); chunk[i] = nullptr; }
    // call right after the copy of `ch` was issued on `st`
    cudaError_t push(ss_chunk *ch, cudaStream_t st) {
        for (int i = 0; i < SS_NPEND; i++)              // hand back every buffer whose copy has completed by now
            if (chunk[i] && i != at && cudaEventQuery(c->ev_pend[i]) == cudaSuccess) give_back(i);
        if (chunk[at]) {
            cudaError_t e = cudaEventSynchronize(c->ev_pend[at]);
            if (e != cudaSuccess) return e;
            give_back(at);
        }
        cudaError_t e = cudaEventRecord(c->ev_pend[at], st);
        if (e != cudaSuccess) return e;
        chunk[at] = ch;
        at = (at + 1) % SS_NPEND;
        return cudaSuccess;
    }
    void drain() {
        for (int i = 0; i < SS_NPEND; i++)
            if (chunk[i]) { cudaEventSynchronize(c->ev_pend[i]); give_back(i); }
    }
};

// Stage one BGZF batch (compressed members + member table + the two host-decoded text pieces) and
// inflate it so that the batch's text starts at `dst` (device).  Copies go on the copy stream, the
// kernel on one of SS_NGZ inflate streams; *done is the event the consumer of the text waits for.
static int inflate_batch(ss_ctx *c, ss_chunk *ch, uint8_t *dst, cudaEvent_t *done) {
    const int b = c->gz_slot;
    if (!c->d_comp[b]) {
        SS_CUDA(cudaMalloc(&c->d_comp[b], c->chunk_bytes + 64));
        SS_CUDA(cudaMalloc(&c->d_members[b], (size_t)SS_BGZF_MAX_MEMBERS * sizeof(ss_member)));
    }
    if (c->gz_used[b]) SS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_gz_done[b], 0));   // kernel done with the slot
    SS_CUDA(cudaMemcpyAsync(c->d_comp[b], ch->text, ch->len, cudaMemcpyHostToDevice, c->copy_stream));
    SS_CUDA(cudaMemcpyAsync(c->d_members[b], ch->members, (size_t)ch->n_members * sizeof(ss_member), cudaMemcpyHostToDevice,
                            c->copy_stream));
    if (ch->pre_len) SS_CUDA(cudaMemcpyAsync(dst, ch->pre_text, ch->pre_len, cudaMemcpyHostToDevice, c->copy_stream));
    if (ch->post_len)
        SS_CUDA(cudaMemcpyAsync(dst + ch->pre_len + ch->inflated_len, ch->post_text, ch->post_len, cudaMemcpyHostToDevice,
                                c->copy_stream));
    SS_CUDA(cudaEventRecord(c->ev_gz_copied[b], c->copy_stream));
    SS_CUDA(cudaStreamWaitEvent(c->gz_stream[b], c->ev_gz_copied[b], 0));
    SS_CUDA(ss_launch_gunzip(c->d_comp[b], c->d_members[b], ch->n_members, dst, c->d_gz + b, c->d_gz + SS_NGZ, c->n_sm,
                             c->gz_stream[b]));
    SS_CUDA(cudaEventRecord(c->ev_gz_done[b], c->gz_stream[b]));
    if (done) *done = c->ev_gz_done[b];
    c->gz_used[b] = true;
    c->gz_slot = (b + 1) % SS_NGZ;
    return SS_OK;
}

static int gz_begin(ss_ctx *c) {
    for (int i = 0; i < SS_NGZ; i++) c->gz_used[i] = false;
    SS_CUDA(cudaMemsetAsync(c->d_gz + SS_NGZ, 0xFF, sizeof(unsigned int), c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}

static void gz_sync(ss_ctx *c) { for (int i = 0; i < SS_NGZ; i++) cudaStreamSynchronize(c->gz_stream[i]); }

// after the stream was synchronised: did every member inflate to its ISIZE?
static int gz_check(ss_ctx *c, const char *what) {
    unsigned int bad = 0xFFFFFFFFu;
    gz_sync(c);
    SS_CUDA(cudaMemcpy(&bad, c->d_gz + SS_NGZ, sizeof bad, cudaMemcpyDeviceToHost));
    if (bad != 0xFFFFFFFFu)
        return fail(SS_ERR_IO, std::string(what) + ": inflate failed on the device (invalid BGZF block, member " +
                                   std::to_string(bad) + " of a batch)");
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// ordinary gzip read files inflated on the device (ss_dgz.cu)
// ---------------------------------------------------------------------------------------------
// A .gz read file that is not blocked gzip: the compressed bytes are uploaded as they are (3-4x fewer than the text
// cross PCIe) and inflated by thousands of decoders at once (block starts found per 128 KiB piece, marker symbols for
// the unknown window, exact stitching).  Replaces the `zcat a b |` of identify.py:82 / Vote_...:359,367, which the host
// threads of ss_ingest.cu otherwise stand in for.  With several ranks a file of many members is split at member
// starts exactly as ss_ingest.cu splits it (same parts, same ownership of the record that straddles a boundary);
// a single-member file is left to the host path (every rank decodes it, chunks dealt round-robin).
struct dgz_file {
    std::string path;
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0;
    size_t lo = 0, hi = 0;            // members that START in [lo, hi) are mine
    size_t first_member = 0;          // the first of them (== size: none)
    size_t up_hi = 0;                 // compressed bytes [first_member, up_hi) are uploaded
    double ratio = 4.0;               // text bytes per compressed byte at the head of the file (sizes the decoder's buffers)
    void close_map() {
        if (map) munmap((void *)map, size);
        if (fd >= 0) close(fd);
        map = nullptr; fd = -1;
    }
};

static bool dgz_enabled() {
    const char *e = getenv("SS_DGZ");
    return !e || atoi(e) != 0;
}

// Does `path` take the device gzip path for this shard?  On true the file stays mapped in `df`.
static bool dgz_eligible(ss_ctx *c, const char *path, int shard, int n_shards, dgz_file &df) {
    if (!dgz_enabled() || !c->device_bgzf) return false;
    df = dgz_file();
    df.path = path;
    df.fd = open(path, O_RDONLY);
    if (df.fd < 0) return false;                                   // the text source reports the error
    struct stat st;
    if (fstat(df.fd, &st) != 0) { df.close_map(); return false; }
    df.size = (size_t)st.st_size;
    size_t min_bytes = 8u << 20;
    if (const char *e = getenv("SS_DGZ_MIN_BYTES")) { long long v = atoll(e); if (v >= 0) min_bytes = (size_t)v; }
    if (df.size < 64 || df.size < min_bytes) { df.close_map(); return false; }
    void *m = mmap(nullptr, df.size, PROT_READ, MAP_PRIVATE, df.fd, 0);
    if (m == MAP_FAILED) { df.map = nullptr; df.close_map(); return false; }
    df.map = (const uint8_t *)m;
    ssi_gz_header h;
    const char *dot = strrchr(path, '.');
    const bool gz = (dot && strcmp(dot + 1, "gz") == 0) || (df.map[0] == 0x1f && df.map[1] == 0x8b);
    if (!gz || ssi_gz_parse_header(df.map, df.map + df.size, &h) != SSI_OK || h.bgzf_bsize) { df.close_map(); return false; }
    {   // 4-line FASTQ?  (FASTA / wrapped FASTQ are rewritten on the host)
        std::vector<uint8_t> head(SS_INGEST_HIST + (64u << 10));
        ssi_gz_stream *g = new ssi_gz_stream;
        ssi_gz_init(*g, df.map, df.size);
        uint8_t *pos = head.data() + SS_INGEST_HIST;
        int rc = ssi_gz_read(*g, &pos, head.data() + head.size());
        const size_t hn = (size_t)(pos - (head.data() + SS_INGEST_HIST));
        const size_t used = g->in_member ? (size_t)(ssi_in_pos(g->s.bits) - df.map) : (size_t)(g->p - df.map);
        delete g;
        if (rc < 0 || hn == 0 || ss_fastx_kind((const char *)head.data() + SS_INGEST_HIST, hn, rc == SSI_OK) != 0) { df.close_map(); return false; }
        if (used > h.header_len + 64) df.ratio = (double)hn / (double)(used - h.header_len);
    }
    df.lo = 0; df.hi = df.size;
    if (n_shards > 1) {
        if (!ss_gz_is_member_split(df.map, df.size, n_shards)) { df.close_map(); return false; }
        const size_t P = ss_gz_split_parts(df.size);
        const size_t p_lo = P * (size_t)shard / (size_t)n_shards, p_hi = P * ((size_t)shard + 1) / (size_t)n_shards;
        df.lo = (size_t)((unsigned __int128)df.size * p_lo / P);
        df.hi = p_hi >= P ? df.size : (size_t)((unsigned __int128)df.size * p_hi / P);
    }
    df.first_member = df.lo == 0 ? 0 : (df.hi > df.lo ? ss_gz_next_member_start(df.map, df.size, df.lo, df.hi) : df.size);
    if (df.first_member >= df.hi) df.first_member = df.size;      // no member of mine: nothing to do, but the file is handled
    // everything up to the end of the member that straddles `hi`
    df.up_hi = df.hi >= df.size ? df.size : ss_gz_next_member_start(df.map, df.size, df.hi, df.size);
    if (df.up_hi < df.size) df.up_hi = std::min(df.size, df.up_hi + 64);
    // room for what the device path still has to allocate: the decoder's buffers for this file's plan, the batch buffer,
    // an upload buffer (what the context already holds from earlier files counts as there)
    size_t free_b = 0, total_b = 0;
    const size_t n_up = df.up_hi - std::min(df.up_hi, df.first_member);
    const ss_dgz_plan plan = ss_dgz_make_plan(n_up, c->n_sm, ss_dgz::shape_from_env(), df.ratio);
    const size_t held = (c->dgz ? c->dgz->held_bytes() : 0) + c->dgz_scratch_cap;
    size_t need = (plan.device_bytes > held ? plan.device_bytes - held : 0) + (1ull << 30);
    if (n_up + 128 > std::min(c->dgz_up_cap[0], c->dgz_up_cap[1])) need += n_up + 128;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < need) {
        df.close_map();
        return false;                                                // not enough room beside the read cache: host threads
    }
    return true;
}

// Inflate one file on the device and hand its text out in pieces of whole FASTQ records:
// sink(const uint8_t *d_text, size_t n) -> SS_* code.  `d_text` stays valid until the sink returns.
// the compressed bytes of one file (range) on their way to the device
struct dgz_upload {
    uint8_t *d_alloc = nullptr;       // 64 bytes of zero padding, the bytes, 64 bytes of zero padding (allocated by the caller)
    size_t n_up = 0;
    double ms = 0;
    int rc = SS_OK;
    std::string msg;
    // bytes [0, watermark) of the range are on the device; `finished` once the upload ended (well or badly)
    std::mutex mu;
    std::condition_variable cv;
    size_t watermark = 0;
    bool finished = false;
    // block until bytes [0, need) are on the device (or the upload ended); returns false when the upload failed
    bool wait(size_t need) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return finished || watermark >= std::min(need, n_up); });
        return rc == SS_OK;
    }
    void finish(int code, const std::string &m) {
        std::lock_guard<std::mutex> lk(mu);
        if (code && !rc) { rc = code; msg = m; }
        if (!code) watermark = n_up;
        finished = true;
        cv.notify_all();
    }
};

// Upload [first_member, up_hi) of the file: several readers, pinned chunks, any order, H2D on the copy stream; the
// watermark follows the completed copies, so the inflate of the SAME file starts on its first batches while the rest
// is still on its way (and the upload of the next file runs beside the inflate of this one).  Own thread: touches the
// text source and the copy stream only.
static void dgz_upload_file(ss_ctx *c, const dgz_file &df, dgz_upload &up) {
    if (df.first_member >= df.size || !up.d_alloc) { up.finish(SS_OK, ""); return; }
    cudaSetDevice(c->device);
    const double t0 = now_ms();
    const size_t base_off = df.first_member, n_up = up.n_up;
    uint8_t *d_comp = up.d_alloc + 64;
    cudaMemsetAsync(up.d_alloc, 0, 64, c->copy_stream);
    cudaMemsetAsync(d_comp + n_up, 0, 64, c->copy_stream);
    int rc = c->src->start_raw(df.path.c_str(), base_off, df.up_hi);
    if (rc) { up.finish(rc, c->src->error()); return; }
    const size_t chunk = c->src->chunk_bytes();
    std::vector<uint8_t> done((n_up + chunk - 1) / chunk, 0);
    size_t next_chunk = 0;
    pending_ring pend(c);
    pend.on_done = [&](ss_chunk *ch) {
        done[(size_t)(ch->file_off - base_off) / chunk] = 1;
        size_t w = next_chunk;
        while (w < done.size() && done[w]) w++;
        if (w != next_chunk) {
            next_chunk = w;
            std::lock_guard<std::mutex> lk(up.mu);
            up.watermark = std::min(n_up, w * chunk);
            up.cv.notify_all();
        }
    };
    cudaError_t ce = cudaSuccess;
    while (ss_chunk *ch = c->src->next()) {
        ce = cudaMemcpyAsync(d_comp + (ch->file_off - base_off), ch->text, ch->len, cudaMemcpyHostToDevice, c->copy_stream);
        if (ce != cudaSuccess) { c->src->release(ch); break; }
        ce = pend.push(ch, c->copy_stream);
        if (ce != cudaSuccess) break;
    }
    pend.drain();
    int src_rc = c->src->finish();
    cudaStreamSynchronize(c->copy_stream);
    up.ms = now_ms() - t0;
    if (ce != cudaSuccess) { up.finish(SS_ERR_CUDA, std::string("CUDA error (") + cudaGetErrorString(ce) + ") in the upload of " + df.path); return; }
    if (src_rc) { up.finish(src_rc, c->src->error()); return; }
    up.finish(SS_OK, "");
}

// Inflate one uploaded file on the device and hand its text out in pieces of whole FASTQ records:
// sink(const uint8_t *d_text, size_t n) -> SS_* code.  `d_text` stays valid until the sink returns.
template <typename Sink>
static int dgz_inflate_file(ss_ctx *c, dgz_file &df, dgz_upload &up, Sink &&sink) {
    if (df.first_member >= df.size) return SS_OK;
    int rc = SS_OK;
    const size_t base_off = df.first_member, n_up = up.n_up;
    uint8_t *d_scratch = nullptr;
    const ss_dgz_plan plan = ss_dgz_make_plan(n_up, c->n_sm, ss_dgz::shape_from_env(), df.ratio);
    const size_t scratch_cap = plan.scratch;
    auto cleanup = [&]() {};
#define DGZ_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return ss_cuda_fail(e_, #x, __FILE__, __LINE__); } } while (0)
    if (c->dgz_scratch_cap != scratch_cap) {
        cudaFree(c->d_dgz_scratch); c->d_dgz_scratch = nullptr; c->dgz_scratch_cap = 0;
        DGZ_TRY(cudaMalloc(&c->d_dgz_scratch, scratch_cap + SS_TEXT_PAD + SS_TILE));
        c->dgz_scratch_cap = scratch_cap;
    }
    d_scratch = c->d_dgz_scratch;
    uint8_t *d_comp = up.d_alloc + 64;
    const bool dbg = getenv("SS_DEBUG_TIMING") != nullptr;
    double t_up = up.ms, t_inflate = 0, t_cut = 0, t_sink = 0;
    // ---- inflate, batch by batch
    if (!c->dgz) c->dgz = new ss_dgz();
    ss_dgz &dz = *c->dgz;
    // a batch is decoded once its compressed bytes (and SS_DGZ_GATE_SLACK behind them) have arrived
    dz.set_input_gate([&up, base_off](size_t need_byte) { return up.wait(need_byte > base_off ? need_byte - base_off : 0); },
                      [&up]() { std::lock_guard<std::mutex> lk(up.mu); return up.finished; });
    rc = dz.open(c->n_sm, c->stream, d_comp - base_off, df.map, df.up_hi, df.first_member, df.hi, plan);
    if (rc) { cleanup(); return fail(rc, "inflate failed on " + df.path + ": " + dz.error()); }
    std::vector<char> h_win(1u << 20);
    size_t carry = 0;
    bool first = true, done = false;
    const std::string no_boundary = df.path + ": no FASTQ record boundary within a batch of the device inflate";
    while (!done) {
        size_t n = 0;
        double t0 = now_ms();
        rc = dz.next(d_scratch + carry, scratch_cap - carry, &n, &done);
        if (rc) { cleanup(); return fail(rc, "inflate failed on " + df.path + ": " + dz.error()); }
        t_inflate += now_ms() - t0; t0 = now_ms();
        size_t len = carry + n, start = 0;
        if (first) {
            const size_t w = std::min(len, h_win.size());
            DGZ_TRY(cudaMemcpy(h_win.data(), d_scratch, w, cudaMemcpyDeviceToHost));
            if (df.lo == 0) {
                rc = check_fastq_head(h_win.data(), w, df.path.c_str());
                if (rc) { cleanup(); return rc; }
            } else {                                                 // the head belongs to the previous part's owner
                start = ss_find_record_start(h_win.data(), w, 1);
                if (start >= w) {
                    if (!done || w < len) { cleanup(); return fail(SS_ERR_FORMAT, no_boundary); }
                    start = len;
                }
            }
            first = false;
        }
        size_t cut = len;
        if (!done) {
            // the last record start of the batch: searched in a window at its end that widens
            cut = 0;
            for (size_t win = 256u << 10; cut == 0; win *= 4) {
                const size_t w = std::min(win, len - start);
                if (w > h_win.size()) h_win.resize(w);
                DGZ_TRY(cudaMemcpy(h_win.data(), d_scratch + len - w, w, cudaMemcpyDeviceToHost));
                const size_t r = ss_find_cut(h_win.data(), 0, w);
                if (r) cut = len - w + r;
                else if (w == len - start) break;
            }
            if (cut == 0) {
                if (len - start + (scratch_cap >> 2) > scratch_cap) { cleanup(); return fail(SS_ERR_FORMAT, no_boundary); }
                cut = start;                                         // nothing whole yet: keep everything, decode on
            }
        } else {
            // the end of my members.  Another part follows: its owner drops the head of its first member up to the
            // first record start behind the member's first byte, so that head is mine (decoded here, on the host).
            const size_t at = dz.stopped_at();
            if (df.hi < df.size && at < df.size && at >= df.hi) {
                std::vector<uint8_t> head(SS_INGEST_HIST + (256u << 10));
                bool found = false;
                while (!found) {
                    ssi_gz_stream *g = new ssi_gz_stream;
                    ssi_gz_init(*g, df.map + at, df.size - at);
                    uint8_t *pos = head.data() + SS_INGEST_HIST;
                    int irc = ssi_gz_read(*g, &pos, head.data() + head.size());
                    delete g;
                    const size_t hn = (size_t)(pos - (head.data() + SS_INGEST_HIST));
                    if (irc < 0 && irc != SSI_ERR_HEADER) { cleanup(); return fail(SS_ERR_IO, "inflate failed on " + df.path + ": invalid compressed data"); }
                    size_t r = ss_find_record_start((const char *)head.data() + SS_INGEST_HIST, hn, 1);
                    if (r < hn || irc == SSI_OK) {
                        if (r > hn) r = hn;
                        if (len + r > scratch_cap) { cleanup(); return fail(SS_ERR_FORMAT, no_boundary); }
                        DGZ_TRY(cudaMemcpy(d_scratch + len, head.data() + SS_INGEST_HIST, r, cudaMemcpyHostToDevice));
                        len += r;
                        found = true;
                    } else if (head.size() > (256u << 20)) { cleanup(); return fail(SS_ERR_FORMAT, no_boundary); }
                    else head.resize(SS_INGEST_HIST + (head.size() - SS_INGEST_HIST) * 4);
                }
                cut = len;
            } else {
                // the file ends here: blank tail lines go, the last line is closed (as the host producers do)
                const size_t w = std::min<size_t>(len - start, 4096);
                DGZ_TRY(cudaMemcpy(h_win.data(), d_scratch + len - w, w, cudaMemcpyDeviceToHost));
                size_t keep = ss_trim_tail(h_win.data(), w);
                if (keep == 0 && w < len - start) {                  // 4 KiB of blank lines: look at everything (rare)
                    std::vector<char> all(len - start);
                    DGZ_TRY(cudaMemcpy(all.data(), d_scratch + start, len - start, cudaMemcpyDeviceToHost));
                    keep = ss_trim_tail(all.data(), len - start);
                    len = start + keep;
                } else len = len - w + keep;
                if (len > start) { DGZ_TRY(cudaMemsetAsync(d_scratch + len, '\n', 1, c->stream)); len++; }
                cut = len;
            }
        }
        // the partial record behind the cut is set aside (the sink may pad behind its text), then moved to the front
        carry = len - cut;
        if (carry) {
            if (carry > c->dgz_carry_cap) {
                cudaFree(c->d_dgz_carry); c->d_dgz_carry = nullptr; c->dgz_carry_cap = 0;
                const size_t cap = std::max<size_t>(carry, 1u << 20);
                DGZ_TRY(cudaMalloc(&c->d_dgz_carry, cap));
                c->dgz_carry_cap = cap;
            }
            DGZ_TRY(cudaMemcpyAsync(c->d_dgz_carry, d_scratch + cut, carry, cudaMemcpyDeviceToDevice, c->stream));
        }
        if (cut > start) {
            const uint8_t *text = d_scratch + start;
            uint8_t *moved = nullptr;
            if (start & 255u) {                                      // (first batch of a later part only) the kernels want an aligned text
                DGZ_TRY(cudaMalloc(&moved, cut - start));
                DGZ_TRY(cudaMemcpyAsync(moved, text, cut - start, cudaMemcpyDeviceToDevice, c->stream));
                DGZ_TRY(cudaMemcpyAsync(d_scratch, moved, cut - start, cudaMemcpyDeviceToDevice, c->stream));
                DGZ_TRY(cudaStreamSynchronize(c->stream));
                cudaFree(moved);
                text = d_scratch;
            }
            t_cut += now_ms() - t0; t0 = now_ms();
            rc = sink(text, cut - start);
            if (rc) { cleanup(); return rc; }
            t_sink += now_ms() - t0;
        }
        if (carry) DGZ_TRY(cudaMemcpyAsync(d_scratch, c->d_dgz_carry, carry, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (dbg)
        fprintf(stderr, "[ss dgz] %s: %llu batches, %llu/%llu pieces used, %llu members; upload %.1f ms (%.1f MB), inflate %.1f ms "
                "(find+decode %.1f, windows+resolve %.1f of which windows %.1f), cuts %.1f ms, sink %.1f ms\n", df.path.c_str(), (unsigned long long)dz.batches(),
                (unsigned long long)dz.pieces_used(), (unsigned long long)(dz.pieces_found() + dz.batches()), (unsigned long long)dz.members(),
                t_up, n_up / 1e6, t_inflate, dz.ms_decode(), dz.ms_resolve(), dz.ms_windows(), t_cut, t_sink);
#undef DGZ_TRY
    cudaStreamSynchronize(c->stream);
    cleanup();
    return SS_OK;
}

// All device-gzip files of a call: ONE uploader thread sends the files up one after the other while this thread
// inflates them in the same order, each batch as soon as its bytes are there.  Two upload buffers (kept by the context,
// growing only: no allocation in a steady run): file i goes into buffer i % 2 once file i - 2 has been inflated.
template <typename Sink>
static int dgz_run_files(ss_ctx *c, std::vector<dgz_file> &files, Sink &&sink) {
    if (files.empty()) return SS_OK;
    const bool dbg = getenv("SS_DEBUG_TIMING") != nullptr;
    const double t_begin = now_ms();
    int rc = ensure_source(c);
    std::vector<std::unique_ptr<dgz_upload>> ups;
    size_t need[2] = {0, 0};
    for (size_t i = 0; i < files.size(); i++) {
        ups.emplace_back(new dgz_upload());
        if (files[i].first_member >= files[i].size) continue;
        ups.back()->n_up = files[i].up_hi - files[i].first_member;
        need[i & 1] = std::max(need[i & 1], ups.back()->n_up + 128);
    }
    for (int k = 0; k < 2 && !rc; k++)
        if (need[k] > c->dgz_up_cap[k]) {
            cudaFree(c->d_dgz_up[k]); c->d_dgz_up[k] = nullptr; c->dgz_up_cap[k] = 0;
            cudaError_t e = cudaMalloc(&c->d_dgz_up[k], need[k]);
            if (e != cudaSuccess) rc = ss_cuda_fail(e, "cudaMalloc(compressed reads)", __FILE__, __LINE__);
            else c->dgz_up_cap[k] = need[k];
        }
    const double t_alloc = now_ms();
    if (!rc) {
        for (size_t i = 0; i < files.size(); i++) if (ups[i]->n_up) ups[i]->d_alloc = c->d_dgz_up[i & 1];
        std::mutex mu;
        std::condition_variable cv;
        size_t inflated = 0;                                          // files [0, inflated) are done with their buffers
        bool stop = false;
        std::thread uploader([&]() {
            for (size_t i = 0; i < files.size(); i++) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&]() { return stop || i < inflated + 2; });
                    if (stop) { for (size_t j = i; j < files.size(); j++) ups[j]->finish(SS_ERR_IO, "cancelled"); return; }
                }
                dgz_upload_file(c, files[i], *ups[i]);
            }
        });
        for (size_t i = 0; i < files.size(); i++) {
            dgz_upload &u = *ups[i];
            if (!rc) {
                rc = dgz_inflate_file(c, files[i], u, sink);
                u.wait(u.n_up);                                      // (an inflate that failed early: let this upload run out)
                if (!rc && u.rc) rc = fail(u.rc, u.msg);
            }
            std::lock_guard<std::mutex> lk(mu);
            if (rc) stop = true;                                     // no further uploads
            inflated = i + 1;
            cv.notify_all();
        }
        uploader.join();
        cudaStreamSynchronize(c->copy_stream);
    }
    for (auto &f : files) f.close_map();
    if (dbg) fprintf(stderr, "[ss dgz] %zu files: %.1f ms in all (buffers %.1f ms)\n", files.size(), now_ms() - t_begin, t_alloc - t_begin);
    return rc;
}

extern "C" int ss_reads_from_files(ss_ctx *c, const char *const *paths, int n_paths, int shard, int n_shards,
                                   ss_reads **out) {
    if (!c || !out || (n_paths > 0 && !paths)) return fail(SS_ERR_ARG, "ss_reads_from_files: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    *out = nullptr;
    SS_CUDA(cudaSetDevice(c->device));
    int rc = ensure_source(c);
    if (rc) return rc;
    ss_text_source &src = *c->src;
    ss_reads *r = new ss_reads();
    r->ctx = c;
    // ordinary gzip files big enough to pay off are inflated on the device, one after the other (ss_dgz.cu)
    std::vector<const char *> rest;
    std::vector<dgz_file> dgz_files;
    for (int i = 0; i < n_paths; i++) {
        dgz_file df;
        if (dgz_eligible(c, paths[i], shard, n_shards, df)) dgz_files.push_back(df);
        else rest.push_back(paths[i]);
    }
    rc = dgz_run_files(c, dgz_files, [&](const uint8_t *d_text, size_t n) {
        if (r->seg.empty() || r->seg.back().len + n + SSI_OUT_SLACK > r->seg.back().cap - SS_TEXT_PAD - SS_TILE) {
            r->seg.emplace_back();
            int arc = alloc_segment(r->seg.back(), std::max<uint64_t>(c->seg_bytes, n + SSI_OUT_SLACK));
            if (arc) { r->seg.pop_back(); return arc; }
        }
        ss_segment &g = r->seg.back();
        SS_CUDA(cudaMemcpyAsync(g.d_text + g.len, d_text, n, cudaMemcpyDeviceToDevice, c->stream));
        SS_CUDA(cudaStreamSynchronize(c->stream));
        g.len += n; r->len += n;
        return (int)SS_OK;
    });
    if (rc) { ss_reads_free(r); return rc; }
    paths = rest.data(); n_paths = (int)rest.size();
    rc = src.start(paths, n_paths, shard, n_shards);
    if (rc) { ss_reads_free(r); return fail(rc, src.error()); }
    // plain inputs have a known size (one segment); gzip streams grow segment by segment
    const uint64_t known = src.plain_bytes() + 64 * (uint64_t)n_paths + 4096;
    const uint64_t seg_default = src.gz_bytes() ? std::max<uint64_t>(c->seg_bytes, c->chunk_bytes) : known;
    pending_ring pend(c);
    cudaError_t ce = cudaSuccess;
    rc = gz_begin(c);
    bool any_gz = false;
    const bool dbg = getenv("SS_DEBUG_TIMING") != nullptr;
    double t_begin = now_ms(), t_wait = 0;
    uint32_t n_chunks = 0;
    while (!rc) {
        double tw = now_ms();
        ss_chunk *ch = src.next();
        t_wait += now_ms() - tw;
        if (!ch) break;
        n_chunks++;
        const uint64_t n = ch->text_len();
        if (r->seg.empty() || r->seg.back().len + n + SSI_OUT_SLACK > r->seg.back().cap - SS_TEXT_PAD - SS_TILE) {
            r->seg.emplace_back();
            rc = alloc_segment(r->seg.back(), std::max<uint64_t>(seg_default, n + SSI_OUT_SLACK));
            if (rc) { r->seg.pop_back(); src.release(ch); break; }
        }
        ss_segment &g = r->seg.back();
        if (ch->kind == SS_CHUNK_BGZF) {
            rc = inflate_batch(c, ch, g.d_text + g.len, nullptr);     // the text is produced in place, in the cache segment
            any_gz = true;
            if (rc) { src.release(ch); break; }
        } else {
            ce = cudaMemcpyAsync(g.d_text + g.len, ch->text, ch->len, cudaMemcpyHostToDevice, c->copy_stream);
            if (ce != cudaSuccess) { src.release(ch); break; }
        }
        ce = pend.push(ch, c->copy_stream);
        if (ce != cudaSuccess) break;
        g.len += n;
        r->len += n;
    }
    double t_loop = now_ms();
    pend.drain();
    int src_rc = src.finish();
    cudaStreamSynchronize(c->copy_stream);
    gz_sync(c);
    cudaStreamSynchronize(c->stream);
    if (dbg)
        fprintf(stderr, "[ss ingest] %u chunks, %.1f MB text: loop %.1f ms (waiting for producers %.1f ms), drain+sync %.1f ms\n",
                n_chunks, r->len / 1e6, t_loop - t_begin, t_wait, now_ms() - t_loop);
    if (ce != cudaSuccess) { ss_reads_free(r); return ss_cuda_fail(ce, "H2D reads", __FILE__, __LINE__); }
    if (rc) { ss_reads_free(r); return rc; }
    if (src_rc) { ss_reads_free(r); return fail(src_rc, src.error()); }
    if (any_gz) { rc = gz_check(c, "ss_reads_from_files"); if (rc) { ss_reads_free(r); return rc; } }
    for (auto &g : r->seg) {
        if (g.cap > ss_reads_device_capacity(g.len) + (64ull << 20)) {      // shrink a mostly empty segment
            ss_segment small;
            if (alloc_segment(small, g.len) == SS_OK) {
                small.len = g.len;
                if (g.len) cudaMemcpy(small.d_text, g.d_text, g.len, cudaMemcpyDeviceToDevice);
                free_segment(g);
                g = small;
            }
        }
        rc = finish_segment(c, g);
        if (rc) { ss_reads_free(r); return rc; }
    }
    *out = r;
    return SS_OK;
}

// host-only: the ingest without a GPU (tools, CPU tests of the chunker / inflate)
extern "C" int ss_ingest_files_host(const char *const *paths, int n_paths, int shard, int n_shards, size_t chunk_bytes,
                                    int n_threads, char *out, size_t out_cap, size_t *out_len, uint32_t *n_chunks) {
    if ((n_paths > 0 && !paths) || !out_len) return fail(SS_ERR_ARG, "ss_ingest_files_host: NULL argument");
    ss_text_source src;
    // BGZF inputs take the same producer path as on the GPU (member batches, host-decoded boundary
    // members); the per-member inflate the device kernel would run is executed here with the same code
    bool bgzf = true;
    size_t bgzf_cap = 6 * chunk_bytes;
    if (const char *e = getenv("SS_BGZF_GPU")) bgzf = atoi(e) != 0;
    if (const char *e = getenv("SS_BGZF_OUT_CAP")) { long long v = atoll(e); if (v >= (1 << 20)) bgzf_cap = (size_t)v; }
    int rc = src.init(chunk_bytes, std::max(1, n_threads) + 3, n_threads, false, bgzf, bgzf_cap);
    if (rc) return fail(rc, src.error());
    rc = src.start(paths, n_paths, shard, n_shards);
    if (rc) return fail(rc, src.error());
    size_t total = 0;
    uint32_t n = 0;
    bool overflow = false, bad_member = false;
    ssi_tables *tabs = new ssi_tables;
    while (ss_chunk *ch = src.next()) {
        const size_t len = ch->text_len();
        if (!out) {
        } else if (total + len + SSI_OUT_SLACK > out_cap) {
            overflow = true;
        } else if (ch->kind == SS_CHUNK_BGZF) {
            uint8_t *dst = (uint8_t *)out + total;
            memcpy(dst, ch->pre_text, ch->pre_len);
            for (uint32_t i = 0; i < ch->n_members; i++) {
                const ss_member &m = ch->members[i];
                ssi_stream st;
                ssi_stream_init(st, ch->text + m.comp_off, ch->text + m.comp_off + m.comp_len);
                uint8_t *w = dst + m.out_off;
                int irc = ssi_inflate(st, *tabs, &w, dst + m.out_off + m.isize + SSI_OUT_SLACK);
                if (irc != SSI_OK || st.out_total != m.isize) bad_member = true;
            }
            memcpy(dst + ch->pre_len + ch->inflated_len, ch->post_text, ch->post_len);
        } else {
            memcpy(out + total, ch->text, ch->len);
        }
        total += len;
        n++;
        src.release(ch);
    }
    delete tabs;
    rc = src.finish();
    if (rc) return fail(rc, src.error());
    if (bad_member) return fail(SS_ERR_IO, "ss_ingest_files_host: inflate failed (invalid BGZF block)");
    *out_len = total;
    if (n_chunks) *n_chunks = n;
    if (overflow) return fail(SS_ERR_ARG, "ss_ingest_files_host: output buffer too small");
    return SS_OK;
}

// host-only: the device gzip inflate pipeline (piece search, marker decode, chain, windows, resolve) on the CPU
extern "C" int ss_dgz_inflate_host(const char *comp, size_t comp_size, size_t first_member, size_t stop_member_at,
                                   uint32_t max_pieces, uint32_t piece_bytes, uint32_t sym_per_byte, char *out, size_t out_cap,
                                   size_t *out_len, size_t *stopped_at, uint64_t *stats4) {
    if ((!comp && comp_size) || !out_len) return fail(SS_ERR_ARG, "ss_dgz_inflate_host: NULL argument");
    std::vector<uint8_t> text;
    std::string err;
    int rc = ss_dgz_host_inflate((const uint8_t *)comp, comp_size, first_member, stop_member_at, max_pieces, piece_bytes,
                                 sym_per_byte ? sym_per_byte : 5u, text, stopped_at, stats4, err);
    if (rc) return fail(rc, "device gzip inflate (host emulation): " + err);
    *out_len = text.size();
    if (out) {
        if (text.size() > out_cap) return fail(SS_ERR_ARG, "ss_dgz_inflate_host: output buffer too small");
        if (!text.empty()) memcpy(out, text.data(), text.size());
    }
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// host-only: index lists of the non-zeros of a dense 2-D array (the mirrors of the per-strain reducers take the dense
// 0/1 strain matrices the reference builds -- pX = X.A, identify_strains_L2_Enet_Pscan_new_sp.py:200-201 -- and need
// them as CSC; NumPy's flatnonzero, strain by strain, is one thread and was most of the mirrors' time)
// ---------------------------------------------------------------------------------------------
namespace {
template <typename T> inline bool dn_nonzero(const T *p, uint64_t i) { return p[i] != (T)0; }

// lists along axis 0: list j = the axis-1 indices i with X[j][i] != 0
template <typename T>
void dn_axis0(const T *X, uint64_t n0, uint64_t n1, int n_threads, uint64_t *ptr, uint32_t *idx) {
    std::vector<std::thread> th;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (uint64_t j; (j = next.fetch_add(1)) < n0;) {
            const T *row = X + j * n1;
            if (!idx) {
                uint64_t c = 0;
                for (uint64_t i = 0; i < n1; i++) c += dn_nonzero(row, i) ? 1u : 0u;
                ptr[j + 1] = c;
            } else {
                uint32_t *out = idx + ptr[j];
                for (uint64_t i = 0; i < n1; i++) if (dn_nonzero(row, i)) *out++ = (uint32_t)i;
            }
        }
    };
    for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

// lists along axis 1: list s = the axis-0 indices r with X[r][s] != 0, ascending (blocks of rows per thread, in order)
template <typename T>
void dn_axis1(const T *X, uint64_t n0, uint64_t n1, int n_threads, uint64_t *ptr, uint32_t *idx) {
    const uint64_t T_ = (uint64_t)std::max(1, n_threads);
    std::vector<std::vector<uint64_t>> cnt(T_, std::vector<uint64_t>(n1, 0));
    auto block = [&](uint64_t t, uint64_t &lo, uint64_t &hi) { lo = n0 * t / T_; hi = n0 * (t + 1) / T_; };
    {
        std::vector<std::thread> th;
        auto work = [&](uint64_t t) {
            uint64_t lo, hi; block(t, lo, hi);
            uint64_t *c = cnt[t].data();
            for (uint64_t r = lo; r < hi; r++) { const T *row = X + r * n1; for (uint64_t s2 = 0; s2 < n1; s2++) c[s2] += dn_nonzero(row, s2) ? 1u : 0u; }
        };
        for (uint64_t t = 1; t < T_; t++) th.emplace_back(work, t);
        work(0);
        for (auto &t : th) t.join();
    }
    if (!idx) {
        for (uint64_t s2 = 0; s2 < n1; s2++) { uint64_t c = 0; for (uint64_t t = 0; t < T_; t++) c += cnt[t][s2]; ptr[s2 + 1] = c; }
        return;
    }
    // where each thread's block starts inside every list
    for (uint64_t s2 = 0; s2 < n1; s2++) { uint64_t at = ptr[s2]; for (uint64_t t = 0; t < T_; t++) { const uint64_t c = cnt[t][s2]; cnt[t][s2] = at; at += c; } }
    std::vector<std::thread> th;
    auto work = [&](uint64_t t) {
        uint64_t lo, hi; block(t, lo, hi);
        uint64_t *at = cnt[t].data();
        for (uint64_t r = lo; r < hi; r++) { const T *row = X + r * n1; for (uint64_t s2 = 0; s2 < n1; s2++) if (dn_nonzero(row, s2)) idx[at[s2]++] = (uint32_t)r; }
    };
    for (uint64_t t = 1; t < T_; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
}

template <typename T>
void dn_run(const void *X, uint64_t n0, uint64_t n1, int axis, int n_threads, uint64_t *ptr, uint32_t *idx) {
    if (axis == 0) dn_axis0((const T *)X, n0, n1, n_threads, ptr, idx);
    else dn_axis1((const T *)X, n0, n1, n_threads, ptr, idx);
}
}   // namespace

// X: C-contiguous n0 x n1 array of `elem_bytes`-byte integers (is_float = 0; bool counts) or IEEE floats (is_float = 1).
// axis 0: n0 lists of axis-1 indices; axis 1: n1 lists of axis-0 indices (ascending).  Call with idx == NULL first: ptr[k + 1]
// receives the length of list k (ptr[0] untouched); turn ptr into offsets and call again with idx of ptr[last] entries.
extern "C" int ss_dense_nonzero_lists(const void *X, uint64_t n0, uint64_t n1, int elem_bytes, int is_float, int axis,
                                      int n_threads, uint64_t *ptr, uint32_t *idx) {
    if ((!X && n0 && n1) || !ptr || (axis != 0 && axis != 1)) return fail(SS_ERR_ARG, "ss_dense_nonzero_lists: bad argument");
    if ((axis == 0 ? n1 : n0) > 0xFFFFFFFFull) return fail(SS_ERR_ARG, "ss_dense_nonzero_lists: more than 2^32 entries along the indexed axis");
    if (n_threads < 1) {
        cpu_set_t set;
        n_threads = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        n_threads = std::max(1, std::min(n_threads, 32));
    }
    if (n0 == 0 || n1 == 0) { if (!idx) for (uint64_t k = 0; k < (axis == 0 ? n0 : n1); k++) ptr[k + 1] = 0; return SS_OK; }
    if (is_float) {
        if (elem_bytes == 4) dn_run<float>(X, n0, n1, axis, n_threads, ptr, idx);
        else if (elem_bytes == 8) dn_run<double>(X, n0, n1, axis, n_threads, ptr, idx);
        else return fail(SS_ERR_ARG, "ss_dense_nonzero_lists: float element size must be 4 or 8");
    } else {
        if (elem_bytes == 1) dn_run<uint8_t>(X, n0, n1, axis, n_threads, ptr, idx);
        else if (elem_bytes == 2) dn_run<uint16_t>(X, n0, n1, axis, n_threads, ptr, idx);
        else if (elem_bytes == 4) dn_run<uint32_t>(X, n0, n1, axis, n_threads, ptr, idx);
        else if (elem_bytes == 8) dn_run<uint64_t>(X, n0, n1, axis, n_threads, ptr, idx);
        else return fail(SS_ERR_ARG, "ss_dense_nonzero_lists: integer element size must be 1, 2, 4 or 8");
    }
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// host-only: the k-mer lists of many tree nodes (Tree_database/kmers/<node>: one line of record ordinals separated by
// single spaces, read by identify.py:118 as list(map(int, line.rstrip().split(" "))) and turned into a set) parsed and
// de-duplicated on all host threads.  A file that is not exactly that shape is reported, not guessed at: the caller
// parses it the slow way with the reference's own semantics.
// ---------------------------------------------------------------------------------------------
enum { SS_NODE_LIST_OK = 0, SS_NODE_LIST_EMPTY = 1, SS_NODE_LIST_UNUSUAL = 2, SS_NODE_LIST_UNREADABLE = 3 };

static int parse_node_list(const char *path, std::vector<uint32_t> &out) {
    out.clear();
    int fd = open(path, O_RDONLY);
    if (fd < 0) return SS_NODE_LIST_UNREADABLE;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return SS_NODE_LIST_UNREADABLE; }
    const size_t size = (size_t)st.st_size;
    if (size == 0) { close(fd); return SS_NODE_LIST_EMPTY; }
    std::vector<char> buf(size);
    for (size_t have = 0; have < size;) {
        ssize_t got = pread(fd, buf.data() + have, size - have, (off_t)have);
        if (got <= 0) { close(fd); return SS_NODE_LIST_UNREADABLE; }
        have += (size_t)got;
    }
    close(fd);
    size_t end = 0;
    while (end < size && buf[end] != '\n') end++;                    // lines[0]
    while (end > 0 && (buf[end - 1] == ' ' || buf[end - 1] == '\r' || buf[end - 1] == '\t' || buf[end - 1] == '\f' || buf[end - 1] == '\v')) end--;   // rstrip()
    if (end == 0) return SS_NODE_LIST_UNUSUAL;                        // int('') raises in the reference
    out.reserve(end / 6 + 1);
    size_t i = 0;
    while (i < end) {
        uint64_t v = 0;
        size_t digits = 0;
        while (i < end && buf[i] >= '0' && buf[i] <= '9') { v = v * 10u + (uint64_t)(buf[i] - '0'); i++; digits++; if (digits > 10) return SS_NODE_LIST_UNUSUAL; }
        if (digits == 0 || v > 0xFFFFFFFFull) return SS_NODE_LIST_UNUSUAL;      // sign, other characters, two spaces, too large
        out.push_back((uint32_t)v);
        if (i < end) { if (buf[i] != ' ') return SS_NODE_LIST_UNUSUAL; i++; if (i == end) return SS_NODE_LIST_UNUSUAL; }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return SS_NODE_LIST_OK;
}

// paths[n] -> ptr[n + 1] (offsets into out), out[<= out_cap] (every list ascending, duplicates removed), status[n]
// (SS_NODE_LIST_*; lists with a status other than OK are empty).  out_cap: half the files' total size is always enough.
extern "C" int ss_node_lists_parse(const char *const *paths, uint32_t n, int n_threads, uint64_t *ptr, uint32_t *out,
                                   uint64_t out_cap, uint8_t *status) {
    if ((n && (!paths || !status)) || !ptr || (!out && out_cap)) return fail(SS_ERR_ARG, "ss_node_lists_parse: NULL argument");
    if (n_threads < 1) {
        cpu_set_t set;
        n_threads = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        n_threads = std::max(1, std::min(n_threads, 32));
    }
    std::vector<std::vector<uint32_t>> lists(n);
    {
        std::atomic<uint32_t> next{0};
        auto work = [&]() { for (uint32_t i; (i = next.fetch_add(1)) < n;) status[i] = (uint8_t)parse_node_list(paths[i], lists[i]); };
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    ptr[0] = 0;
    for (uint32_t i = 0; i < n; i++) ptr[i + 1] = ptr[i] + lists[i].size();
    if (ptr[n] > out_cap) return fail(SS_ERR_ARG, "ss_node_lists_parse: output buffer too small");
    {
        std::atomic<uint32_t> next{0};
        auto work = [&]() { for (uint32_t i; (i = next.fetch_add(1)) < n;) if (!lists[i].empty()) memcpy(out + ptr[i], lists[i].data(), lists[i].size() * sizeof(uint32_t)); };
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    return SS_OK;
}

// host-only: the two decode-table forms against each other on random prefix codes
extern "C" int ss_dgz_tables_selftest_host(uint64_t seed, uint32_t trials, uint64_t *n_checked) {
    std::string err;
    if (ss_dgz_tables_selftest(seed, trials, n_checked, err)) return fail(SS_ERR_FORMAT, "decode-table self-test: " + err);
    return SS_OK;
}

// host-only: how a gzip file (range) of `compressed_bytes` would be cut up for the device inflate
extern "C" int ss_dgz_plan_host(size_t compressed_bytes, int n_sm, double head_ratio, uint64_t *out6) {
    if (!out6 || n_sm < 1) return fail(SS_ERR_ARG, "ss_dgz_plan_host: bad argument");
    const ss_dgz_shape sh = ss_dgz::shape_from_env();
    const ss_dgz_plan pl = ss_dgz_make_plan(compressed_bytes, n_sm, sh, head_ratio);
    out6[0] = pl.piece; out6[1] = pl.max_pieces; out6[2] = pl.expand; out6[3] = pl.scratch; out6[4] = pl.device_bytes;
    out6[5] = (uint64_t)n_sm * ss_dgz::decoders_per_sm(sh);
    return SS_OK;
}

extern "C" int ss_reads_from_device(ss_ctx *c, void *dev_ptr, size_t len, size_t capacity, ss_reads **out) {
    if (!c || !out || (!dev_ptr && len)) return fail(SS_ERR_ARG, "ss_reads_from_device: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    *out = nullptr;
    if (((uintptr_t)dev_ptr & 255u) != 0) return fail(SS_ERR_ARG, "ss_reads_from_device: pointer must be 256-byte aligned");
    SS_CUDA(cudaSetDevice(c->device));
    ss_reads *r = new ss_reads();
    r->ctx = c; r->len = len;
    r->seg.emplace_back();
    ss_segment &g = r->seg.back();
    g.d_text = (uint8_t *)dev_ptr; g.len = len; g.cap = capacity; g.owns_text = false;
    int rc = finish_segment(c, g);
    if (rc) { ss_reads_free(r); return rc; }
    *out = r;
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// count drivers
// ---------------------------------------------------------------------------------------------
static int check_format_result(ss_ctx *c, const char *what) {
    if (c->h_stats[4] != ~0ull) {
        char buf[256];
        snprintf(buf, sizeof buf,
                 "%s: FASTQ record framing breaks at text byte %llu of a chunk (a file must be 4-line FASTQ throughout, "
                 "or FASTA / multi-line FASTQ from its first record on)",
                 what, c->h_stats[4]);
        return fail(SS_ERR_FORMAT, buf);
    }
    return SS_OK;
}

static int reset_pass(ss_ctx *c, const ss_kmerset *s) {
    if (!s->counters_clean)     // a pass that failed half-way; normally K3b has cleared what the last pass counted
        SS_CUDA(cudaMemsetAsync(s->d_slot_cnt, 0, (4 * s->n_buckets + 1) * sizeof(uint32_t), c->stream));
    s->counters_clean = false;
    SS_CUDA(cudaMemsetAsync(c->d_stats, 0, 4 * sizeof(unsigned long long), c->stream));
    SS_CUDA(cudaMemsetAsync(c->d_stats + 4, 0xFF, sizeof(unsigned long long), c->stream));
    SS_CUDA(cudaMemsetAsync(c->d_stats + 5, 0, sizeof(unsigned long long), c->stream));
    return SS_OK;
}

static void fill_stats(ss_ctx *c, ss_stats *st) {
    st->n_kmers = c->h_stats[0]; st->n_hits = c->h_stats[1];
    st->n_second_probe = c->h_stats[2]; st->n_reads = c->h_stats[3];
    st->n_table_probes = c->h_stats[5];   // 0 when the set has no filter: every k-mer probes the table
}

// ---- binned probing ---------------------------------------------------------------------------
// A probe of a table that does not fit in L2 costs one random 128-byte DRAM line, which
// bounds a pass at ~4e10 table probes/s.  The first pass over the sample barely notices (its filter sends
// 2-3 % of the k-mers to the table), but a cluster's own k-mer set is hit by most k-mers of that cluster's
// reads.  For those passes the k-mers that pass the filter are first grouped by table range (bins, written
// as warp-owned 1 KiB chunks) and then probed bin after bin with the range resident in L2.
// Knobs (environment): SS_BIN=0 never, 1 decide from a sample (default), 2 always when a filter exists;
// SS_BIN_MIN_RATE (0.08) table probes per k-mer in the sample; SS_BIN_MIN_TABLE_MB (160); SS_BIN_SLICE_MB (3/4 of
// the L2) table + counter bytes per bin; SS_BIN_POOL_MB (4096) key pool; SS_BIN_FILTER (1).
static double env_double(const char *name, double dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atof(e) : dflt;
}

static int ensure_bin_pool(ss_ctx *c) {
    uint64_t want = (uint64_t)(env_double("SS_BIN_POOL_MB", 4096.0) * 1048576.0);
    if (c->d_bin_keys && c->bin_pool_want == want) return SS_OK;
    cudaFree(c->d_bin_keys); cudaFree(c->d_bin_fill); cudaFree(c->d_bin_n); cudaFreeHost(c->h_bin);
    c->d_bin_keys = nullptr; c->d_bin_fill = nullptr; c->d_bin_n = nullptr; c->h_bin = nullptr;
    c->bin_pool_want = want;
    size_t free_b = 0, total_b = 0;
    SS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (want > free_b / 3) want = free_b / 3;
    uint64_t chunks = want / (SS_BIN_CHUNK * 8ull);
    if (chunks > 0x1800000ull) chunks = 0x1800000ull;            // key indices are 32-bit (chunk * 128 + i)
    if (chunks < SS_BIN_PMAX) return fail(SS_ERR_NOMEM, "binned probing: no room for the key pool");
    SS_CUDA(cudaMalloc(&c->d_bin_keys, chunks * SS_BIN_CHUNK * 8ull));
    SS_CUDA(cudaMalloc(&c->d_bin_fill, chunks * sizeof(uint32_t)));
    SS_CUDA(cudaMalloc(&c->d_bin_n, 2 * SS_BIN_PMAX * sizeof(uint32_t)));
    SS_CUDA(cudaMallocHost(&c->h_bin, 6 * sizeof(unsigned long long) + 2 * SS_BIN_PMAX * sizeof(uint32_t)));
    c->bin_pool_chunks = chunks;
    return SS_OK;
}

// One round of the binned mode over tiles [lo, hi) of a segment; on a bin overflow (all k-mers alike, e.g.
// a poly-G tail that is in the set) the round is redone with direct probes -- nothing was counted yet.
static int binned_round(ss_ctx *c, const ss_kmerset *s, const ss_segment &g, uint32_t lo, uint32_t hi,
                        const ss_bin_view &bv, bool use_filter, unsigned long long *acc, uint32_t *n_binned) {
    const bool dbg = getenv("SS_DEBUG_TIMING") != nullptr;
    static cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};
    if (dbg) {
        if (!ev_t[0]) for (int i = 0; i < 3; i++) cudaEventCreate(&ev_t[i]);
        cudaEventRecord(ev_t[0], c->stream);
    }
    uint32_t *h_n = reinterpret_cast<uint32_t *>(c->h_bin + 6);
    SS_CUDA(cudaMemsetAsync(c->d_bin_n, 0, 2 * SS_BIN_PMAX * sizeof(uint32_t), c->stream));
    SS_CUDA(cudaMemsetAsync(c->d_stats + 8, 0, 7 * sizeof(unsigned long long), c->stream));
    SS_CUDA(ss_launch_bin_scan(g.d_text, g.len, lo, hi, g.d_tile_line, s->view(), bv, use_filter, c->d_stats + 8,
                               c->d_stats + 4, c->n_sm, c->stream));
    if (dbg) cudaEventRecord(ev_t[1], c->stream);
    SS_CUDA(cudaMemcpyAsync(c->h_bin, c->d_stats + 8, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaMemcpyAsync(h_n, c->d_bin_n, 2 * SS_BIN_PMAX * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    if (h_n[SS_BIN_PMAX]) {
        SS_CUDA(ss_launch_probe_range(g.d_text, g.len, lo, hi, g.d_tile_line, s->view(), c->d_stats, c->d_stats + 4,
                                      c->n_sm, c->stream));
        return SS_OK;
    }
    acc[0] += c->h_bin[0]; acc[3] += c->h_bin[3]; acc[5] += c->h_bin[5];
    SS_CUDA(ss_launch_bin_probe(bv, h_n, s->view(), c->d_stats, c->d_stats + 8, c->n_sm, c->stream));
    if (dbg) {
        cudaEventRecord(ev_t[2], c->stream);
        cudaEventSynchronize(ev_t[2]);
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ev_t[0], ev_t[1]);
        cudaEventElapsedTime(&b, ev_t[1], ev_t[2]);
        uint64_t chunks = 0;
        for (uint32_t p = 0; p < bv.P; p++) chunks += h_n[p];
        fprintf(stderr, "[ss] binned round tiles %u..%u: scan %.3f ms, probe (+sync) %.3f ms, %llu chunks, %u bins, filter %d\n",
                lo, hi, a, b, (unsigned long long)chunks, bv.P, (int)use_filter);
    }
    (*n_binned)++;
    return SS_OK;
}

extern "C" int ss_count_device(ss_ctx *c, const ss_kmerset *s, const ss_reads *r, uint32_t *dev_counts, ss_stats *st) {
    if (!c || !s || !r || !dev_counts) return fail(SS_ERR_ARG, "ss_count_device: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    if (s->ctx != c || r->ctx != c) return fail(SS_ERR_ARG, "ss_count_device: handle belongs to another context");
    double t0 = now_ms();
    SS_CUDA(cudaSetDevice(c->device));
    int rc = reset_pass(c, s);
    if (rc) return rc;
    SS_CUDA(cudaEventRecord(c->ev_a, c->stream));
    uint32_t launches = 0, n_binned = 0, index_launches = 0;
    rc = ensure_index(c, r, &index_launches);           // first pass over these reads: K1 (kept for later passes)
    if (rc) return rc;
    SS_CUDA(cudaEventRecord(c->ev_i, c->stream));
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};     // scan statistics of the binned rounds (summed on the host)

    // ---- direct or binned?  Decided from the first tiles of the pass.
    const int bin_mode = (int)env_double("SS_BIN", 1.0);
    const uint64_t table_bytes = s->n_buckets * (sizeof(ss_bucket) + 4 * sizeof(uint32_t));
    uint64_t total_tiles = 0;
    for (const ss_segment &g : r->seg) total_tiles += g.n_tiles;
    const bool force = bin_mode == 2;                   // tests: any size, SS_BIN_ROUND_TILES tiles per round
    const uint32_t sample_tiles = force ? 64 : 16384;   // 16 MB of text
    ss_bin_view bv = {};
    uint32_t round_tiles = 0;
    bool use_filter = true;
    size_t first_seg = 0;
    uint32_t first_lo = 0;
    if (bin_mode > 0 && s->d_filter && total_tiles > 0 &&
        (force || (total_tiles >= 8ull * sample_tiles &&
                   table_bytes >= (uint64_t)(env_double("SS_BIN_MIN_TABLE_MB", 160.0) * 1048576.0)))) {
        while (r->seg[first_seg].n_tiles == 0) first_seg++;
        const ss_segment &g = r->seg[first_seg];
        double rate, probes_per_tile, kmers_per_tile;
        if (!force && s->seen_reads_id == r->id) {        // same set, same reads: the sample was taken before
            rate = s->seen_rate; probes_per_tile = s->seen_per_tile; kmers_per_tile = s->seen_kmers_per_tile;
        } else {
            const uint32_t n_sample = g.n_tiles < sample_tiles ? g.n_tiles : sample_tiles;
            SS_CUDA(ss_launch_probe_range(g.d_text, g.len, 0, n_sample, g.d_tile_line, s->view(), c->d_stats, c->d_stats + 4,
                                          c->n_sm, c->stream));
            SS_CUDA(cudaMemcpyAsync(c->h_stats, c->d_stats, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            SS_CUDA(cudaEventRecord(c->ev_d, c->stream));
            // the GPU goes on with the next tiles (direct) while the host waits for the sample's numbers
            first_lo = force ? n_sample : (g.n_tiles < 4 * sample_tiles ? g.n_tiles : 4 * sample_tiles);
            SS_CUDA(ss_launch_probe_range(g.d_text, g.len, n_sample, first_lo, g.d_tile_line, s->view(), c->d_stats,
                                          c->d_stats + 4, c->n_sm, c->stream));
            launches += first_lo > n_sample ? 2 : 1;
            SS_CUDA(cudaEventSynchronize(c->ev_d));
            rate = c->h_stats[0] ? (double)c->h_stats[5] / (double)c->h_stats[0] : 0.0;
            probes_per_tile = (double)c->h_stats[5] / (double)n_sample;
            kmers_per_tile = (double)c->h_stats[0] / (double)n_sample;
            s->seen_reads_id = r->id; s->seen_rate = rate; s->seen_per_tile = probes_per_tile; s->seen_kmers_per_tile = kmers_per_tile;
        }
        // SS_BIN_FILTER=0: bin every k-mer without asking the filter (one L2 load less per k-mer, more keys to bin;
        // measured equal at 80 % hits, so the filter stays on by default)
        use_filter = (int)env_double("SS_BIN_FILTER", 1.0) != 0;
        const double per_tile = use_filter ? probes_per_tile : kmers_per_tile;               // keys to bin per tile
        if ((force || rate >= env_double("SS_BIN_MIN_RATE", 0.08)) && ensure_bin_pool(c) == SS_OK) {
            // table + counter bytes per bin: 3/4 of the L2 (measured on B200, 545 MB table at 80 % hits: 8 bins of
            // 89 MB 20.4 ms, 16 bins 21.4 ms, 32 bins 26.4 ms -- every (warp, bin) pair is an open write stream of the
            // scan, and fewer streams cost the scan less than the larger range costs the probes)
            const uint64_t slice = (uint64_t)(env_double("SS_BIN_SLICE_MB", 0.75 * c->prop.l2CacheSize / 1048576.0) * 1048576.0);
            uint32_t P = 2;
            while (P < SS_BIN_PMAX && table_bytes / P > slice) P *= 2;
            bv.keys = c->d_bin_keys; bv.fill = c->d_bin_fill; bv.n_chunks = c->d_bin_n;
            bv.P = P; bv.cap = (uint32_t)(c->bin_pool_chunks / P);
            bv.shift = 32; for (uint32_t q = P; q > 1; q >>= 1) bv.shift--;
            // every warp of the scan kernel may hold one partly filled chunk per bin
            const uint64_t warps = (uint64_t)c->n_sm * ss_bin_scan_ctas_per_sm() * SS_CTA_WARPS;
            if (force) {
                round_tiles = (uint32_t)env_double("SS_BIN_ROUND_TILES", 1024.0);
                if (round_tiles < 1) round_tiles = 1;
            } else if (bv.cap > warps + 64) {
                const double usable = (double)(bv.cap - warps) * SS_BIN_CHUNK * P / 1.3;   // margin: sample error + bin skew
                double rt = usable / (per_tile > 1.0 ? per_tile : 1.0);
                round_tiles = rt > 4e9 ? 0xFFFFFFF0u : (uint32_t)rt;
                if (round_tiles < 4 * sample_tiles) round_tiles = 0;                       // pool too small to pay off
            }
        }
    }

    for (size_t gi = 0; gi < r->seg.size(); gi++) {
        const ss_segment &g = r->seg[gi];
        if (!g.n_tiles) continue;
        uint32_t lo = (first_lo && gi == first_seg) ? first_lo : 0;
        if (!round_tiles) {
            SS_CUDA(ss_launch_probe_range(g.d_text, g.len, lo, g.n_tiles, g.d_tile_line, s->view(), c->d_stats,
                                          c->d_stats + 4, c->n_sm, c->stream));
            launches++;
            continue;
        }
        // equal rounds, so that the last one is not a sliver
        const uint32_t n_rounds = (g.n_tiles - lo + round_tiles - 1) / round_tiles;
        const uint32_t step = n_rounds ? (g.n_tiles - lo + n_rounds - 1) / n_rounds : 0;
        while (lo < g.n_tiles) {
            uint32_t hi = g.n_tiles - lo > step ? lo + step : g.n_tiles;
            rc = binned_round(c, s, g, lo, hi, bv, use_filter, acc, &n_binned);
            if (rc) return rc;
            launches += 2;
            lo = hi;
        }
    }
    SS_CUDA(cudaEventRecord(c->ev_b, c->stream));
    SS_CUDA(ss_launch_scatter(s->d_slot_cnt, s->d_ord_of_slot, s->n_buckets, s->d_dup, s->n_dup, s->d_slot_of, s->n_records,
                              dev_counts, c->n_sm, c->stream));
    SS_CUDA(cudaEventRecord(c->ev_c, c->stream));
    SS_CUDA(cudaMemcpyAsync(c->h_stats, c->d_stats, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    s->counters_clean = true;
    c->h_stats[0] += acc[0]; c->h_stats[3] += acc[3]; c->h_stats[5] += acc[5];
    rc = check_format_result(c, "ss_count");
    if (rc) return rc;
    if (st) {
        memset(st, 0, sizeof *st);
        fill_stats(c, st);
        st->text_bytes = r->len;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_a, c->ev_i); st->ms_index = index_launches ? ms : 0.0;
        cudaEventElapsedTime(&ms, c->ev_i, c->ev_b); st->ms_probe = ms;
        cudaEventElapsedTime(&ms, c->ev_b, c->ev_c); st->ms_gather = ms;
        st->probe_launches = launches;
        st->total_launches = st->probe_launches + index_launches + (s->n_records ? 1 + (s->n_dup ? 1 : 0) : 0);
        st->binned_rounds = n_binned;
        st->bins = n_binned ? bv.P : 0;
        st->ms_total = now_ms() - t0;
    }
    return SS_OK;
}

extern "C" int ss_count(ss_ctx *c, const ss_kmerset *s, const ss_reads *r, uint32_t *counts, ss_stats *st) {
    if (!c || !s || !r || !counts) return fail(SS_ERR_ARG, "ss_count: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    double t0 = now_ms();
    SS_CUDA(cudaSetDevice(c->device));
    int rc = ensure_dense(c, s->n_records);
    if (rc) return rc;
    rc = ss_count_device(c, s, r, c->d_dense, st);
    if (rc) return rc;
    if (s->n_records) SS_CUDA(cudaMemcpy(counts, c->d_dense, s->n_records * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (st) st->ms_total = now_ms() - t0;
    return SS_OK;
}

static int ensure_chunks(ss_ctx *c) {
    if (c->d_chunk[0]) return SS_OK;
    const size_t text_cap = std::max(c->chunk_bytes, c->device_bgzf ? c->bgzf_out_cap + SSI_OUT_SLACK : 0);
    size_t cap = ss_reads_device_capacity(text_cap);
    uint32_t tiles = (uint32_t)(text_cap / SS_TILE) + 2;
    for (int i = 0; i < SS_NSLOT; i++) {
        SS_CUDA(cudaMalloc(&c->d_chunk[i], cap));
        SS_CUDA(cudaMalloc(&c->d_chunk_line[i], ss_index_words(tiles) * sizeof(uint32_t)));
    }
    return SS_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// the double-buffered device side of the streaming drivers
struct stream_state { int slot = 0; uint32_t probe_launches = 0, total_launches = 0; uint64_t bytes = 0; bool used[SS_NSLOT] = {}; };

// copy one record-aligned chunk (n bytes at src, host) into device slot ss.slot and scan it; a BGZF
// batch (`batch` != NULL) is inflated into the slot by the device instead
static int stream_chunk(ss_ctx *c, const ss_kmerset *s, const char *src, size_t n, stream_state &ss,
                        ss_chunk *batch = nullptr) {
    int b = ss.slot;
    if (ss.used[b]) SS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));   // kernels done with it
    if (batch) {
        SS_CUDA(cudaMemsetAsync(c->d_chunk[b] + n, '\n', ss_reads_device_capacity(n) - n, c->copy_stream));
        cudaEvent_t inflated;
        int rc = inflate_batch(c, batch, c->d_chunk[b], &inflated);
        if (rc) return rc;
        SS_CUDA(cudaStreamWaitEvent(c->stream, inflated, 0));
        ss.total_launches++;
    } else {
        SS_CUDA(cudaMemcpyAsync(c->d_chunk[b], src, n, cudaMemcpyHostToDevice, c->copy_stream));
        SS_CUDA(cudaMemsetAsync(c->d_chunk[b] + n, '\n', ss_reads_device_capacity(n) - n, c->copy_stream));
        SS_CUDA(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        SS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    }
    uint32_t tiles = (uint32_t)((n + SS_TILE - 1) / SS_TILE);
    SS_CUDA(ss_launch_index(c->d_chunk[b], tiles, c->d_chunk_line[b], 0, c->n_sm, c->stream));
    SS_CUDA(ss_launch_probe(c->d_chunk[b], n, tiles, c->d_chunk_line[b], s->view(), c->d_stats, c->d_stats + 4,
                            c->n_sm, c->stream));
    SS_CUDA(cudaEventRecord(c->ev_done[b], c->stream));
    ss.used[b] = true;
    ss.probe_launches++; ss.total_launches += 2; ss.bytes += n;
    ss.slot = (b + 1) % SS_NSLOT;
    return SS_OK;
}

// stream `len` bytes of whole FASTQ records (caller's host memory) through the chunk pipeline
static int stream_text(ss_ctx *c, const ss_kmerset *s, const char *buf, size_t len, stream_state &ss) {
    size_t pos = 0;
    const bool pinned = len ? is_pinned(buf) : true;
    const size_t chunk = c->chunk_bytes;
    while (pos < len) {
        size_t end = pos + chunk >= len ? len : ss_find_cut(buf, pos, pos + chunk);
        if (end > pos + chunk || end <= pos)
            return fail(SS_ERR_FORMAT, "reads: no FASTQ record boundary within a chunk (a record longer than SS_CHUNK_BYTES?)");
        size_t n = end - pos;
        int b = ss.slot;
        const char *src = buf + pos;
        if (!pinned) {   // stage pageable memory through our own pinned buffer so the copy stays asynchronous
            if (!c->h_pinned[b]) SS_CUDA(cudaMallocHost(&c->h_pinned[b], chunk));
            if (ss.used[b]) SS_CUDA(cudaEventSynchronize(c->ev_copied[b]));
            memcpy(c->h_pinned[b], src, n);
            src = (const char *)c->h_pinned[b];
        }
        int rc = stream_chunk(c, s, src, n, ss);
        if (rc) return rc;
        pos = end;
    }
    return SS_OK;
}

// common tail of the streaming drivers: gather, copy out, statistics
static int finish_streamed(ss_ctx *c, const ss_kmerset *s, stream_state &ss, uint32_t *counts, ss_stats *st, double t0,
                           const char *what) {
    SS_CUDA(cudaEventRecord(c->ev_b, c->stream));
    SS_CUDA(ss_launch_scatter(s->d_slot_cnt, s->d_ord_of_slot, s->n_buckets, s->d_dup, s->n_dup, s->d_slot_of, s->n_records,
                              c->d_dense, c->n_sm, c->stream));
    SS_CUDA(cudaEventRecord(c->ev_c, c->stream));
    if (s->n_records)
        SS_CUDA(cudaMemcpyAsync(counts, c->d_dense, s->n_records * sizeof(uint32_t), cudaMemcpyDefault, c->stream));
    SS_CUDA(cudaMemcpyAsync(c->h_stats, c->d_stats, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    s->counters_clean = true;
    int rc = check_format_result(c, what);
    if (rc) return rc;
    if (st) {
        memset(st, 0, sizeof *st);
        fill_stats(c, st);
        st->text_bytes = ss.bytes;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); st->ms_probe = ms;   // copies + index + probe, overlapped
        cudaEventElapsedTime(&ms, c->ev_b, c->ev_c); st->ms_gather = ms;
        st->probe_launches = ss.probe_launches;
        st->total_launches = ss.total_launches + (s->n_records ? 1 + (s->n_dup ? 1 : 0) : 0);
        st->ms_total = now_ms() - t0;
    }
    return SS_OK;
}

static int count_streamed(ss_ctx *c, const ss_kmerset *s, const char *const *bufs, const size_t *lens, int n,
                          uint32_t *counts, ss_stats *st) {
    double t0 = now_ms();
    SS_CUDA(cudaSetDevice(c->device));
    int rc = ensure_chunks(c);
    if (rc) return rc;
    rc = ensure_dense(c, s->n_records);
    if (rc) return rc;
    rc = reset_pass(c, s);
    if (rc) return rc;
    SS_CUDA(cudaEventRecord(c->ev_a, c->stream));
    stream_state ss;
    std::vector<std::vector<char>> stores((size_t)n);
    for (int i = 0; i < n; i++) {
        const char *b = bufs[i];
        size_t bl = lens[i];
        rc = normalize_host_text(b, bl, stores[(size_t)i], "reads");   // FASTA / wrapped FASTQ: rewritten copy
        if (rc) return rc;
        size_t l = trim_tail(b, bl);
        rc = check_fastq_head(b, l, "reads");
        if (rc) return rc;
        rc = stream_text(c, s, b, l, ss);   // every buffer starts a record: line index restarts at 0
        if (rc) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream); return rc; }
    }
    return finish_streamed(c, s, ss, counts, st, t0, "ss_count_host");
}

extern "C" int ss_count_host(ss_ctx *c, const ss_kmerset *s, const char *const *bufs, const size_t *lens, int n,
                             uint32_t *counts, ss_stats *st) {
    if (!c || !s || !counts || (n > 0 && (!bufs || !lens))) return fail(SS_ERR_ARG, "ss_count_host: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    if (s->ctx != c) return fail(SS_ERR_ARG, "ss_count_host: set belongs to another context");
    return count_streamed(c, s, bufs, lens, n, counts, st);
}

// End to end from files: producer threads read / inflate into pinned chunks while earlier chunks are
// copied and scanned; nothing but two device chunk buffers is resident.
extern "C" int ss_count_files(ss_ctx *c, const ss_kmerset *s, const char *const *paths, int n_paths, int shard,
                              int n_shards, uint32_t *counts, ss_stats *st) {
    if (!c || !s || !counts || (n_paths > 0 && !paths)) return fail(SS_ERR_ARG, "ss_count_files: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    if (s->ctx != c) return fail(SS_ERR_ARG, "ss_count_files: set belongs to another context");
    double t0 = now_ms();
    SS_CUDA(cudaSetDevice(c->device));
    int rc = ensure_chunks(c);
    if (rc) return rc;
    rc = ensure_dense(c, s->n_records);
    if (rc) return rc;
    rc = ensure_source(c);
    if (rc) return rc;
    ss_text_source &src = *c->src;
    rc = reset_pass(c, s);
    if (rc) return rc;
    SS_CUDA(cudaEventRecord(c->ev_a, c->stream));
    stream_state ss;
    // ordinary gzip files big enough to pay off are inflated on the device, batch after batch, each batch scanned
    // where it was inflated (ss_dgz.cu)
    std::vector<const char *> rest;
    {
        std::vector<dgz_file> dgz_files;
        for (int i = 0; i < n_paths; i++) {
            dgz_file df;
            if (dgz_eligible(c, paths[i], shard, n_shards, df)) dgz_files.push_back(df);
            else rest.push_back(paths[i]);
        }
        rc = dgz_run_files(c, dgz_files, [&](const uint8_t *d_text, size_t n) {
            const uint32_t tiles = (uint32_t)((n + SS_TILE - 1) / SS_TILE);
            if (ss_index_words(tiles) > c->dgz_line_words) {
                cudaFree(c->d_dgz_line); c->d_dgz_line = nullptr; c->dgz_line_words = 0;
                const size_t words = ss_index_words(tiles) + (1u << 16);
                SS_CUDA(cudaMalloc(&c->d_dgz_line, words * sizeof(uint32_t)));
                c->dgz_line_words = words;
            }
            SS_CUDA(cudaMemsetAsync((uint8_t *)d_text + n, '\n', ss_reads_device_capacity(n) - n, c->stream));
            SS_CUDA(ss_launch_index(d_text, tiles, c->d_dgz_line, 0, c->n_sm, c->stream));
            SS_CUDA(ss_launch_probe(d_text, n, tiles, c->d_dgz_line, s->view(), c->d_stats, c->d_stats + 4, c->n_sm, c->stream));
            SS_CUDA(cudaStreamSynchronize(c->stream));
            ss.probe_launches++; ss.total_launches += 2; ss.bytes += n;
            return (int)SS_OK;
        });
        if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    }
    paths = rest.data(); n_paths = (int)rest.size();
    rc = src.start(paths, n_paths, shard, n_shards);
    if (rc) return fail(rc, src.error());
    pending_ring pend(c);
    bool any_gz = false;
    rc = gz_begin(c);
    while (!rc) {
        ss_chunk *ch = src.next();
        if (!ch) break;
        const bool gz = ch->kind == SS_CHUNK_BGZF;
        any_gz |= gz;
        rc = stream_chunk(c, s, (const char *)ch->text, ch->text_len(), ss, gz ? ch : nullptr);
        if (rc) { src.release(ch); break; }
        cudaError_t e = pend.push(ch, c->copy_stream);
        if (e != cudaSuccess) { rc = ss_cuda_fail(e, "ingest event", __FILE__, __LINE__); break; }
    }
    pend.drain();
    int src_rc = src.finish();
    if (rc || src_rc) {
        cudaStreamSynchronize(c->copy_stream); gz_sync(c); cudaStreamSynchronize(c->stream);
        return rc ? rc : fail(src_rc, src.error());
    }
    rc = finish_streamed(c, s, ss, counts, st, t0, "ss_count_files");
    if (any_gz) { int grc = gz_check(c, "ss_count_files"); if (grc) return grc; }   // a bad block explains a framing error
    return rc;
}

extern "C" int ss_l2_finalize(ss_ctx *c, const ss_kmerset *s, const uint32_t *dev_counts, int64_t *py_o) {
    if (!c || !s || !dev_counts || !py_o) return fail(SS_ERR_ARG, "ss_l2_finalize: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    SS_CUDA(cudaSetDevice(c->device));
    uint64_t n = s->n_records;
    if (!n) return SS_OK;
    long long *d = nullptr;
    SS_CUDA(cudaMalloc(&d, n * sizeof(long long)));
    cudaError_t e = ss_launch_l2_finalize(dev_counts, s->d_flags, s->d_row_of, n, d, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(py_o, d, n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return ss_cuda_fail(e, "ss_l2_finalize", __FILE__, __LINE__);
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// reducers
// ---------------------------------------------------------------------------------------------
template <typename T>
struct dev_buf {
    T *p = nullptr;
    ~dev_buf() { cudaFree(p); }
    cudaError_t alloc(uint64_t n) { return cudaMalloc(&p, std::max<uint64_t>(n, 1) * sizeof(T)); }
};

// Persistent device forms of the two index structures the reducers walk: the CSR node -> record ordinals of
// Tree_database/kmers/<node> (K4) and the CSC strain -> rows of all_strains_re.npz (K5).  Uploaded and validated
// once; a reduction then moves only its inputs (y, mask) and its few outputs.
struct ss_node_index {
    ss_ctx *ctx = nullptr;
    uint32_t n_nodes = 0;
    uint64_t nnz = 0;
    unsigned long long *d_ptr = nullptr, *d_sum = nullptr;
    uint32_t *d_ord = nullptr, *d_len = nullptr, *d_cov = nullptr, *d_max = nullptr;
};

struct ss_strain_matrix {
    ss_ctx *ctx = nullptr;
    uint32_t n_strains = 0;
    uint64_t n_rows = 0, nnz = 0;
    unsigned long long *d_ptr = nullptr, *d_out = nullptr;      // d_out: total | covered | sum
    uint32_t *d_rows = nullptr;
    long long *d_y = nullptr;
    uint8_t *d_mask = nullptr;
};

extern "C" int ss_node_index_free(ss_node_index *x) {
    if (!x) return SS_OK;
    cudaSetDevice(x->ctx->device);
    cudaFree(x->d_ptr); cudaFree(x->d_sum); cudaFree(x->d_ord); cudaFree(x->d_len); cudaFree(x->d_cov); cudaFree(x->d_max);
    delete x;
    return SS_OK;
}

extern "C" int ss_node_index_create(ss_ctx *c, const uint64_t *node_ptr, const uint32_t *ordinals, uint32_t n_nodes,
                                    ss_node_index **out) {
    if (!c || !node_ptr || !out) return fail(SS_ERR_ARG, "ss_node_index_create: NULL argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_nodes; i++)
        if (node_ptr[i + 1] < node_ptr[i]) return fail(SS_ERR_ARG, "ss_node_index_create: node_ptr is not non-decreasing");
    const uint64_t nnz = node_ptr[n_nodes] - node_ptr[0];
    if (nnz && !ordinals) return fail(SS_ERR_ARG, "ss_node_index_create: ordinals is NULL");
    SS_CUDA(cudaSetDevice(c->device));
    ss_node_index *x = new ss_node_index();
    x->ctx = c; x->n_nodes = n_nodes; x->nnz = nnz;
    const uint64_t nn = std::max<uint32_t>(n_nodes, 1);
    cudaError_t e = cudaMalloc(&x->d_ptr, (nn + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_sum, nn * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_ord, std::max<uint64_t>(nnz, 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_len, nn * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_cov, nn * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_max, nn * sizeof(uint32_t));
    if (e == cudaSuccess) {
        std::vector<unsigned long long> rel(n_nodes + 1);           // offsets relative to ordinals[node_ptr[0]]
        for (uint32_t i = 0; i <= n_nodes; i++) rel[i] = node_ptr[i] - node_ptr[0];
        e = cudaMemcpy(x->d_ptr, rel.data(), (n_nodes + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && nnz) e = cudaMemcpy(x->d_ord, ordinals + node_ptr[0], nnz * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ss_node_index_free(x); return ss_cuda_fail(e, "ss_node_index_create", __FILE__, __LINE__); }
    *out = x;
    return SS_OK;
}

extern "C" int ss_node_index_reduce(ss_ctx *c, const ss_kmerset *s, const ss_node_index *x, const uint32_t *dev_counts,
                                    uint32_t *length, uint32_t *covered, uint64_t *sum, uint32_t *max_count) {
    if (!c || !s || !x || !dev_counts || !length || !covered || !sum) return fail(SS_ERR_ARG, "ss_node_index_reduce: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    if (x->ctx != c || s->ctx != c) return fail(SS_ERR_ARG, "ss_node_index_reduce: handle belongs to another context");
    SS_CUDA(cudaSetDevice(c->device));
    const uint32_t n = x->n_nodes;
    if (n == 0) return SS_OK;
    SS_CUDA(ss_launch_node_reduce(dev_counts, s->d_flags, x->d_ptr, x->d_ord, n, s->n_records, x->d_len, x->d_cov, x->d_sum,
                                  x->d_max, c->stream));
    SS_CUDA(cudaMemcpyAsync(length, x->d_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaMemcpyAsync(covered, x->d_cov, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaMemcpyAsync(sum, x->d_sum, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (max_count) SS_CUDA(cudaMemcpyAsync(max_count, x->d_max, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}

extern "C" int ss_node_reduce(ss_ctx *c, const ss_kmerset *s, const uint32_t *dev_counts, const uint64_t *node_ptr,
                              const uint32_t *ordinals, uint32_t n_nodes, uint32_t *length, uint32_t *covered,
                              uint64_t *sum) {
    if (!c || !s || !dev_counts || !node_ptr || !length || !covered || !sum)
        return fail(SS_ERR_ARG, "ss_node_reduce: NULL argument");
    if (n_nodes == 0) return SS_OK;
    ss_node_index *x = nullptr;
    int rc = ss_node_index_create(c, node_ptr, ordinals, n_nodes, &x);
    if (rc) return rc;
    rc = ss_node_index_reduce(c, s, x, dev_counts, length, covered, sum, nullptr);
    ss_node_index_free(x);
    return rc;
}

extern "C" int ss_strain_matrix_free(ss_strain_matrix *m) {
    if (!m) return SS_OK;
    cudaSetDevice(m->ctx->device);
    cudaFree(m->d_ptr); cudaFree(m->d_out); cudaFree(m->d_rows); cudaFree(m->d_y); cudaFree(m->d_mask);
    delete m;
    return SS_OK;
}

extern "C" int ss_strain_matrix_create(ss_ctx *c, const uint64_t *col_ptr, const uint32_t *rows, uint32_t n_strains,
                                       uint64_t n_rows, ss_strain_matrix **out) {
    if (!c || !col_ptr || !out) return fail(SS_ERR_ARG, "ss_strain_matrix_create: NULL argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_strains; i++)
        if (col_ptr[i + 1] < col_ptr[i]) return fail(SS_ERR_ARG, "ss_strain_matrix_create: col_ptr is not non-decreasing");
    const uint64_t nnz = col_ptr[n_strains] - col_ptr[0];
    if (nnz && !rows) return fail(SS_ERR_ARG, "ss_strain_matrix_create: rows is NULL");
    for (uint64_t i = 0; i < nnz; i++)                            // once per matrix, not per reduction
        if (rows[col_ptr[0] + i] >= n_rows) return fail(SS_ERR_ARG, "ss_strain_matrix_create: row index out of range");
    SS_CUDA(cudaSetDevice(c->device));
    ss_strain_matrix *m = new ss_strain_matrix();
    m->ctx = c; m->n_strains = n_strains; m->n_rows = n_rows; m->nnz = nnz;
    const uint64_t ns = std::max<uint32_t>(n_strains, 1);
    cudaError_t e = cudaMalloc(&m->d_ptr, (ns + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_out, 3 * ns * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_rows, std::max<uint64_t>(nnz, 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_y, std::max<uint64_t>(n_rows, 1) * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_mask, std::max<uint64_t>(n_rows, 1));
    if (e == cudaSuccess) {
        std::vector<unsigned long long> rel(n_strains + 1);
        for (uint32_t i = 0; i <= n_strains; i++) rel[i] = col_ptr[i] - col_ptr[0];
        e = cudaMemcpy(m->d_ptr, rel.data(), (n_strains + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && nnz) e = cudaMemcpy(m->d_rows, rows + col_ptr[0], nnz * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ss_strain_matrix_free(m); return ss_cuda_fail(e, "ss_strain_matrix_create", __FILE__, __LINE__); }
    *out = m;
    return SS_OK;
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}

extern "C" int ss_strain_matrix_reduce(ss_ctx *c, ss_strain_matrix *m, const int64_t *y, const uint8_t *row_mask,
                                       uint64_t *total, uint64_t *covered, uint64_t *sum) {
    if (!c || !m || !y || !total || !covered || !sum) return fail(SS_ERR_ARG, "ss_strain_matrix_reduce: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(c->mu);
    if (m->ctx != c) return fail(SS_ERR_ARG, "ss_strain_matrix_reduce: handle belongs to another context");
    SS_CUDA(cudaSetDevice(c->device));
    const uint32_t n = m->n_strains;
    if (n == 0) return SS_OK;
    const long long *d_y = (const long long *)y;                  // y (and the mask) may already live on the device
    if (!is_device_ptr(y)) {
        if (m->n_rows) SS_CUDA(cudaMemcpyAsync(m->d_y, y, m->n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        d_y = m->d_y;
    }
    const uint8_t *d_mask = row_mask;
    if (row_mask && !is_device_ptr(row_mask)) {
        if (m->n_rows) SS_CUDA(cudaMemcpyAsync(m->d_mask, row_mask, m->n_rows, cudaMemcpyHostToDevice, c->stream));
        d_mask = m->d_mask;
    }
    SS_CUDA(ss_launch_strain_reduce(m->d_ptr, m->d_rows, n, d_y, d_mask, m->d_out, m->d_out + n, m->d_out + 2 * (uint64_t)n,
                                    c->stream));
    SS_CUDA(cudaMemcpyAsync(total, m->d_out, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaMemcpyAsync(covered, m->d_out + n, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaMemcpyAsync(sum, m->d_out + 2 * (uint64_t)n, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}

extern "C" int ss_strain_reduce(ss_ctx *c, const uint64_t *col_ptr, const uint32_t *rows, uint32_t n_strains,
                                const int64_t *y, const uint8_t *row_mask, uint64_t n_rows, uint64_t *total,
                                uint64_t *covered, uint64_t *sum) {
    if (!c || !col_ptr || !y || !total || !covered || !sum) return fail(SS_ERR_ARG, "ss_strain_reduce: NULL argument");
    if (n_strains == 0) return SS_OK;
    ss_strain_matrix *m = nullptr;
    int rc = ss_strain_matrix_create(c, col_ptr, rows, n_strains, n_rows, &m);
    if (rc) return rc;
    rc = ss_strain_matrix_reduce(c, m, y, row_mask, total, covered, sum);
    ss_strain_matrix_free(m);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// measurement + synthetic workload helpers
// ---------------------------------------------------------------------------------------------
extern "C" int ss_bench_random_gather(ss_ctx *c, uint64_t bytes, uint64_t n_probes, int iters, double *gbps) {
    if (!c || !gbps) return fail(SS_ERR_ARG, "ss_bench_random_gather: NULL argument");
    SS_CUDA(cudaSetDevice(c->device));
    uint64_t n_sectors = bytes / 32;
    if (n_sectors == 0 || n_probes == 0) return fail(SS_ERR_ARG, "ss_bench_random_gather: empty");
    dev_buf<uint8_t> buf;
    SS_CUDA(buf.alloc(n_sectors * 32));
    SS_CUDA(cudaMemsetAsync(buf.p, 0x5A, n_sectors * 32, c->stream));
    double best = 0;
    for (int i = 0; i < iters + 1; i++) {
        SS_CUDA(cudaEventRecord(c->ev_a, c->stream));
        SS_CUDA(ss_launch_random_gather(buf.p, n_sectors, n_probes, 0x1234 + 7919ull * i, c->d_stats + 6, c->n_sm, c->stream));
        SS_CUDA(cudaEventRecord(c->ev_b, c->stream));
        SS_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0;
        SS_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
        if (i > 0) best = std::max(best, (double)n_probes * 32.0 / (ms * 1e-3) / 1e9);
    }
    *gbps = best;
    return SS_OK;
}

extern "C" size_t ss_synth_read_record_bytes(const ss_synth_params *p) {
    return p ? (size_t)p->header_len + 1 + p->read_len + 1 + 2 + p->read_len + 1 : 0;
}
extern "C" size_t ss_synth_db_record_bytes(const ss_synth_params *p) { return p ? (size_t)p->k + 4 : 0; }

static int check_synth(const ss_synth_params *p) {
    if (!p) return fail(SS_ERR_ARG, "synth: params is NULL");
    if (p->n_leaves < 1 || p->k < 1 || p->k > 32 || p->block_len < p->k + 1 || p->genome_len < p->block_len ||
        p->read_len < 1 || p->read_len > p->genome_len || p->header_len < 3 || p->n_sources < 1 ||
        p->n_sources > SS_SYNTH_MAX_SOURCES)
        return fail(SS_ERR_ARG, "synth: bad parameters");
    for (uint32_t i = 0; i < p->n_sources; i++)
        if (p->source_leaf[i] >= p->n_leaves) return fail(SS_ERR_ARG, "synth: source_leaf out of range");
    return SS_OK;
}

extern "C" int ss_synth_reads_device(ss_ctx *c, const ss_synth_params *p, void *dev_text, uint64_t n_reads,
                                     uint64_t first_read) {
    if (!c || !dev_text) return fail(SS_ERR_ARG, "ss_synth_reads_device: NULL argument");
    int rc = check_synth(p);
    if (rc) return rc;
    SS_CUDA(cudaSetDevice(c->device));
    SS_CUDA(ss_launch_synth_reads_impl(*p, (uint8_t *)dev_text, n_reads, first_read, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}

extern "C" int ss_synth_db_host(ss_ctx *c, const ss_synth_params *p, const uint32_t *node_sizes, uint32_t n_nodes,
                                char *text_out, uint32_t *node_of_record) {
    if (!c || !node_sizes || !text_out) return fail(SS_ERR_ARG, "ss_synth_db_host: NULL argument");
    int rc = check_synth(p);
    if (rc) return rc;
    SS_CUDA(cudaSetDevice(c->device));
    ss_synth_db_layout lay;
    std::string why = ss_synth_db_layout_build(p, node_sizes, n_nodes, lay);
    if (!why.empty()) return fail(why == "too many records" ? SS_ERR_UNSUPPORTED : SS_ERR_ARG, "ss_synth_db_host: " + why);
    const std::vector<unsigned long long> &node_off = lay.node_off;
    const std::vector<uint32_t> &blk_list = lay.blk_list, &blk_off = lay.blk_off;
    const uint64_t N = lay.n_records, a = lay.perm_a;
    if (N == 0) return SS_OK;
    ss_synth_db_plan plan;
    dev_buf<unsigned long long> d_off;
    dev_buf<uint32_t> d_list, d_boff, d_node;
    dev_buf<uint8_t> d_text;
    const uint64_t rec = p->k + 4;
    SS_CUDA(d_off.alloc(n_nodes + 1)); SS_CUDA(d_list.alloc(blk_list.size())); SS_CUDA(d_boff.alloc(blk_off.size()));
    SS_CUDA(d_text.alloc(N * rec));
    if (node_of_record) SS_CUDA(d_node.alloc(N));
    SS_CUDA(cudaMemcpy(d_off.p, node_off.data(), (n_nodes + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    SS_CUDA(cudaMemcpy(d_list.p, blk_list.data(), blk_list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    SS_CUDA(cudaMemcpy(d_boff.p, blk_off.data(), blk_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    plan.node_off = d_off.p; plan.n_nodes = n_nodes; plan.blk_list = d_list.p; plan.blk_off = d_boff.p;
    plan.n_records = N; plan.perm_a = a; plan.perm_c = N / 3;
    SS_CUDA(ss_launch_synth_db(*p, plan, d_text.p, node_of_record ? d_node.p : nullptr, c->stream));
    SS_CUDA(cudaMemcpyAsync(text_out, d_text.p, N * rec, cudaMemcpyDeviceToHost, c->stream));
    if (node_of_record)
        SS_CUDA(cudaMemcpyAsync(node_of_record, d_node.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA(cudaStreamSynchronize(c->stream));
    return SS_OK;
}
