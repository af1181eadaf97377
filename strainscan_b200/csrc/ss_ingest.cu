// ss_ingest.cu -- producer side of the read-file ingest (see ss_ingest.cuh).
#include "ss_ingest.cuh"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/strainscan_b200.h"
#include "ss_fastx.h"
#include "ss_inflate.cuh"
#include "ss_pgz.cuh"

// ---------------------------------------------------------------------------------------------
// FASTQ framing helpers
// ---------------------------------------------------------------------------------------------
// length of buf without trailing blank lines / whitespace
size_t ss_trim_tail(const char *buf, size_t len) {
    while (len > 0 && (buf[len - 1] == '\n' || buf[len - 1] == '\r' || buf[len - 1] == ' ' || buf[len - 1] == '\t'))
        len--;
    return len;
}

// first byte >= from that starts a FASTQ record: a line opening with '@' whose line+2 opens with '+'
// (a quality line may open with '@', but then line+2 is a sequence line, never '+')
size_t ss_find_record_start(const char *buf, size_t len, size_t from) {
    if (from == 0) return 0;
    if (from >= len) return len;
    size_t p = from;
    if (buf[p - 1] != '\n') {
        const char *nl = (const char *)memchr(buf + p, '\n', len - p);
        if (!nl) return len;
        p = (size_t)(nl - buf) + 1;
    }
    while (p < len) {
        const char *n1 = (const char *)memchr(buf + p, '\n', len - p);
        if (!n1) return len;
        size_t l1 = (size_t)(n1 - buf) + 1;
        if (buf[p] == '@' && l1 < len) {
            const char *n2 = (const char *)memchr(buf + l1, '\n', len - l1);
            if (!n2) return len;
            size_t l2 = (size_t)(n2 - buf) + 1;
            if (l2 < len && buf[l2] == '+') return p;
        }
        p = l1;
    }
    return len;
}

// a record start in (lo, hi) to cut a chunk at, as late as a cheap search finds one: the window before `hi` starts at
// SS_INGEST_BOUNDARY and is widened 4x at a time down to `lo` (long-read files: one record can be megabytes).
// Returns 0 when [lo, hi) holds no record start after lo.
size_t ss_find_cut(const char *buf, size_t lo, size_t hi) {
    for (size_t win = SS_INGEST_BOUNDARY;; win *= 4) {
        const bool whole = win >= hi - lo;
        size_t from = whole ? lo + 1 : hi - win;
        size_t cut = ss_find_record_start(buf, hi, from);
        if (cut < hi) return cut;
        if (whole) return 0;
    }
}

static bool ends_with_gz(const char *path) {   // identify.py:81: re.split('\.', path)[-1] == 'gz'
    const char *dot = strrchr(path, '.');
    return dot && strcmp(dot + 1, "gz") == 0;
}

int ss_read_whole_file(const char *path, std::vector<char> &out, std::string &err) {
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return SS_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> raw((size_t)sz + 16, 0);
    size_t rd = sz ? fread(raw.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (rd != (size_t)sz) { err = std::string("short read on ") + path; return SS_ERR_IO; }
    bool gz = ends_with_gz(path) || (sz >= 2 && (unsigned char)raw[0] == 0x1f && (unsigned char)raw[1] == 0x8b);
    if (!gz) { out.insert(out.end(), raw.begin(), raw.begin() + sz); return SS_OK; }
    ssi_gz_stream *g = new ssi_gz_stream;
    ssi_gz_init(*g, (const uint8_t *)raw.data(), (size_t)sz);
    const size_t win = 8u << 20;
    std::vector<uint8_t> buf(SS_INGEST_HIST + win);
    uint8_t *text = buf.data() + SS_INGEST_HIST;
    int rc;
    do {
        uint8_t *pos = text;
        rc = ssi_gz_read(*g, &pos, text + win);
        size_t got = (size_t)(pos - text);
        out.insert(out.end(), (char *)text, (char *)pos);
        if (got >= SS_INGEST_HIST) memcpy(text - SS_INGEST_HIST, pos - SS_INGEST_HIST, SS_INGEST_HIST);
        else if (got) { memmove(text - SS_INGEST_HIST, text - SS_INGEST_HIST + got, SS_INGEST_HIST - got); memcpy(text - got, text, got); }
    } while (rc == SSI_MORE_OUTPUT);
    delete g;
    if (rc != SSI_OK) { err = std::string("inflate failed on ") + path; return SS_ERR_IO; }
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------
ss_text_source::~ss_text_source() {
    finish();
    for (auto &b : bufs_) { if (pinned_) cudaFreeHost(b.base); else free(b.base); }
}

// room behind the text area of every chunk buffer: the two host-decoded text pieces of a BGZF batch
#define SS_PIECE_BYTES (3u * 65536u + 4096u)

int ss_text_source::init(size_t chunk_bytes, int n_buffers, int n_threads, bool pinned, bool device_bgzf,
                         size_t bgzf_out_cap) {
    if (!bufs_.empty()) return SS_OK;
    if (chunk_bytes < 4 * (size_t)SS_INGEST_BOUNDARY) { err_msg_ = "ingest: chunk size must be at least 256 KiB"; return SS_ERR_ARG; }
    chunk_bytes = (chunk_bytes + 15) & ~(size_t)15;
    chunk_bytes_ = chunk_bytes;
    n_threads_ = std::max(1, n_threads);
    pinned_ = pinned;
    device_bgzf_ = device_bgzf && chunk_bytes >= (1u << 20) && bgzf_out_cap >= (1u << 18);
    bgzf_out_cap_ = bgzf_out_cap;
    bufs_.resize((size_t)std::max(3, n_buffers));
    for (auto &b : bufs_) {
        void *p = nullptr;
        const size_t bytes = SS_INGEST_HIST + chunk_bytes + 4096 + 2 * (size_t)SS_PIECE_BYTES +
                             (size_t)SS_BGZF_MAX_MEMBERS * sizeof(ss_member);
        cudaError_t e = cudaSuccess;
        if (pinned) e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        else p = malloc(bytes);
        if (e != cudaSuccess || !p) {
            for (auto &x : bufs_) { if (pinned) cudaFreeHost(x.base); else free(x.base); }
            bufs_.clear();
            err_msg_ = std::string("ingest: cannot allocate host chunk buffers: ") + (pinned ? cudaGetErrorString(e) : "out of memory");
            return SS_ERR_NOMEM;
        }
        b.base = (uint8_t *)p;
        b.text = b.base + SS_INGEST_HIST;
        b.cap = chunk_bytes;
        b.len = 0;
        b.members = (ss_member *)(b.text + chunk_bytes + 4096 + 2 * (size_t)SS_PIECE_BYTES);
    }
    return SS_OK;
}

void ss_text_source::fail(int code, const std::string &msg) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!err_code_) { err_code_ = code; err_msg_ = msg; }
    stop_ = true;
    cv_free_.notify_all();
    cv_ready_.notify_all();
}

// the record start at or after byte `off` of a plain file (off itself when it is 0 or the file end)
int ss_text_source::plain_boundary(const file_map &f, size_t off, size_t *out) {
    if (off == 0 || off >= f.size) { *out = std::min(off, f.size); return SS_OK; }
    size_t win = 4 * (size_t)SS_INGEST_BOUNDARY;
    while (true) {
        size_t lo = off - 1, n = std::min(win, f.size - lo);
        std::vector<char> w(n);
        ssize_t got = pread(f.fd, w.data(), n, (off_t)lo);
        if (got != (ssize_t)n) { err_msg_ = "short read on " + f.path; return SS_ERR_IO; }
        size_t p = ss_find_record_start(w.data(), n, 1);
        if (p < n) { *out = lo + p; return SS_OK; }
        if (lo + n >= f.size) { *out = f.size; return SS_OK; }          // no further record: the rest joins the previous part
        if (win >= (64u << 20)) { err_msg_ = f.path + ": no FASTQ record boundary within 64 MiB"; return SS_ERR_FORMAT; }
        win *= 4;
    }
}

static size_t bgzf_resync(const uint8_t *p, size_t size, size_t off);

// ---- multi-member gzip: members found by their header ------------------------------------------------
// Does a gzip member start at byte `at`?  Magic, method, reserved flag bits, a header that parses -- and a deflate
// stream behind it that decodes: the first 64 KiB of text (or the whole member, if shorter) are inflated with an EMPTY
// history, so in bytes that merely look like a header the first match that reaches back before the "member start",
// the first unused code or the first broken block header ends the trial; a stream that ends inside the trial must be
// followed by its own length (ISIZE).  (A header check alone is not enough: a
// fixed-Huffman block "header" is valid whatever follows, and the 3-byte magic turns up every 2^27 bytes of
// compressed data.)
static bool gz_member_at(const uint8_t *p, size_t size, size_t at, ssi_tables &t) {
    if (at + 18 > size || p[at] != 0x1f || p[at + 1] != 0x8b || p[at + 2] != 8 || (p[at + 3] & 0xE0)) return false;
    ssi_gz_header h;
    if (ssi_gz_parse_header(p + at, p + size, &h) != SSI_OK) return false;
    ssi_stream s;
    ssi_stream_init(s, p + at + h.header_len, p + size);
    std::vector<uint8_t> buf((64u << 10) + 4 * SSI_OUT_SLACK);
    uint8_t *pos = buf.data();
    const int rc = ssi_inflate(s, t, &pos, buf.data() + buf.size());
    if (rc == SSI_MORE_OUTPUT) return true;
    if (rc != SSI_OK) return false;
    // a member that ends inside the trial: its trailer must carry the length just produced
    ssi_drop(s.bits, s.bits.cnt & 7u);
    const uint8_t *q = ssi_in_pos(s.bits);
    if (p + size - q < 8) return false;
    const uint32_t isize = (uint32_t)q[4] | ((uint32_t)q[5] << 8) | ((uint32_t)q[6] << 16) | ((uint32_t)q[7] << 24);
    return isize == (uint32_t)s.out_total;
}

// first member start in [from, limit), or `size` when there is none (a member needs 18 bytes: header + trailer)
static size_t gz_next_member(const uint8_t *p, size_t size, size_t from, size_t limit, ssi_tables &t) {
    limit = std::min(limit, size >= 18 ? size - 17 : 0);
    for (size_t at = from; at < limit;) {
        const uint8_t *q = (const uint8_t *)memchr(p + at, 0x1f, limit - at);
        if (!q) return size;
        at = (size_t)(q - p);
        if (gz_member_at(p, size, at, t)) return at;
        at++;
    }
    return size;
}

size_t ss_gz_next_member_start(const uint8_t *map, size_t size, size_t from, size_t limit) {
    ssi_tables *t = new ssi_tables;
    const size_t r = gz_next_member(map, size, from, limit, *t);
    delete t;
    return r;
}

bool ss_gz_is_member_split(const uint8_t *map, size_t size, int n_shards) {
    int want_split = 1;
    if (const char *e = getenv("SS_GZ_SPLIT")) want_split = atoi(e);
    if (!want_split || size < (1u << 20)) return false;
    const size_t m2 = ss_gz_next_member_start(map, size, 1, 256u << 20);
    return m2 < size && (want_split == 2 || size / m2 >= 4 * (size_t)std::max(1, n_shards));
}

size_t ss_gz_split_parts(size_t size);

// fixed parts of a member-split gzip file: a function of the file size only, so every rank (whatever its thread
// count) cuts the file the same way
size_t ss_gz_split_parts(size_t size) {
    size_t part = 32u << 20;
    if (const char *e = getenv("SS_GZ_PART_BYTES")) { long long v = atoll(e); if (v >= (64 << 10)) part = (size_t)v; }
    return std::max<size_t>(1, std::min<size_t>(4096, size / part));
}

int ss_text_source::start(const char *const *paths, int n_paths, int shard, int n_shards) {
    if (bufs_.empty()) { err_msg_ = "ingest: init() was not called"; return SS_ERR_ARG; }
    if (n_shards < 1 || shard < 0 || shard >= n_shards) { err_msg_ = "reads: bad shard / n_shards"; return SS_ERR_ARG; }
    finish();
    err_code_ = 0; err_msg_.clear(); stop_ = false;
    shard_ = shard; n_shards_ = n_shards;
    raw_mode_ = false;
    plain_bytes_ = gz_bytes_ = 0;
    n_gz_jobs_ = 0;
    files_.clear(); jobs_.clear(); next_job_ = 0;
    free_.clear(); ready_.clear();
    for (auto &b : bufs_) free_.push_back(&b);

    for (int i = 0; i < n_paths; i++) {
        file_map f;
        f.path = paths[i];
        f.fd = open(paths[i], O_RDONLY);
        if (f.fd < 0) { err_msg_ = std::string("cannot open ") + paths[i]; finish(); return SS_ERR_IO; }
        struct stat st;
        if (fstat(f.fd, &st) != 0) { err_msg_ = std::string("cannot stat ") + paths[i]; close(f.fd); finish(); return SS_ERR_IO; }
        f.size = (size_t)st.st_size;
        unsigned char magic[2] = {0, 0};
        ssize_t got = pread(f.fd, magic, 2, 0);
        f.gz = ends_with_gz(paths[i]) || (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b);
        if (f.gz && f.size) {
            void *m = mmap(nullptr, f.size, PROT_READ, MAP_PRIVATE, f.fd, 0);
            if (m == MAP_FAILED) { err_msg_ = std::string("cannot mmap ") + paths[i]; close(f.fd); finish(); return SS_ERR_IO; }
            madvise(m, f.size, MADV_SEQUENTIAL);
            f.map = (const uint8_t *)m;
            ssi_gz_header h;
            f.bgzf = device_bgzf_ && ssi_gz_parse_header(f.map, f.map + f.size, &h) == SSI_OK && h.bgzf_bsize >= h.header_len + 8 &&
                     h.bgzf_bsize <= f.size;
            // A file of many gzip members (lane files concatenated, block-wise compressors) is split at member
            // starts: decided from the distance to the second member alone, i.e. identically on every rank.
            if (!f.bgzf) f.split = ss_gz_is_member_split(f.map, f.size, n_shards);
        }
        if (f.size) {   // dialect from the head of the text: FASTA and wrapped FASTQ are rewritten on the host (ss_fastx.h)
            std::vector<uint8_t> head(SS_INGEST_HIST + (64u << 10));
            size_t hn = 0;
            bool whole = false;
            if (f.gz) {
                ssi_gz_stream *g = new ssi_gz_stream;
                ssi_gz_init(*g, f.map, f.size);
                uint8_t *pos = head.data() + SS_INGEST_HIST;
                int rc = ssi_gz_read(*g, &pos, head.data() + head.size());
                delete g;
                hn = (size_t)(pos - (head.data() + SS_INGEST_HIST));
                whole = rc == SSI_OK;
                if (rc < 0) hn = 0;                              // the producer reports the inflate error
            } else {
                ssize_t got2 = pread(f.fd, head.data() + SS_INGEST_HIST, 64u << 10, 0);
                hn = got2 > 0 ? (size_t)got2 : 0;
                whole = hn == f.size;
            }
            int kind = hn ? ss_fastx_kind((const char *)head.data() + SS_INGEST_HIST, hn, whole) : 0;
            if (kind < 0) {
                err_msg_ = f.path + ": not a FASTQ / FASTA file (first byte is neither '@' nor '>')";
                if (f.map) munmap((void *)f.map, f.size);
                close(f.fd);
                finish();
                return SS_ERR_FORMAT;
            }
            if (kind == 1) { f.normalize = true; f.bgzf = false; }
        }
        files_.push_back(f);
    }
    // jobs: one per gzip stream (serial by nature); plain files in record-aligned parts
    for (size_t fi = 0; fi < files_.size(); fi++) {
        const file_map &f = files_[fi];
        if (f.size == 0) continue;
        if (f.normalize) {
            job j; j.file = (int)fi; j.lo = 0; j.hi = f.size; j.first_of_file = true;
            jobs_.push_back(j);
            gz_bytes_ += f.size;                                 // size after rewriting is unknown: grow by segments
            continue;
        }
        if (f.bgzf) {
            // members inflate independently: cut the file into parts at member starts (found by their
            // 16-byte BGZF header signature and confirmed by walking the chain), one producer per part
            // (part size is a constant, not a function of the thread count: with several ranks the batches are dealt
            // by part and batch number, and every rank must cut the file the same way)
            size_t part_min = 128u << 20;
            if (const char *e = getenv("SS_BGZF_PART_BYTES")) { long long v = atoll(e); if (v >= (256 << 10)) part_min = (size_t)v; }
            int parts = (int)std::max<size_t>(1, std::min<size_t>(4096, f.size / (part_min + 1) + 1));
            size_t prev = 0;
            for (int pi = 1; pi <= parts; pi++) {
                size_t b = pi == parts ? f.size : bgzf_resync(f.map, f.size, (size_t)((unsigned __int128)f.size * pi / parts));
                if (b > prev) {
                    job j; j.file = (int)fi; j.lo = prev; j.hi = b; j.first_of_file = (prev == 0);
                    jobs_.push_back(j);
                }
                prev = std::max(prev, b);
            }
            gz_bytes_ += f.size;
            continue;
        }
        if (f.gz && f.split) {
            // fixed parts; rank r owns the members that START in parts [P r / N, P (r + 1) / N)
            const size_t P = ss_gz_split_parts(f.size);
            const size_t p_lo = P * (size_t)shard_ / (size_t)n_shards_, p_hi = P * ((size_t)shard_ + 1) / (size_t)n_shards_;
            for (size_t pi = p_lo; pi < p_hi; pi++) {
                job j; j.file = (int)fi;
                j.lo = (size_t)((unsigned __int128)f.size * pi / P);
                j.hi = pi + 1 == P ? f.size : (size_t)((unsigned __int128)f.size * (pi + 1) / P);
                j.first_of_file = (pi == 0);
                if (j.hi > j.lo) jobs_.push_back(j);
            }
            gz_bytes_ += f.size;
            continue;
        }
        if (f.gz) {
            job j; j.file = (int)fi; j.lo = 0; j.hi = f.size; j.first_of_file = true;
            jobs_.push_back(j);
            gz_bytes_ += f.size;
            n_gz_jobs_++;
            continue;
        }
        // parts per rank: enough to keep the producers busy, at least ~4 chunks each
        size_t mine = f.size / (size_t)n_shards_;
        int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads_, mine / (4 * chunk_bytes_ + 1) + 1));
        size_t total_parts = (size_t)parts * (size_t)n_shards_;
        size_t prev = 0;
        int rc = plain_boundary(f, (size_t)((unsigned __int128)f.size * ((size_t)shard_ * parts) / total_parts), &prev);
        if (rc) { finish(); return rc; }
        for (int p = 0; p < parts; p++) {
            size_t idx = (size_t)shard_ * parts + p + 1, b = f.size;
            if (idx < total_parts) {
                rc = plain_boundary(f, (size_t)((unsigned __int128)f.size * idx / total_parts), &b);
                if (rc) { finish(); return rc; }
            }
            if (b > prev) {
                job j; j.file = (int)fi; j.lo = prev; j.hi = b; j.first_of_file = (prev == 0);
                jobs_.push_back(j);
                plain_bytes_ += b - prev;
            }
            prev = std::max(prev, b);
        }
    }
    int nt = (int)std::min<size_t>((size_t)n_threads_, jobs_.size());
    // every producer holds one buffer while it fills it: keep at least one more for the consumer side
    nt = std::min(nt, std::max(1, (int)bufs_.size() - 2));
    active_ = nt;
    for (int t = 0; t < nt; t++) threads_.emplace_back([this]() { worker(); });
    return SS_OK;
}

int ss_text_source::start_raw(const char *path, size_t lo, size_t hi) {
    if (bufs_.empty()) { err_msg_ = "ingest: init() was not called"; return SS_ERR_ARG; }
    finish();
    err_code_ = 0; err_msg_.clear(); stop_ = false;
    shard_ = 0; n_shards_ = 1; plain_bytes_ = gz_bytes_ = 0; n_gz_jobs_ = 0;
    files_.clear(); jobs_.clear(); next_job_ = 0;
    free_.clear(); ready_.clear();
    for (auto &b : bufs_) free_.push_back(&b);
    file_map f;
    f.path = path;
    f.fd = open(path, O_RDONLY);
    if (f.fd < 0) { err_msg_ = std::string("cannot open ") + path; return SS_ERR_IO; }
    struct stat st;
    if (fstat(f.fd, &st) != 0) { err_msg_ = std::string("cannot stat ") + path; close(f.fd); return SS_ERR_IO; }
    f.size = (size_t)st.st_size;
    files_.push_back(f);
    hi = std::min(hi, f.size);
    for (size_t at = lo; at < hi; at += chunk_bytes_) {
        job j; j.file = 0; j.lo = at; j.hi = std::min(hi, at + chunk_bytes_);
        jobs_.push_back(j);
    }
    raw_mode_ = true;
    int nt = (int)std::min<size_t>((size_t)n_threads_, jobs_.size());
    nt = std::min(nt, std::max(1, (int)bufs_.size() - 2));
    active_ = nt;
    for (int t = 0; t < nt; t++) threads_.emplace_back([this]() { worker(); });
    return SS_OK;
}

void ss_text_source::run_raw(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    ss_chunk *c = acquire();
    if (!c) return;
    const size_t want = j.hi - j.lo;
    for (size_t have = 0; have < want;) {
        ssize_t got = pread(f.fd, c->text + have, want - have, (off_t)(j.lo + have));
        if (got <= 0) { release(c); fail(SS_ERR_IO, "short read on " + f.path); return; }
        have += (size_t)got;
    }
    c->kind = SS_CHUNK_RAW;
    c->file_off = j.lo;
    c->len = want;
    c->inflated_len = want;                      // text_len() of a non-TEXT chunk: pre + inflated + post
    emit(c);
}

int ss_text_source::finish() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
        cv_free_.notify_all();
        cv_ready_.notify_all();
    }
    for (auto &t : threads_) t.join();
    threads_.clear();
    for (auto &f : files_) {
        if (f.map) munmap((void *)f.map, f.size);
        if (f.fd >= 0) close(f.fd);
    }
    files_.clear();
    return err_code_;
}

// ---------------------------------------------------------------------------------------------
// queue
// ---------------------------------------------------------------------------------------------
ss_chunk *ss_text_source::acquire() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_free_.wait(lk, [this]() { return stop_ || !free_.empty(); });
    if (stop_) return nullptr;
    ss_chunk *c = free_.front();
    free_.pop_front();
    c->len = 0;
    c->kind = SS_CHUNK_TEXT;
    c->n_members = 0;
    c->pre_text = c->post_text = nullptr;
    c->pre_len = c->post_len = c->inflated_len = 0;
    return c;
}

void ss_text_source::emit(ss_chunk *c) {
    std::lock_guard<std::mutex> lk(mu_);
    if (c->text_len() == 0) { free_.push_back(c); cv_free_.notify_one(); return; }
    ready_.push_back(c);
    cv_ready_.notify_one();
}

void ss_text_source::release(ss_chunk *c) {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back(c);
    cv_free_.notify_one();
}

ss_chunk *ss_text_source::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_ready_.wait(lk, [this]() { return err_code_ || !ready_.empty() || active_ == 0; });
    if (err_code_ || ready_.empty()) return nullptr;
    ss_chunk *c = ready_.front();
    ready_.pop_front();
    return c;
}

void ss_text_source::worker() {
    while (true) {
        job j;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (stop_ || next_job_ >= jobs_.size()) break;
            j = jobs_[next_job_++];
        }
        const file_map &f = files_[(size_t)j.file];
        if (raw_mode_) run_raw(j);
        else if (f.normalize) run_normalize(j);
        else if (f.bgzf) run_bgzf(j);
        else if (f.gz && f.split) run_gz_members(j);
        else if (f.gz) {
            // an ordinary gzip stream: decoded by several threads per round when cores are to spare (ss_pgz.cuh)
            int t = std::max(1, n_threads_ / std::max(1, n_gz_jobs_));
            size_t span = 4u << 20;                                  // compressed bytes per piece (measured: 4 MiB > 2 MiB)
            if (const char *e = getenv("SS_PGZ_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) t = v; }
            if (const char *e = getenv("SS_PGZ_SPAN")) { long long v = atoll(e); if (v >= (64 << 10)) span = (size_t)v; }
            if (t >= 2 && f.size >= 4 * span) run_gz_parallel(j, t, span);
            else run_gz(j);
        }
        else run_plain(j);
    }
    std::lock_guard<std::mutex> lk(mu_);
    active_--;
    cv_ready_.notify_all();
}

// ---------------------------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------------------------
static int check_head(const uint8_t *text, size_t len, const std::string &path, std::string &msg) {
    if (len == 0) return SS_OK;
    if (text[0] == '>') {
        msg = path + ": FASTA read input is not supported by the GPU path (convert to 4-line FASTQ)";
        return SS_ERR_FORMAT;
    }
    if (text[0] != '@') { msg = path + ": not a FASTQ file (first byte is not '@')"; return SS_ERR_FORMAT; }
    return SS_OK;
}

// Cut a filled chunk at its last record start inside the boundary window; the tail is carried over.
// Returns the cut (== fill when `last`, after trimming blank tail lines and closing the last line).
static size_t cut_chunk(ss_chunk *c, size_t fill, bool last, bool file_end = true) {
    char *t = (char *)c->text;
    if (last && !file_end) return fill;      // an inner part of a plain file ends right before a record start
    if (last) {
        size_t n = ss_trim_tail(t, fill);
        if (n) t[n++] = '\n';       // slack behind cap
        return n;
    }
    return ss_find_cut(t, 0, fill);   // 0: no boundary found
}

void ss_text_source::run_plain(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    size_t pos = j.lo;
    std::vector<uint8_t> carry;
    bool first = j.first_of_file;
    while (pos < j.hi || !carry.empty()) {
        ss_chunk *c = acquire();
        if (!c) return;
        size_t fill = carry.size();
        if (fill) memcpy(c->text, carry.data(), fill);
        carry.clear();
        size_t want = std::min(c->cap - fill, j.hi - pos);
        size_t done = 0;
        while (done < want) {
            ssize_t got = pread(f.fd, c->text + fill + done, want - done, (off_t)(pos + done));
            if (got <= 0) { release(c); fail(SS_ERR_IO, "short read on " + f.path); return; }
            done += (size_t)got;
        }
        pos += want; fill += want;
        if (first) {
            std::string m;
            int rc = check_head(c->text, ss_trim_tail((const char *)c->text, fill), f.path, m);
            if (rc) { release(c); fail(rc, m); return; }
            first = false;
        }
        const bool last = pos >= j.hi;
        size_t cut = cut_chunk(c, fill, last, j.hi >= f.size);
        if (!last) {
            if (cut == 0) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within a chunk (a record longer than SS_CHUNK_BYTES?)"); return; }
            carry.assign(c->text + cut, c->text + fill);
        }
        c->len = cut;
        emit(c);
    }
}

// One thread decodes the stream; the text goes through the same stream_writer as the parallel decoder's, so both
// cut the stream into the SAME chunks (filled exactly to the chunk size, cut at the last record start): the
// round-robin dealing of a stream that every rank decodes does not depend on which decoder a rank picked.
void ss_text_source::run_gz(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    ssi_gz_stream *g = new ssi_gz_stream;
    ssi_gz_init(*g, f.map, f.size);
    stream_writer w;
    w.src = this; w.j = &j; w.fill_threads = 1;
    const size_t win = 4u << 20;
    std::vector<uint8_t> buf(SS_INGEST_HIST + win + 2 * SSI_OUT_SLACK);
    uint8_t *text = buf.data() + SS_INGEST_HIST;
    while (true) {
        uint8_t *pos = text;
        int rc = ssi_gz_read(*g, &pos, text + win + 2 * SSI_OUT_SLACK);
        const size_t got = (size_t)(pos - text);
        if (rc < 0) {
            w.abandon();
            const char *why = rc == SSI_ERR_TRUNC ? "unexpected end of file" : rc == SSI_ERR_HEADER ? "not in gzip format"
                              : rc == SSI_ERR_SIZE ? "length error" : "invalid compressed data";
            fail(SS_ERR_IO, "inflate failed on " + f.path + ": " + why);
            break;
        }
        if (got && !w.append(got, [text](uint8_t *dst, size_t off, size_t len) { memcpy(dst, text + off, len); })) break;
        if (rc == SSI_OK) { w.finish(); break; }
        // keep the last 32 KiB in front of the window for the matches of the next round
        if (got >= SS_INGEST_HIST) memcpy(text - SS_INGEST_HIST, pos - SS_INGEST_HIST, SS_INGEST_HIST);
        else { memmove(text - SS_INGEST_HIST, text - SS_INGEST_HIST + got, SS_INGEST_HIST - got); memcpy(text - got, text, got); }
    }
    delete g;
}

// Member-split mode: this job owns the gzip members that START in [j.lo, j.hi) of the file.  They are decoded one
// after the other straight into the chunk buffers (32 KiB of history carried from chunk to chunk); all chunks are
// this rank's.  Members need not hold whole records.  The same rule as for BGZF parts settles who owns what: the
// text of a part's first member belongs to the PREVIOUS part up to the first record start behind the member's first
// byte, so the owner of a part drops that head and, at the other end, decodes on into the next part's first member
// until it has seen that record start (same search, same bytes: both owners agree without talking to each other).
// Checked, loudly: the member the job stops in front of is what the next part's owner will find as its first.
void ss_text_source::run_gz_members(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    ssi_tables *tabs = new ssi_tables;
    const size_t start = j.lo == 0 ? 0 : gz_next_member(f.map, f.size, j.lo, j.hi, *tabs);
    if (start >= j.hi) { delete tabs; return; }                       // no member starts in this part
    ssi_gz_stream *g = new ssi_gz_stream;
    ssi_gz_init(*g, f.map + start, f.size - start);
    g->stop_p = f.map + j.hi;
    ss_chunk *c = acquire();
    if (!c) { delete g; delete tabs; return; }
    auto why = [](int rc) {
        return rc == SSI_ERR_TRUNC ? "unexpected end of file" : rc == SSI_ERR_HEADER ? "not in gzip format"
               : rc == SSI_ERR_SIZE ? "length error" : "invalid compressed data";
    };
    std::vector<uint8_t> carry;
    size_t fill = 0;
    size_t drop = j.lo == 0 ? 0 : (size_t)-1;        // bytes of the first chunk that belong to the previous part (-1: not known yet)
    size_t mark = (size_t)-1;                        // offset in the current chunk where the NEXT part's first member begins
    bool file_end = false, done = false;
    while (!done) {
        // decode: up to the part boundary first (stop_p), then little by little into the next part's first member
        uint8_t *pos = c->text + fill;
        uint8_t *lim = mark == (size_t)-1 ? c->text + c->cap : std::min(c->text + c->cap, pos + (64u << 10) + 2 * SSI_OUT_SLACK);
        int rc = ssi_gz_read(*g, &pos, lim);
        fill = (size_t)(pos - c->text);
        if (rc < 0) { release(c); fail(SS_ERR_IO, "inflate failed on " + f.path + ": " + why(rc)); break; }
        size_t cut = 0;
        bool emit_all = false;
        if (rc == SSI_OK && mark == (size_t)-1) {
            // the decoder stands between two members at or behind j.hi, at the file end, or in front of bytes that are
            // no member (zcat: "trailing garbage ignored" -- the same thing here only if no member follows them)
            const size_t end = (size_t)(g->p - f.map);
            file_end = end >= f.size;
            if (!file_end && !gz_member_at(f.map, f.size, end, *tabs)) {
                if (gz_next_member(f.map, f.size, end + 1, f.size, *tabs) < f.size) {
                    release(c); fail(SS_ERR_IO, f.path + ": bytes that are no gzip member are followed by more members; set SS_GZ_SPLIT=0"); break;
                }
                file_end = true;
            }
            // (at the file end: nobody may find a member start inside the bytes I decoded as part of my last member)
            if (file_end ? (j.hi < f.size && gz_next_member(f.map, f.size, j.hi, f.size, *tabs) < f.size)
                         : (end < j.hi || gz_next_member(f.map, f.size, j.hi, end + 1, *tabs) != end)) {
                release(c); fail(SS_ERR_IO, f.path + ": gzip member boundaries are ambiguous (a member header pattern inside compressed data); set SS_GZ_SPLIT=0"); break;
            }
            if (file_end) { emit_all = true; }
            else { mark = fill; g->stop_p = nullptr; continue; }      // go on into the next part's first member
        }
        if (mark != (size_t)-1 && !emit_all) {
            // looking for the first record start behind the first byte of the next part's first member
            size_t r = ss_find_record_start((const char *)c->text, fill, mark + 1);
            if (r < fill) { cut = r; done = true; }
            else if (rc == SSI_OK) { file_end = true; emit_all = true; }             // the file ends before another record starts
            else if (fill + (64u << 10) + 4 * SSI_OUT_SLACK <= c->cap) continue;      // decode some more into this chunk
        }
        if (emit_all) {
            cut = ss_trim_tail((const char *)c->text, fill);
            if (cut) c->text[cut++] = '\n';
            done = true;
        }
        size_t tail = 0, hist = 0;
        if (!done) {                                                   // chunk full: cut at the last record start, carry the tail
            cut = cut_chunk(c, fill, false);
            if (cut == 0) {
                release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within a chunk (a record longer than SS_CHUNK_BYTES?)"); break;
            }
            // the tail and the 32 KiB of history the decoder needs go through a private buffer: this thread must not
            // sit on its full chunk while it waits for an empty one (all producers doing that at once would starve
            // the consumer, which hands buffers back only as its copies complete)
            tail = fill - cut; hist = std::min(fill, tail + (size_t)SS_INGEST_HIST);
            carry.assign(c->text + fill - hist, c->text + fill);
            if (mark != (size_t)-1) {
                if (mark >= cut) mark -= cut;                          // the boundary member begins in the carried tail
                else { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record start within a chunk behind a gzip member start"); break; }
            }
        }
        // the head of the job's first chunk belongs to the previous part
        size_t off = 0;
        if (j.lo == 0 && drop == 0) {                                  // the file's first chunk: is this FASTQ at all?
            std::string m;
            int hrc = check_head(c->text, ss_trim_tail((const char *)c->text, cut), f.path, m);
            if (hrc) { release(c); fail(hrc, m); break; }
            drop = 1;                                                  // checked
        }
        if (drop == (size_t)-1) {
            off = ss_find_record_start((const char *)c->text, cut, 1);
            if (off >= cut && !done) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record start within a chunk behind a gzip member start"); break; }
            off = std::min(off, cut);
            drop = off;
        }
        if (off) memmove(c->text, c->text + off, cut - off);
        c->len = cut - off;
        emit(c);
        c = nullptr;
        if (done) break;
        c = acquire();
        if (!c) break;
        memcpy(c->text + tail - hist, carry.data(), hist);            // history in front of the text area, tail at its start
        fill = tail;
    }
    delete g;
    delete tabs;
}

// ---------------------------------------------------------------------------------------------
// stream -> chunks, and the parallel gzip producer
// ---------------------------------------------------------------------------------------------
bool ss_text_source::stream_writer::append(size_t n, const std::function<void(uint8_t *, size_t, size_t)> &fill_fn) {
    const file_map &f = src->files_[(size_t)j->file];
    size_t done = 0;
    while (n) {
        if (!c) {
            c = src->acquire();
            if (!c) return false;
        }
        size_t take = std::min(c->cap - fill, n);
        pgz_parallel_fill(c->text + fill, take, fill_threads,
                          [&](uint8_t *dst, size_t off, size_t len) { fill_fn(dst, done + off, len); });
        fill += take; done += take; n -= take;
        if (fill < c->cap) break;
        if (first) {
            std::string m;
            int hrc = check_head(c->text, ss_trim_tail((const char *)c->text, fill), f.path, m);
            if (hrc) { abandon(); src->fail(hrc, m); return false; }
            first = false;
        }
        size_t cut = cut_chunk(c, fill, false);
        if (cut == 0) { abandon(); src->fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within a chunk (a record longer than SS_CHUNK_BYTES?)"); return false; }
        ss_chunk *c2 = src->acquire();
        if (!c2) { abandon(); return false; }
        size_t tail = fill - cut;
        memcpy(c2->text, c->text + cut, tail);
        if ((chunk_idx + (uint64_t)j->file) % (uint64_t)src->n_shards_ == (uint64_t)src->shard_) { c->len = cut; src->emit(c); }
        else src->release(c);
        chunk_idx++;
        c = c2;
        fill = tail;
    }
    return true;
}

bool ss_text_source::stream_writer::finish() {
    if (!c) return true;
    const file_map &f = src->files_[(size_t)j->file];
    if (first) {
        std::string m;
        int hrc = check_head(c->text, ss_trim_tail((const char *)c->text, fill), f.path, m);
        if (hrc) { abandon(); src->fail(hrc, m); return false; }
    }
    size_t cut = cut_chunk(c, fill, true);
    if ((chunk_idx + (uint64_t)j->file) % (uint64_t)src->n_shards_ == (uint64_t)src->shard_) { c->len = cut; src->emit(c); }
    else src->release(c);
    c = nullptr;
    return true;
}

void ss_text_source::stream_writer::abandon() {
    if (c) src->release(c);
    c = nullptr;
}

// FASTA / wrapped FASTQ: read (and inflate) the whole file, rewrite it as 4-line FASTQ, chunk it
void ss_text_source::run_normalize(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    std::vector<char> raw, norm;
    std::string err;
    int rc = ss_read_whole_file(f.path.c_str(), raw, err);
    if (rc) { fail(rc, err); return; }
    if (ss_fastx_normalize(raw.data(), raw.size(), norm)) {
        fail(SS_ERR_FORMAT, f.path + ": malformed FASTA / multi-line FASTQ (record framing breaks)");
        return;
    }
    std::vector<char>().swap(raw);
    stream_writer w;
    w.src = this; w.j = &j; w.fill_threads = 2;
    w.first = false;                                            // the rewritten text opens with '@' by construction
    const char *src = norm.data();
    if (!w.append(norm.size(), [src](uint8_t *dst, size_t off, size_t len) { memcpy(dst, src + off, len); })) return;
    w.finish();
}

void ss_text_source::run_gz_parallel(const job &j, int threads, size_t span) {
    const file_map &f = files_[(size_t)j.file];
    stream_writer w;
    w.src = this; w.j = &j; w.fill_threads = std::max(1, threads / 2);
    size_t p = 0;
    uint64_t n_members = 0;
    auto why = [](int rc) {
        return rc == SSI_ERR_TRUNC ? "unexpected end of file" : rc == SSI_ERR_HEADER ? "not in gzip format"
               : rc == SSI_ERR_SIZE ? "length error" : "invalid compressed data";
    };
    while (p < f.size) {
        ssi_gz_header h;
        int rc = ssi_gz_parse_header(f.map + p, f.map + f.size, &h);
        if (rc == SSI_ERR_HEADER && n_members) break;               // trailing garbage, as zcat
        if (rc) { w.abandon(); fail(SS_ERR_IO, "inflate failed on " + f.path + ": " + why(rc)); return; }
        pgz_member m;
        m.base = f.map; m.size = f.size; m.bit = (uint64_t)(p + h.header_len) * 8u;
        bool ok = true;
        struct report {                                             // SS_DEBUG_TIMING: where a member's time went
            pgz_member &m; int threads;
            ~report() {
                if (getenv("SS_DEBUG_TIMING") && m.rounds > 1)
                    fprintf(stderr, "[ss pgz] %d threads, %llu rounds, %llu/%llu pieces used, %.1f MB out (%.1f MB through markers): "
                            "find %.3f s, decode %.3f s, stitch+emit %.3f s\n", threads, (unsigned long long)m.rounds,
                            (unsigned long long)m.pieces_used, (unsigned long long)m.pieces_found, m.member_out / 1e6,
                            m.marker_syms / 1e6, m.t_find, m.t_decode, m.t_stitch);
            }
        } rep{m, threads};
        rc = pgz_member_decode(m, threads, span,
                               [&](size_t n, const std::function<void(uint8_t *, size_t, size_t)> &fill) { if (ok) ok = w.append(n, fill); },
                               [&]() { std::lock_guard<std::mutex> lk(mu_); return ok && !stop_; });
        if (!ok) return;                                            // stopped, or append() reported the failure
        if (rc == 0) { w.abandon(); return; }                       // asked to stop
        if (rc < 0) { w.abandon(); fail(SS_ERR_IO, "inflate failed on " + f.path + ": " + why(rc)); return; }
        size_t q = (size_t)((m.bit + 7u) >> 3);
        if (f.size - q < 8) { w.abandon(); fail(SS_ERR_IO, "inflate failed on " + f.path + ": unexpected end of file"); return; }
        const uint8_t *tr = f.map + q;
        uint32_t isize = (uint32_t)tr[4] | ((uint32_t)tr[5] << 8) | ((uint32_t)tr[6] << 16) | ((uint32_t)tr[7] << 24);
        if (isize != (uint32_t)m.member_out) { w.abandon(); fail(SS_ERR_IO, "inflate failed on " + f.path + ": length error"); return; }
        p = q + 8;
        n_members++;
    }
    if (n_members == 0) { w.abandon(); fail(SS_ERR_IO, "inflate failed on " + f.path + ": not in gzip format"); return; }
    w.finish();
}

// ---------------------------------------------------------------------------------------------
// BGZF (blocked gzip: every member carries its own size in a "BC" extra field and inflates on its own).
// The producer only walks the member chain and stages compressed bytes; the DEVICE inflates
// (ss_gunzip.cu).  To hand out batches that begin and end on FASTQ record boundaries without a device
// round trip, the member at every batch boundary is inflated here, on the host, and split at a record
// start: its head closes this batch (post_text), its tail opens the next one (pre_text).
// ---------------------------------------------------------------------------------------------
namespace {
struct bgzf_member { uint32_t hdr, bsize, isize; };

int bgzf_parse(const uint8_t *p, size_t size, size_t at, bgzf_member *m) {
    ssi_gz_header h;
    if (ssi_gz_parse_header(p + at, p + size, &h) != SSI_OK || h.bgzf_bsize < h.header_len + 8 || at + h.bgzf_bsize > size) return -1;
    m->hdr = h.header_len; m->bsize = h.bgzf_bsize;
    const uint8_t *q = p + at + h.bgzf_bsize - 4;
    m->isize = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    return m->isize <= 65536u ? 0 : -1;            // BGZF blocks hold at most 64 KiB of text
}

// inflate one member on the host, appending its text to `out`
int bgzf_host_inflate(const uint8_t *p, size_t at, const bgzf_member &m, std::vector<uint8_t> &out) {
    size_t o = out.size();
    out.resize(o + m.isize + SSI_OUT_SLACK);
    ssi_stream s;
    ssi_tables t;
    ssi_stream_init(s, p + at + m.hdr, p + at + m.bsize - 8);
    uint8_t *w = out.data() + o;
    int rc = ssi_inflate(s, t, &w, out.data() + out.size());
    if (rc != SSI_OK || s.out_total != m.isize) return -1;
    out.resize(o + m.isize);
    return 0;
}
}   // namespace

// first member start at or after `off`: the fixed bytes of a BGZF header (magic, deflate, FEXTRA, XLEN = 6,
// "BC", SLEN = 2) confirmed by three chained members (or the file end)
static size_t bgzf_resync(const uint8_t *p, size_t size, size_t off) {
    for (size_t at = off; at + 18 <= size; at++) {
        const uint8_t *q = (const uint8_t *)memchr(p + at, 0x1f, size - 18 - at + 1);
        if (!q) return size;
        at = (size_t)(q - p);
        if (q[1] != 0x8b || q[2] != 8 || q[3] != 4 || q[12] != 'B' || q[13] != 'C') continue;
        size_t w = at;
        bool ok = true;
        bgzf_member m;
        for (int k = 0; k < 3 && w < size && ok; k++) {
            ok = bgzf_parse(p, size, w, &m) == 0;
            w += m.bsize;
        }
        if (ok) return at;
    }
    return size;
}

// Text of the member(s) at file offset `at` (a member start), extended by up to two more members until a
// FASTQ record start shows up behind its first byte: *cut is that record start, *next the offset of the
// first member not consumed.  The part that ends at `at` and the part that starts there both call this
// and so agree on where one stops and the other begins.  Returns 0, or -1 on a broken chain;
// *cut == text.size() when there is no record start before the file ends.
static int bgzf_boundary(const uint8_t *p, size_t size, size_t at, std::vector<uint8_t> &text, size_t *cut, size_t *next) {
    text.clear();
    *cut = 0;
    bgzf_member m;
    for (int tries = 0; at < size; tries++) {
        if (bgzf_parse(p, size, at, &m) || bgzf_host_inflate(p, at, m, text)) return -1;
        at += m.bsize;
        *cut = ss_find_record_start((const char *)text.data(), text.size(), 1);
        if (*cut < text.size() || tries == 2) break;
    }
    *next = at;
    if (*cut > text.size()) *cut = text.size();
    return 0;
}

void ss_text_source::run_bgzf(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    const uint8_t *p = f.map;                       // headers, trailers and boundary members are read through the map;
    const size_t size = f.size;                     // the bulk of the compressed bytes goes pread -> pinned buffer
    const std::string bad = "inflate failed on " + f.path + ": broken BGZF member chain";
    const size_t out_cap = bgzf_out_cap_ - 2 * (size_t)SS_PIECE_BYTES;
    std::vector<uint8_t> carry, btext;
    bgzf_member m;
    size_t pos = j.lo;
    uint64_t batch_idx = 0;
    if (j.lo == 0) {                                // head check on the first data member's text
        size_t at = 0;
        while (at < size) {
            if (bgzf_parse(p, size, at, &m)) { fail(SS_ERR_IO, bad); return; }
            if (m.isize) break;
            at += m.bsize;
        }
        if (at < size) {
            if (bgzf_host_inflate(p, at, m, btext)) { fail(SS_ERR_IO, bad); return; }
            std::string msg;
            int rc = check_head(btext.data(), ss_trim_tail((const char *)btext.data(), btext.size()), f.path, msg);
            if (rc) { fail(rc, msg); return; }
        }
    } else {                                        // a later part of the file: it owns the text behind the first record start
        size_t cut = 0;
        if (bgzf_boundary(p, size, j.lo, btext, &cut, &pos)) { fail(SS_ERR_IO, bad); return; }
        carry.assign(btext.begin() + (long)cut, btext.end());
    }
    bool done = false;
    while (!done) {
        ss_chunk *c = acquire();
        if (!c) return;
        c->kind = SS_CHUNK_BGZF;
        uint8_t *pieces = c->text + c->cap + 4096;
        memcpy(pieces, carry.data(), carry.size());
        c->pre_text = pieces; c->pre_len = carry.size();
        carry.clear();
        // members for the device: a contiguous run of the file, up to (not including) the batch's boundary member
        const size_t run_lo = pos;
        size_t run_hi = pos, out_sum = 0;
        uint32_t n = 0;
        bool err = false, part_end = false;
        while (true) {
            if (pos >= j.hi || pos >= size) { part_end = true; break; }
            if (bgzf_parse(p, size, pos, &m)) { err = true; break; }
            if (m.isize == 0) { pos += m.bsize; if (n) run_hi = pos; continue; }     // empty member (EOF marker)
            const bool fits = pos + m.bsize - run_lo + 16 <= c->cap && out_sum + m.isize <= out_cap && n < SS_BGZF_MAX_MEMBERS;
            if (!fits) break;
            // the last data member of the whole file is decoded on the host (its tail may need trimming)
            if (j.hi >= size) {
                size_t nx = pos + m.bsize;
                bgzf_member e;
                while (nx < size && bgzf_parse(p, size, nx, &e) == 0 && e.isize == 0) nx += e.bsize;
                if (nx >= size) break;
            }
            ss_member &t = c->members[n++];
            t.comp_off = (uint32_t)(pos - run_lo + m.hdr); t.comp_len = m.bsize - m.hdr - 8;
            t.out_off = (uint32_t)(c->pre_len + out_sum); t.isize = m.isize;
            out_sum += m.isize;
            pos += m.bsize;
            run_hi = pos;
        }
        if (err) { release(c); fail(SS_ERR_IO, bad); return; }
        for (size_t have = 0, want = n ? run_hi - run_lo : 0; have < want;) {
            ssize_t got = pread(f.fd, c->text + have, want - have, (off_t)(run_lo + have));
            if (got <= 0) { release(c); fail(SS_ERR_IO, "short read on " + f.path); return; }
            have += (size_t)got;
        }
        // the text that closes this batch
        uint8_t *post = pieces + SS_PIECE_BYTES;
        size_t post_len = 0;
        if (part_end && j.hi < size) {              // the next part's first member(s): mine up to the agreed record start
            size_t cut = 0, nx = 0;
            if (bgzf_boundary(p, size, j.hi, btext, &cut, &nx)) { release(c); fail(SS_ERR_IO, bad); return; }
            memcpy(post, btext.data(), cut);
            post_len = cut;
            done = true;
        } else if (part_end || pos >= size) {       // nothing left at all
            done = true;
        } else {                                    // boundary member(s) inside my part, split at a late record start
            btext.clear();
            size_t cut = 0;
            bool found = false, file_end = false;
            for (int tries = 0; tries < 3 && !found; tries++) {
                if (bgzf_parse(p, size, pos, &m) || bgzf_host_inflate(p, pos, m, btext)) { err = true; break; }
                pos += m.bsize;
                bgzf_member e;
                while (pos < size && bgzf_parse(p, size, pos, &e) == 0 && e.isize == 0) pos += e.bsize;
                if (pos >= size) { file_end = true; break; }
                const size_t bn = btext.size();
                cut = ss_find_record_start((const char *)btext.data(), bn, bn > 32768 ? bn - 32768 : 1);
                if (cut >= bn) cut = ss_find_record_start((const char *)btext.data(), bn, 1);
                found = cut < bn;
                if (pos >= j.hi) break;             // do not run into the next part's members
            }
            if (err) { release(c); fail(SS_ERR_IO, bad); return; }
            if (file_end) {                         // the file's last data member: trim blank tail lines, close the last line
                size_t bn = ss_trim_tail((const char *)btext.data(), btext.size());
                memcpy(post, btext.data(), bn);
                if (bn) post[bn++] = '\n';
                post_len = bn;
                done = true;
            } else if (found) {
                memcpy(post, btext.data(), cut);
                post_len = cut;
                carry.assign(btext.begin() + (long)cut, btext.end());
            } else if (pos >= j.hi) {               // no record start before my part ends: hand everything to the closing step
                if (btext.size() > SS_PIECE_BYTES - 8) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within three BGZF blocks"); return; }
                memcpy(post, btext.data(), btext.size());
                post_len = btext.size();
                size_t cut2 = 0, nx = 0;
                std::vector<uint8_t> more;
                if (j.hi < size) {
                    if (bgzf_boundary(p, size, j.hi, more, &cut2, &nx)) { release(c); fail(SS_ERR_IO, bad); return; }
                    if (post_len + cut2 > SS_PIECE_BYTES - 8) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within three BGZF blocks"); return; }
                    memcpy(post + post_len, more.data(), cut2);
                    post_len += cut2;
                }
                done = true;
            } else {
                release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within three BGZF blocks"); return;
            }
        }
        c->post_text = post; c->post_len = post_len;
        c->len = n ? run_hi - run_lo : 0; c->n_members = n; c->inflated_len = out_sum;
        if (n == 0) {                                   // nothing for the device: deliver it as plain text
            memmove(c->text, c->pre_text, c->pre_len);
            memcpy(c->text + c->pre_len, c->post_text, c->post_len);
            c->kind = SS_CHUNK_TEXT;
            c->len = c->pre_len + c->post_len;
            c->pre_len = c->post_len = 0;
        }
        if ((batch_idx + (uint64_t)j.file + (uint64_t)(j.lo >> 20)) % (uint64_t)n_shards_ == (uint64_t)shard_) emit(c);
        else release(c);
        batch_idx++;
    }
}
