// ss_ingest.cu -- producer side of the read-file ingest (see ss_ingest.cuh).
#include "ss_ingest.cuh"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>

#include "../../include/strainscan_b200.h"
#include "ss_inflate.cuh"

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

static bool ends_with_gz(const char *path) {   // identify.py:81: re.split('\.', path)[-1] == 'gz'
    const char *dot = strrchr(path, '.');
    return dot && strcmp(dot + 1, "gz") == 0;
}

// ---------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------
ss_text_source::~ss_text_source() {
    finish();
    for (auto &b : bufs_) { if (pinned_) cudaFreeHost(b.base); else free(b.base); }
}

int ss_text_source::init(size_t chunk_bytes, int n_buffers, int n_threads, bool pinned) {
    if (!bufs_.empty()) return SS_OK;
    if (chunk_bytes < 4 * (size_t)SS_INGEST_BOUNDARY) { err_msg_ = "ingest: chunk size must be at least 256 KiB"; return SS_ERR_ARG; }
    chunk_bytes_ = chunk_bytes;
    n_threads_ = std::max(1, n_threads);
    pinned_ = pinned;
    bufs_.resize((size_t)std::max(3, n_buffers));
    for (auto &b : bufs_) {
        void *p = nullptr;
        const size_t bytes = SS_INGEST_HIST + chunk_bytes + 4096;
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

int ss_text_source::start(const char *const *paths, int n_paths, int shard, int n_shards) {
    if (bufs_.empty()) { err_msg_ = "ingest: init() was not called"; return SS_ERR_ARG; }
    if (n_shards < 1 || shard < 0 || shard >= n_shards) { err_msg_ = "reads: bad shard / n_shards"; return SS_ERR_ARG; }
    finish();
    err_code_ = 0; err_msg_.clear(); stop_ = false;
    shard_ = shard; n_shards_ = n_shards;
    plain_bytes_ = gz_bytes_ = 0;
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
        }
        files_.push_back(f);
    }
    // jobs: one per gzip stream (serial by nature); plain files in record-aligned parts
    for (size_t fi = 0; fi < files_.size(); fi++) {
        const file_map &f = files_[fi];
        if (f.size == 0) continue;
        if (f.gz) {
            job j; j.file = (int)fi; j.lo = 0; j.hi = f.size; j.first_of_file = true;
            jobs_.push_back(j);
            gz_bytes_ += f.size;
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
    return c;
}

void ss_text_source::emit(ss_chunk *c) {
    std::lock_guard<std::mutex> lk(mu_);
    if (c->len == 0) { free_.push_back(c); cv_free_.notify_one(); return; }
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
        if (files_[(size_t)j.file].gz) run_gz(j);
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
static size_t cut_chunk(ss_chunk *c, size_t fill, bool last) {
    char *t = (char *)c->text;
    if (last) {
        size_t n = ss_trim_tail(t, fill);
        if (n) t[n++] = '\n';       // slack behind cap
        return n;
    }
    size_t from = fill > SS_INGEST_BOUNDARY ? fill - SS_INGEST_BOUNDARY : 1;
    size_t cut = ss_find_record_start(t, fill, from);
    return cut >= fill ? 0 : cut;   // 0: no boundary found
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
        size_t cut = cut_chunk(c, fill, last);
        if (!last) {
            if (cut == 0) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within 64 KiB"); return; }
            carry.assign(c->text + cut, c->text + fill);
        }
        c->len = cut;
        emit(c);
    }
}

void ss_text_source::run_gz(const job &j) {
    const file_map &f = files_[(size_t)j.file];
    ssi_gz_stream *g = new ssi_gz_stream;
    ssi_gz_init(*g, f.map, f.size);
    ss_chunk *c = acquire();
    if (!c) { delete g; return; }
    size_t fill = 0;
    uint64_t chunk_idx = 0;
    bool first = true;
    while (true) {
        uint8_t *pos = c->text + fill;
        int rc = ssi_gz_read(*g, &pos, c->text + c->cap);
        fill = (size_t)(pos - c->text);
        if (rc < 0) {
            release(c);
            const char *why = rc == SSI_ERR_TRUNC ? "unexpected end of file" : rc == SSI_ERR_HEADER ? "not in gzip format"
                              : rc == SSI_ERR_SIZE ? "length error" : "invalid compressed data";
            fail(SS_ERR_IO, "inflate failed on " + f.path + ": " + why);
            break;
        }
        const bool last = rc == SSI_OK;
        if (first) {
            std::string m;
            int hrc = check_head(c->text, ss_trim_tail((const char *)c->text, fill), f.path, m);
            if (hrc) { release(c); fail(hrc, m); break; }
            first = false;
        }
        size_t cut = cut_chunk(c, fill, last);
        ss_chunk *c2 = nullptr;
        if (!last) {
            if (cut == 0) { release(c); fail(SS_ERR_FORMAT, f.path + ": no FASTQ record boundary within 64 KiB"); break; }
            c2 = acquire();
            if (!c2) { release(c); break; }
            // carry the tail (the bytes after the cut) and keep 32 KiB of history in front of the point
            // where decoding continues
            size_t tail = fill - cut, h = std::min(fill, tail + (size_t)SS_INGEST_HIST);
            memcpy(c2->text + tail - h, c->text + fill - h, h);
            fill = tail;
        }
        // gzip streams cannot be range-split: every rank decodes the whole stream and keeps its chunks
        if ((chunk_idx + (uint64_t)j.file) % (uint64_t)n_shards_ == (uint64_t)shard_) { c->len = cut; emit(c); }
        else release(c);
        chunk_idx++;
        if (last) break;
        c = c2;
    }
    delete g;
}
