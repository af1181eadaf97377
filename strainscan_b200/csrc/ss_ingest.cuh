// ss_ingest.cuh -- host ingest of read files: a pool of producer threads turns plain / gzip'ed FASTQ
// files into record-aligned text chunks in pinned host buffers, ready for cudaMemcpyAsync.
//
// Replaces the read-file argv and the `zcat a b |` pipe of library/identify.py:81-87 and
// library/Vote_Strain_L2_Lasso_new_sp.py:357-372.  Counts are additive over reads, so chunks may be
// delivered in any order (files and file parts are produced concurrently).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define SS_INGEST_HIST 32768u            // deflate history kept in front of every chunk's text area
#define SS_INGEST_BOUNDARY (64u << 10)   // a FASTQ record boundary is searched in the last 64 KiB of a chunk

struct ss_chunk {
    uint8_t *base = nullptr;   // pinned allocation: [SS_INGEST_HIST history][cap text bytes][slack]
    uint8_t *text = nullptr;   // whole FASTQ records, starts at a record start, ends with '\n'
    size_t len = 0;
    size_t cap = 0;
};

// FASTQ framing helpers shared with ss_api.cu
size_t ss_trim_tail(const char *buf, size_t len);
size_t ss_find_record_start(const char *buf, size_t len, size_t from);

class ss_text_source {
public:
    ss_text_source() = default;
    ~ss_text_source();
    ss_text_source(const ss_text_source &) = delete;
    ss_text_source &operator=(const ss_text_source &) = delete;

    // Allocate `n_buffers` pinned chunk buffers of `chunk_bytes` text capacity (kept for the life of the
    // object).  Needs a current CUDA device.
    // `pinned` = false uses plain host memory (host-only tools and tests; no CUDA context needed).
    int init(size_t chunk_bytes, int n_buffers, int n_threads, bool pinned = true);
    // Start producing shard `shard` of `n_shards` of the given files.  Plain files are split into
    // record-aligned byte ranges (one range per rank and producer); gzip streams are decoded whole by
    // every rank and their chunks are dealt round-robin.
    int start(const char *const *paths, int n_paths, int shard, int n_shards);
    ss_chunk *next();              // blocks; nullptr when everything was delivered or on error
    void release(ss_chunk *c);     // the buffer may be refilled (call once the H2D copy has completed)
    int finish();                  // joins the producers; returns the SS_ERR_* code of the first failure
    const std::string &error() const { return err_msg_; }
    size_t chunk_bytes() const { return chunk_bytes_; }
    uint64_t plain_bytes() const { return plain_bytes_; }   // bytes of plain text this shard will deliver
    uint64_t gz_bytes() const { return gz_bytes_; }         // compressed bytes of the gzip inputs
    bool ready() const { return !bufs_.empty(); }

private:
    struct file_map { int fd = -1; const uint8_t *map = nullptr; size_t size = 0; std::string path; bool gz = false; };
    struct job { int file = 0; size_t lo = 0, hi = 0; bool first_of_file = false; };

    void worker();
    void run_plain(const job &j);
    void run_gz(const job &j);
    ss_chunk *acquire();
    void emit(ss_chunk *c);
    void fail(int code, const std::string &msg);
    int plain_boundary(const file_map &f, size_t off, size_t *out);

    size_t chunk_bytes_ = 0;
    int n_threads_ = 1;
    bool pinned_ = true;
    std::vector<ss_chunk> bufs_;
    std::vector<file_map> files_;
    std::vector<job> jobs_;
    size_t next_job_ = 0;
    int active_ = 0;
    int shard_ = 0, n_shards_ = 1;
    uint64_t plain_bytes_ = 0, gz_bytes_ = 0;

    std::mutex mu_;
    std::condition_variable cv_free_, cv_ready_;
    std::deque<ss_chunk *> free_, ready_;
    std::vector<std::thread> threads_;
    int err_code_ = 0;
    std::string err_msg_;
    bool stop_ = false;
};
