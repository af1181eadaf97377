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
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define SS_INGEST_HIST 32768u            // deflate history kept in front of every chunk's text area
#define SS_INGEST_BOUNDARY (64u << 10)   // a FASTQ record boundary is searched in the last 64 KiB of a chunk first (ss_find_cut)

// One independently inflatable gzip member (BGZF block) of a batch; the device kernel's work item.
struct ss_member {
    uint32_t comp_off;    // raw deflate stream inside the batch's compressed bytes
    uint32_t comp_len;
    uint32_t out_off;     // where its text goes, relative to the start of the batch's text
    uint32_t isize;       // inflated size (ISIZE of the trailer)
};
#define SS_BGZF_MAX_MEMBERS 16384u

// A chunk is either plain TEXT (whole FASTQ records in `text[0, len)`) or a BGZF batch, whose text is
//   [pre_text][members inflated on the DEVICE][post_text]      (pre_len + inflated_len + post_len bytes)
// with the compressed members in text[0, len), the two host-decoded text pieces behind them (the
// boundary member of every batch is inflated by the producer so that batches begin and end on record
// boundaries) and the member table in `members`.
enum { SS_CHUNK_TEXT = 0, SS_CHUNK_BGZF = 1, SS_CHUNK_RAW = 2 };
struct ss_chunk {
    uint8_t *base = nullptr;   // pinned allocation: [SS_INGEST_HIST history][cap text bytes][slack][member table]
    uint8_t *text = nullptr;   // TEXT: whole FASTQ records ending with '\n'; BGZF: compressed members
    size_t len = 0;
    size_t cap = 0;
    int kind = SS_CHUNK_TEXT;
    ss_member *members = nullptr;      // SS_BGZF_MAX_MEMBERS entries (pinned)
    uint32_t n_members = 0;
    uint64_t file_off = 0;             // RAW: bytes [file_off, file_off + len) of the file, verbatim (start_raw)
    const uint8_t *pre_text = nullptr, *post_text = nullptr;   // inside this buffer, behind the compressed bytes
    size_t pre_len = 0, post_len = 0, inflated_len = 0;
    size_t text_len() const { return kind == SS_CHUNK_TEXT ? len : pre_len + inflated_len + post_len; }
};

// FASTQ framing helpers shared with ss_api.cu
// whole file into memory; gzip'ed files (by suffix or magic) are inflated.  Returns SS_OK or an SS_ERR_* code.
int ss_read_whole_file(const char *path, std::vector<char> &out, std::string &err);
size_t ss_trim_tail(const char *buf, size_t len);
size_t ss_find_record_start(const char *buf, size_t len, size_t from);
size_t ss_find_cut(const char *buf, size_t lo, size_t hi);
// multi-member gzip files (ss_ingest.cu): is the file one (decided from the distance to its second member, the same on
// every rank), its fixed parts, and the first member start in [from, limit) (`size` = none)
bool ss_gz_is_member_split(const uint8_t *map, size_t size, int n_shards);
size_t ss_gz_split_parts(size_t size);
size_t ss_gz_next_member_start(const uint8_t *map, size_t size, size_t from, size_t limit);

class ss_text_source {
public:
    ss_text_source() = default;
    ~ss_text_source();
    ss_text_source(const ss_text_source &) = delete;
    ss_text_source &operator=(const ss_text_source &) = delete;

    // Allocate `n_buffers` pinned chunk buffers of `chunk_bytes` text capacity (kept for the life of the
    // object).  Needs a current CUDA device.
    // `pinned` = false uses plain host memory (host-only tools and tests; no CUDA context needed).
    // `device_bgzf`: deliver BGZF files as SS_CHUNK_BGZF batches for the device inflate kernel (at most
    // `bgzf_out_cap` text bytes per batch); otherwise they are decoded on the host like any gzip stream.
    int init(size_t chunk_bytes, int n_buffers, int n_threads, bool pinned = true, bool device_bgzf = false,
             size_t bgzf_out_cap = 0);
    // Start producing shard `shard` of `n_shards` of the given files.  Plain files are split into
    // record-aligned byte ranges (one range per rank and producer); multi-member gzip files are split at
    // member starts (fixed file parts, every member decoded by exactly one rank); a single gzip stream is
    // decoded whole by every rank and its chunks (always cut the same way) are dealt round-robin.
    int start(const char *const *paths, int n_paths, int shard, int n_shards);
    // Deliver bytes [lo, hi) of ONE file verbatim as SS_CHUNK_RAW chunks (any order, several readers): the upload of a
    // compressed file that the device inflates.
    int start_raw(const char *path, size_t lo, size_t hi);
    ss_chunk *next();              // blocks; nullptr when everything was delivered or on error
    void release(ss_chunk *c);     // the buffer may be refilled (call once the H2D copy has completed)
    int finish();                  // joins the producers; returns the SS_ERR_* code of the first failure
    const std::string &error() const { return err_msg_; }
    size_t chunk_bytes() const { return chunk_bytes_; }
    uint64_t plain_bytes() const { return plain_bytes_; }   // bytes of plain text this shard will deliver
    uint64_t gz_bytes() const { return gz_bytes_; }         // compressed bytes of the gzip inputs
    size_t bgzf_out_cap() const { return bgzf_out_cap_; }
    bool ready() const { return !bufs_.empty(); }

private:
    struct file_map { int fd = -1; const uint8_t *map = nullptr; size_t size = 0; std::string path; bool gz = false, bgzf = false, normalize = false, split = false; };
    struct job { int file = 0; size_t lo = 0, hi = 0; bool first_of_file = false; };

    void worker();
    void run_plain(const job &j);
    void run_gz(const job &j);
    void run_gz_members(const job &j);
    void run_gz_parallel(const job &j, int threads, size_t span);
    void run_bgzf(const job &j);
    void run_normalize(const job &j);
    void run_raw(const job &j);
    bool raw_mode_ = false;
    // in-order byte stream of one file -> record-aligned chunks (used by the parallel gzip decoder)
    struct stream_writer {
        ss_text_source *src = nullptr;
        const job *j = nullptr;
        ss_chunk *c = nullptr;
        size_t fill = 0;
        uint64_t chunk_idx = 0;
        bool first = true;
        int fill_threads = 1;
        // n bytes produced by fill(dst, off, len) (callable on disjoint slices from several threads);
        // false: stopped or failed (fail() was called)
        bool append(size_t n, const std::function<void(uint8_t *, size_t, size_t)> &fill);
        bool finish();
        void abandon();
    };
    ss_chunk *acquire();
    void emit(ss_chunk *c);
    void fail(int code, const std::string &msg);
    int plain_boundary(const file_map &f, size_t off, size_t *out);

    size_t chunk_bytes_ = 0;
    int n_threads_ = 1;
    bool pinned_ = true, device_bgzf_ = false;
    size_t bgzf_out_cap_ = 0;
    std::vector<ss_chunk> bufs_;
    std::vector<file_map> files_;
    std::vector<job> jobs_;
    size_t next_job_ = 0;
    int active_ = 0;
    int shard_ = 0, n_shards_ = 1;
    int n_gz_jobs_ = 0;
    uint64_t plain_bytes_ = 0, gz_bytes_ = 0;

    std::mutex mu_;
    std::condition_variable cv_free_, cv_ready_;
    std::deque<ss_chunk *> free_, ready_;
    std::vector<std::thread> threads_;
    int err_code_ = 0;
    std::string err_msg_;
    bool stop_ = false;
};
