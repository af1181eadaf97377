/*
 * strainscan_b200.h -- C ABI of the B200-native match+count engine for StrainScan's
 * identification hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference has no FFI for this path: it shells out to a bundled third-party binary and
 * re-parses its text output.  Every entry point below replaces one piece of that process/file
 * protocol; the reference lines it stands in for are cited per function
 * (paths relative to the StrainScan tree):
 *
 *   library/identify.py:73-103                      jellyfish_count()      (L1, k = 31)
 *   library/identify_low_mem.py:67-90               jellyfish_count()      (L1, -e DBs)
 *   library/identify_low_depth.py:46-74             jellyfish_count()      (-b 1)
 *   library/Vote_Strain_L2_Lasso_new_sp.py:354-403  vote_strain_L2() count block (L2, k = -k)
 *   library/identify.py:115-127                     match_node()           (per-node gather)
 *   library/identify_strains_L2_Enet_Pscan_new_sp.py:33-49,94-134   per-strain hit statistics
 *
 * Semantics are those of `jellyfish count -m K --if F reads... ; jellyfish dump -c` WITHOUT -C
 * (forward strand only, case folded, any non-ACGT byte breaks the window, zero-count records
 * stay valid), see SURVEY.md Appendix A.3 and DESIGN.md.
 *
 * All functions return SS_OK (0) or an SS_ERR_* code; ss_last_error() gives the message of the
 * calling thread's last failure.  There is NO CPU fallback: ss_init fails with SS_ERR_NO_DEVICE
 * when no sm_100 GPU is usable.  Calls block.  A context and the k-mer sets / read caches made from it carry
 * per-pass state (statistics, slot counters, scratch vectors, streaming slots): the count / load / reduce entry
 * points serialise on a per-context lock, so sharing a context between threads is safe but not concurrent --
 * use one context per thread (or per stream, ss_set_stream) to run passes side by side.  No temp files are
 * written (the reference writes temp_<uuid>.jf/.fa into cwd).
 */
#ifndef STRAINSCAN_B200_H
#define STRAINSCAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_OK               0
#define SS_ERR_NO_DEVICE    1   /* no CUDA device of compute capability 10.x */
#define SS_ERR_CUDA         2   /* a CUDA runtime call or kernel failed */
#define SS_ERR_IO           3   /* file open/read/inflate failure */
#define SS_ERR_FORMAT       4   /* input is not FASTQ / FASTA, or its record framing breaks */
#define SS_ERR_ARG          5   /* bad argument (NULL, k out of 1..32, ...) */
#define SS_ERR_NOMEM        6
#define SS_ERR_UNSUPPORTED  7

/* per-record flag bits returned by ss_kmerset_flags() */
#define SS_REC_IN_SET     1u    /* upper(record) has length k and is all ACGT: a key of the dump */
#define SS_REC_IS_LAST    2u    /* no later record has the same upper-cased string (identify.py:94) */
#define SS_REC_RAW_UPPER  4u    /* the raw record string is itself a dumped key (Vote_...:315) */

typedef struct ss_ctx     ss_ctx;      /* one GPU: streams, staging buffers, scratch */
typedef struct ss_kmerset ss_kmerset;  /* GPU-resident probe table seeded from a k-mer FASTA */
typedef struct ss_reads   ss_reads;    /* GPU-resident FASTQ text (the read cache) */
typedef struct ss_node_index    ss_node_index;     /* GPU-resident CSR node -> record ordinals (kmers/<node>) */
typedef struct ss_strain_matrix ss_strain_matrix;  /* GPU-resident CSC strain -> rows (all_strains_re.npz) */

typedef struct ss_stats {
    uint64_t text_bytes;      /* FASTQ bytes scanned */
    uint64_t n_reads;         /* sequence lines seen */
    uint64_t n_kmers;         /* valid k-windows probed (the unit of work) */
    uint64_t n_hits;          /* probes that found their k-mer */
    uint64_t n_second_probe;  /* probes that needed a second 32-byte sector */
    double   ms_index;        /* device time: line-index kernel K1 (0 when the read cache was indexed by an earlier pass) */
    double   ms_probe;        /* device time: fused scan/encode/probe/count kernel */
    double   ms_gather;       /* device time: slot -> record-ordinal gather */
    double   ms_h2d;          /* device time of host->device copies issued by this call */
    double   ms_total;        /* wall time of the call */
    uint32_t probe_launches;  /* launches of the probe kernel inside this call */
    uint32_t total_launches;  /* all kernel launches inside this call */
    uint64_t n_table_probes;  /* k-mers that passed the L2-resident prefilter and probed the HBM table
                                 (0 when the set is small enough to need no filter) */
    uint32_t binned_rounds;   /* rounds of the binned mode (large table + many table probes: survivors are
                                 grouped by table range first, then probed range after range out of L2) */
    uint32_t bins;            /* table ranges used by those rounds (0 = direct probing throughout) */
} ss_stats;

/* ---- context ------------------------------------------------------------------------------ */

/* Bind to CUDA device `device`.  Replaces locating/spawning library/jellyfish-linux
 * (identify.py:77, Vote_...:354). */
int ss_init(int device, ss_ctx **ctx);
int ss_shutdown(ss_ctx *ctx);
const char *ss_last_error(void);
/* Run the kernels on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) so the
 * caller can order its own work (NCCL all-reduce, CUDA events) after them AND the next pass after that work;
 * NULL = the context's own non-blocking stream.  To name the default stream pass cudaStreamLegacy
 * ((cudaStream_t)0x1), not 0 -- strainscan_b200.Engine.follow_torch_stream() does. */
int ss_set_stream(ss_ctx *ctx, void *stream);
/* name (<= name_cap bytes), SM count, total memory */
int ss_device_info(const ss_ctx *ctx, char *name, size_t name_cap, int *n_sm, uint64_t *mem_bytes);

/* ---- k-mer set: replaces `--if <fasta>` seeding + the adapters' FASTA re-read --------------- */

/* Load a k-mer FASTA (Tree_database/kmer.fa, Kmer_Sets_L2/Kmer_Sets/C<id>/all_kmer.fasta;
 * records ">id\nKMER\n", record ordinal = line pair index as in identify.py:92-95) and build the
 * open-addressing probe table in HBM.  1 <= k <= 32.  Replaces `--if` (identify.py:82,86;
 * Vote_...:359-371) and kmer_index_dict (identify.py:90-95). */
int ss_kmerset_from_fasta(ss_ctx *ctx, const char *path, int k, ss_kmerset **set);
int ss_kmerset_from_text(ss_ctx *ctx, const char *text, size_t len, int k, ss_kmerset **set);
/* Same as ss_kmerset_from_fasta with a binary cache of the parsed records (packed 2-bit k-mers, validity flags,
 * header ids: 9 bytes per record instead of ~36 of text) at `cache_path`: used when it matches the FASTA's size,
 * mtime and k, (re)written otherwise (best effort).  The FASTA stays the source of truth.  *cache_hit (may be NULL)
 * tells which way it went.  Removes the text parse the adapters pay per call (identify.py:90-95 re-reads kmer.fa
 * every time). */
int ss_kmerset_from_fasta_cached(ss_ctx *ctx, const char *path, int k, const char *cache_path, int *cache_hit,
                                 ss_kmerset **set);
int ss_kmerset_free(ss_kmerset *set);
uint64_t ss_kmerset_records(const ss_kmerset *set);   /* int(len(lines)/2) */
uint64_t ss_kmerset_distinct(const ss_kmerset *set);  /* distinct valid k-mers (= dump lines) */
int ss_kmerset_k(const ss_kmerset *set);
uint64_t ss_kmerset_table_bytes(const ss_kmerset *set);
/* flags[n_records]: SS_REC_* bits.  valid_kmers of identify.py:410 = IN_SET & IS_LAST. */
int ss_kmerset_flags(const ss_kmerset *set, uint8_t *flags);
/* ids[n_records]: integer after '>' (all_kmer.fasta: kid, Build_kmer_sets_...:397-399); 0 if none */
int ss_kmerset_header_ids(const ss_kmerset *set, uint64_t *ids);

/* ---- reads: replaces the read-file argv / `zcat a b |` pipe --------------------------------- */

/* Read (and inflate, if the name ends in .gz -- identify.py:81) the files in argv order and keep
 * shard `shard` of `n_shards` (record-aligned byte ranges) resident in HBM.  Paired-end = two
 * paths, concatenated exactly as the reference does (identify.py:75-76, Vote_...:367,371). */
int ss_reads_from_files(ss_ctx *ctx, const char *const *paths, int n_paths, int shard, int n_shards,
                        ss_reads **reads);
/* Host-only helper (no GPU needed): the record-aligned byte range [lo, hi) that shard `shard` of
 * `n_shards` owns in FASTQ text `buf`.  A record start is a line opening with '@' whose line + 2
 * opens with '+' (a quality line may open with '@'; its line + 2 is a sequence line).  The ranges
 * of all shards partition [0, len).  Used by ss_reads_from_files / ss_count_files. */
int ss_fastq_shard_range(const char *buf, size_t len, int shard, int n_shards, size_t *lo, size_t *hi);
/* Host-only helper (no GPU needed): run the read-file ingest that feeds ss_reads_from_files /
 * ss_count_files -- producer threads, the library's own gzip inflate (replaces the `zcat a b |` of
 * identify.py:82), record-aligned chunking, sharding -- and write the delivered chunks back to back
 * into `out` (chunk order is arbitrary; every chunk holds whole records).  out may be NULL to size. */
int ss_ingest_files_host(const char *const *paths, int n_paths, int shard, int n_shards, size_t chunk_bytes,
                         int n_threads, char *out, size_t out_cap, size_t *out_len, uint32_t *n_chunks);
/* Host-only helper (no GPU needed): the pipeline that inflates ORDINARY gzip streams on the device (block starts found
 * for pieces of `piece_bytes` compressed bytes, each piece decoded with an unknown window into marker symbols, pieces
 * stitched by exact block-boundary match, windows resolved in order; strainscan_b200/csrc/ss_dgz.cuh) run on the CPU with
 * the same source, for tests against zlib.  Decodes the members that start in [first_member, stop_member_at) of the
 * compressed bytes.  stats4 (may be NULL): pieces found, pieces used, batches, members.  Replaces the `zcat` of
 * identify.py:82 like ss_ingest_files_host does. */
int ss_dgz_inflate_host(const char *comp, size_t comp_size, size_t first_member, size_t stop_member_at,
                        uint32_t max_pieces, uint32_t piece_bytes, uint32_t sym_per_byte, char *out, size_t out_cap,
                        size_t *out_len, size_t *stopped_at, uint64_t *stats4);
/* Host-only helper (no GPU needed): index lists of the non-zeros of a dense C-contiguous n0 x n1 array, on n_threads host
 * threads (< 1: all the process may use).  The mirrors of the per-strain reducers (l2_shim: cal_cov_all, get_candidate_arr,
 * get_remainc -- identify_strains_L2_Enet_Pscan_new_sp.py:44-49, 94-134) take the dense 0/1 matrices the reference builds
 * (pX = X.A, :200-201) and need them as column lists for ss_strain_matrix_create.  elem_bytes-byte integers / bool
 * (is_float = 0) or IEEE floats (is_float = 1, elem_bytes 4 or 8).  axis 0: n0 lists of axis-1 indices; axis 1: n1 lists
 * of axis-0 indices, ascending.  Two calls: with idx == NULL, ptr[k + 1] receives the length of list k; the caller turns
 * ptr into offsets (ptr[0] = 0, running sum) and calls again with idx of ptr[last] entries. */
int ss_dense_nonzero_lists(const void *X, uint64_t n0, uint64_t n1, int elem_bytes, int is_float, int axis,
                           int n_threads, uint64_t *ptr, uint32_t *idx);
/* Host-only helper (no GPU needed): the k-mer lists of n tree nodes (Tree_database/kmers/<node>, one line of record
 * ordinals separated by single spaces; identify.py:115-119 reads it with list(map(int, line.rstrip().split(" "))) and
 * makes a set of it) parsed and de-duplicated on n_threads host threads (< 1: all the process may use).  ptr[n + 1] =
 * offsets into out, every list ascending; status[i]: 0 parsed, 1 empty file (the reference's "lines == []" case),
 * 2 not of that exact shape (signs, other separators, values over 2^32 - 1: left to the caller's slow parser, which has
 * the reference's semantics), 3 unreadable.  out_cap: half the total size of the files is always enough. */
int ss_node_lists_parse(const char *const *paths, uint32_t n, int n_threads, uint64_t *ptr, uint32_t *out,
                        uint64_t out_cap, uint8_t *status);
/* Host-only self-test (no GPU needed): the 16-bit Huffman decode tables of the device gzip decoder (ss_dgz2.cuh) against
 * the 32-bit tables of the host decoder (ss_inflate.cuh) on `trials` random prefix codes -- literal/length codes of
 * 257..286 symbols and distance codes of 1..30 symbols with lengths up to 15, the single-code and the empty distance
 * code, over-subscribed and incomplete codes; every 15-bit pattern must decode alike.  *n_checked = patterns compared. */
int ss_dgz_tables_selftest_host(uint64_t seed, uint32_t trials, uint64_t *n_checked);
/* Host-only helper (no GPU needed): how the device inflate would cut up a gzip file (range) of `compressed_bytes` on a
 * GPU of `n_sm` SMs whose head inflates to `head_ratio` text bytes per compressed byte.  out6 = piece bytes, pieces per
 * batch at most, symbols of room per compressed byte of a piece, text bytes per batch, device bytes the plan takes,
 * decoders in flight (one wave).  The kernel shape and the SS_DGZ_* overrides are read from the environment as in a
 * real call (strainscan_b200/csrc/ss_dgz_host.h: ss_dgz_make_plan). */
int ss_dgz_plan_host(size_t compressed_bytes, int n_sm, double head_ratio, uint64_t *out6);
/* Same from in-memory FASTQ text (each buffer = one file's uncompressed contents). */
int ss_reads_from_host(ss_ctx *ctx, const char *const *bufs, const size_t *lens, int n_bufs,
                       ss_reads **reads);
/* Borrow a device buffer that already holds FASTQ text of whole records (`len` bytes).  It must
 * be 256-byte aligned with capacity >= ss_reads_device_capacity(len); the tail is overwritten with
 * '\n' padding.  The caller keeps ownership (e.g. a torch uint8 tensor). */
int ss_reads_from_device(ss_ctx *ctx, void *dev_ptr, size_t len, size_t capacity, ss_reads **reads);
size_t ss_reads_device_capacity(size_t len);
uint64_t ss_reads_bytes(const ss_reads *reads);
int ss_reads_free(ss_reads *reads);
/* The line index of a read cache (line number mod 4 at every 992-byte unit, K1) is built by the first
 * counting pass over it -- inside that pass, reported as ss_stats.ms_index -- and reused by the later
 * passes (one per identified cluster, Vote_...:295-296).  This forgets it, so that the next pass pays
 * for it again: what a one-shot L1 pass costs (used by bench.py for `value`). */
int ss_reads_drop_index(ss_reads *reads);

/* ---- match + count: replaces `jellyfish count` + `jellyfish dump -c` + the dump parse -------- */

/* counts[i] (i = record ordinal, n_records entries) = number of forward-strand occurrences of
 * upper(record i) in the reads, 0 for records that are not keys.  HOST output.
 * Replaces identify.py:82-101 minus the dict (see strainscan_b200/identify_shim.py). */
int ss_count(ss_ctx *ctx, const ss_kmerset *set, const ss_reads *reads, uint32_t *counts, ss_stats *stats);
/* Same, DEVICE output (n_records uint32), left on the context's stream and synchronised before
 * return; used for the multi-GPU sum (NCCL all-reduce by the caller) and by the reducers. */
int ss_count_device(ss_ctx *ctx, const ss_kmerset *set, const ss_reads *reads, uint32_t *dev_counts,
                    ss_stats *stats);
/* End-to-end from HOST text: stages the buffers to the GPU in chunks (pinned double buffering,
 * copy overlapped with the kernels), counts, copies the dense vector back.  `counts` may be a host
 * or a device pointer (unified addressing decides). */
int ss_count_host(ss_ctx *ctx, const ss_kmerset *set, const char *const *bufs, const size_t *lens,
                  int n_bufs, uint32_t *counts, ss_stats *stats);
/* End-to-end from files (plain or .gz), shard `shard` of `n_shards`. */
int ss_count_files(ss_ctx *ctx, const ss_kmerset *set, const char *const *paths, int n_paths,
                   int shard, int n_shards, uint32_t *counts, ss_stats *stats);

/* L2 adapter on a dense DEVICE vector (after any cross-GPU sum): rows whose raw record is not a
 * dumped key -> 0, count == 1 -> 0 (remove_1, Vote_...:312-322), rows ordered by kid.
 * Row order, one rule everywhere (also l2_shim.kid_row_order): rows follow the FASTA header ids when these are an
 * exact permutation of 1..n (all_kmer.fasta: ">kid", Build_kmer_sets_...:397-399); any other header layout keeps
 * FASTA record order.  py_o[n_records] int64 HOST output. */
int ss_l2_finalize(ss_ctx *ctx, const ss_kmerset *set, const uint32_t *dev_counts, int64_t *py_o);

/* ---- reducers over a dense DEVICE count vector ------------------------------------------------ */

/* Per-node hit statistics, match_node() for many nodes at once (identify.py:115-127):
 * CSR node -> record ordinals (node_ptr[n_nodes+1], ordinals[node_ptr[n_nodes]], HOST arrays).
 * Per node: length = #distinct valid ordinals, covered = #valid ordinals with count > 0,
 * sum = their count total (before the 100x-median outlier trim, which stays on the host). */
int ss_node_reduce(ss_ctx *ctx, const ss_kmerset *set, const uint32_t *dev_counts,
                   const uint64_t *node_ptr, const uint32_t *ordinals, uint32_t n_nodes,
                   uint32_t *length, uint32_t *covered, uint64_t *sum);

/* Per-strain hit statistics (stat_cov / get_remainc / get_candidate_arr,
 * identify_strains_L2_Enet_Pscan_new_sp.py:33-49,94-134) on the CSC form of all_strains_re.npz:
 * col_ptr[n_strains+1], rows[nnz] (HOST), y = py_o (HOST int64, n_rows), row_mask (HOST uint8,
 * n_rows, may be NULL = all rows).  Per strain: total = #masked rows with X=1,
 * covered = #masked rows with X=1 and y > 1, sum = sum of y over covered rows. */
int ss_strain_reduce(ss_ctx *ctx, const uint64_t *col_ptr, const uint32_t *rows, uint32_t n_strains,
                     const int64_t *y, const uint8_t *row_mask, uint64_t n_rows,
                     uint64_t *total, uint64_t *covered, uint64_t *sum);

/* Persistent forms of the two reducers: the index structure is validated and uploaded ONCE and every reduction
 * moves only its inputs and its per-node / per-strain outputs.  identify_low_depth.identify_ranks visits every
 * node twice (identify_low_depth.py:113-132) and Pre_Scan asks for up to ~45 per-strain reductions per cluster
 * (identify_strains...:241-334), all over the same kmers/<node> lists resp. the same all_strains_re.npz.
 * ss_node_index_reduce also returns max_count[n_nodes] (may be NULL): with max < 100 the 100x-median outlier trim of
 * del_outlier (identify.py:106-112) cannot remove anything, so coverage = covered / length needs no host gather.
 * ss_strain_matrix_reduce: y (int64[n_rows]) and row_mask (uint8[n_rows] or NULL) may be HOST or DEVICE pointers. */
int ss_node_index_create(ss_ctx *ctx, const uint64_t *node_ptr, const uint32_t *ordinals, uint32_t n_nodes,
                         ss_node_index **index);
int ss_node_index_free(ss_node_index *index);
int ss_node_index_reduce(ss_ctx *ctx, const ss_kmerset *set, const ss_node_index *index, const uint32_t *dev_counts,
                         uint32_t *length, uint32_t *covered, uint64_t *sum, uint32_t *max_count);
int ss_strain_matrix_create(ss_ctx *ctx, const uint64_t *col_ptr, const uint32_t *rows, uint32_t n_strains,
                            uint64_t n_rows, ss_strain_matrix **matrix);
int ss_strain_matrix_free(ss_strain_matrix *matrix);
int ss_strain_matrix_reduce(ss_ctx *ctx, ss_strain_matrix *matrix, const int64_t *y, const uint8_t *row_mask,
                            uint64_t *total, uint64_t *covered, uint64_t *sum);

/* ---- measurement + synthetic workload helpers (bench.py, full-size parity tests) ---------------- */

/* K0: uniform-random 32-byte sector gathers over a `bytes`-sized buffer; returns useful GB/s
 * (best of `iters`).  The denominator of the probe kernel's random-access roofline. */
int ss_bench_random_gather(ss_ctx *ctx, uint64_t bytes, uint64_t n_probes, int iters, double *gbps);

/* Counter-based synthetic pan-genome (strainscan_b200/csrc/ss_synth.cuh): a heap-ordered binary
 * cluster search tree over n_leaves clusters; every genome block belongs to one ancestor of the
 * leaf, so its k-mers are specific to that node, as Build_tree.py's node k-mer sets are. */
#define SS_SYNTH_MAX_SOURCES 8
typedef struct ss_synth_params {
    uint64_t seed;
    uint32_t n_leaves;       /* clusters */
    uint32_t genome_len;     /* bases per cluster genome */
    uint32_t block_len;      /* ownership granularity */
    uint32_t k;
    uint32_t read_len;
    uint32_t header_len;     /* bytes of the '@' header line without its newline */
    uint32_t n_sources;      /* strains the reads are drawn from (<= SS_SYNTH_MAX_SOURCES) */
    uint32_t source_leaf[SS_SYNTH_MAX_SOURCES];    /* cluster index 0..n_leaves-1 */
    uint32_t source_strain[SS_SYNTH_MAX_SOURCES];  /* strain index inside the cluster */
    uint32_t source_cum[SS_SYNTH_MAX_SOURCES];     /* cumulative abundance scaled to 2^32-1 */
    uint32_t p_offtarget;    /* P(read is random sequence outside the DB) x 2^32 */
    uint32_t p_sub;          /* per-base substitution probability x 2^32 */
    uint32_t p_n;            /* per-base N probability x 2^32 */
    uint32_t snp_rate;       /* per-base strain SNP probability x 2^32 */
} ss_synth_params;

/* bytes of one synthetic FASTQ record / one synthetic k-mer FASTA record */
size_t ss_synth_read_record_bytes(const ss_synth_params *p);
size_t ss_synth_db_record_bytes(const ss_synth_params *p);
/* Write reads [first_read, first_read + n_reads) as 4-line FASTQ into DEVICE memory. */
int ss_synth_reads_device(ss_ctx *ctx, const ss_synth_params *p, void *dev_text, uint64_t n_reads,
                          uint64_t first_read);
/* Write the synthetic Tree_database/kmer.fa (both strands as separate records, shuffled by an affine
 * permutation) into HOST memory: node_sizes[n_nodes] records per node (even numbers; n_nodes =
 * 2*n_leaves-1).  text_out must hold sum(node_sizes) * ss_synth_db_record_bytes(); node_of_record
 * (may be NULL) receives the owner node of every record, i.e. the content of kmers/<node>. */
int ss_synth_db_host(ss_ctx *ctx, const ss_synth_params *p, const uint32_t *node_sizes, uint32_t n_nodes,
                     char *text_out, uint32_t *node_of_record);

#ifdef __cplusplus
}
#endif
#endif
