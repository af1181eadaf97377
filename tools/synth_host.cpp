// synth_host.cpp -- host-side generator of the synthetic workload (bench / test tooling, NOT the product).
//
// Writes byte-for-byte the same kmer.fa and FASTQ text as the device generators of
// strainscan_b200/csrc/ss_synth.cu (both call the record writers of ss_synth.cuh), with host threads and
// without a GPU or the product library.  bench.py --impl reference uses it so that the reference arm never
// loads libstrainscan_b200.so; tests/test_gpu_parity.py checks the two generators against each other.
//
//   g++ -O3 -std=c++17 -shared -fPIC -pthread -o tools/libss_synth_host.so tools/synth_host.cpp
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../strainscan_b200/csrc/ss_synth.cuh"

static thread_local std::string g_err;

template <typename F>
static void parallel_for(uint64_t n, int n_threads, F f) {
    n_threads = std::max(1, n_threads);
    if (n < 4096) n_threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([=]() {
            const uint64_t lo = n * (uint64_t)t / (uint64_t)n_threads, hi = n * (uint64_t)(t + 1) / (uint64_t)n_threads;
            for (uint64_t i = lo; i < hi; i++) f(i);
        });
    for (auto &x : th) x.join();
}

extern "C" {

const char *ssh_last_error(void) { return g_err.c_str(); }

size_t ssh_read_record_bytes(const ss_synth_params *p) { return (size_t)p->header_len + 1 + p->read_len + 1 + 2 + p->read_len + 1; }
size_t ssh_db_record_bytes(const ss_synth_params *p) { return (size_t)p->k + 4; }

// reads [first_read, first_read + n_reads) as 4-line FASTQ into `out`
int ssh_reads(const ss_synth_params *p, uint8_t *out, uint64_t n_reads, uint64_t first_read, int n_threads) {
    if (!p || !out) { g_err = "ssh_reads: NULL argument"; return 1; }
    const ss_synth_params P = *p;
    const uint64_t rec = ssh_read_record_bytes(p);
    parallel_for(n_reads, n_threads, [=](uint64_t r) { ss_synth_write_read(P, first_read + r, out + r * rec); });
    return 0;
}

// the synthetic Tree_database/kmer.fa (sum(node_sizes) records) and, optionally, the owner node of every record
int ssh_db(const ss_synth_params *p, const uint32_t *node_sizes, uint32_t n_nodes, uint8_t *text_out,
           uint32_t *node_of_record, int n_threads) {
    if (!p || !node_sizes || !text_out) { g_err = "ssh_db: NULL argument"; return 1; }
    ss_synth_db_layout lay;
    std::string why = ss_synth_db_layout_build(p, node_sizes, n_nodes, lay);
    if (!why.empty()) { g_err = "ssh_db: " + why; return 1; }
    ss_synth_db_plan plan;
    plan.node_off = lay.node_off.data(); plan.n_nodes = n_nodes;
    plan.blk_list = lay.blk_list.data(); plan.blk_off = lay.blk_off.data();
    plan.n_records = lay.n_records; plan.perm_a = lay.perm_a; plan.perm_c = lay.perm_c;
    const ss_synth_params P = *p;
    parallel_for(lay.n_records, n_threads, [=](uint64_t i) { ss_synth_write_db_record(P, plan, i, text_out, node_of_record); });
    return 0;
}

}   // extern "C"
