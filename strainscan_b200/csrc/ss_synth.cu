// ss_synth.cu -- device-side synthetic database / read generators (bench + full-size parity tests).
#include "ss_common.cuh"
#include "ss_kernels.cuh"
#include "ss_synth.cuh"

// one thread per FASTA record ">1\nKMER\n"
__global__ void ss_synth_db_kernel(ss_synth_params p, ss_synth_db_plan plan, uint8_t *__restrict__ text,
                                   uint32_t *__restrict__ node_of_record) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plan.n_records) return;
    ss_synth_write_db_record(p, plan, i, text, node_of_record);
}

// one thread per FASTQ record
__global__ void ss_synth_reads_kernel(ss_synth_params p, uint8_t *__restrict__ text, unsigned long long n_reads,
                                      unsigned long long first_read) {
    unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned long long rec = (unsigned long long)p.header_len + 1 + p.read_len + 1 + 2 + p.read_len + 1;
    ss_synth_write_read(p, first_read + r, text + r * rec);
}

cudaError_t ss_launch_synth_db(const ss_synth_params &p, const ss_synth_db_plan &plan, uint8_t *text,
                               uint32_t *node_of_record, cudaStream_t st) {
    if (plan.n_records == 0) return cudaSuccess;
    ss_synth_db_kernel<<<(unsigned)((plan.n_records + 255) / 256), 256, 0, st>>>(p, plan, text, node_of_record);
    return cudaGetLastError();
}

cudaError_t ss_launch_synth_reads_impl(const ss_synth_params &p, uint8_t *text, uint64_t n_reads,
                                       uint64_t first_read, cudaStream_t st) {
    if (n_reads == 0) return cudaSuccess;
    ss_synth_reads_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, st>>>(p, text, n_reads, first_read);
    return cudaGetLastError();
}
