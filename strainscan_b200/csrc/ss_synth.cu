// ss_synth.cu -- device-side synthetic database / read generators (bench + full-size parity tests).
#include "ss_common.cuh"
#include "ss_kernels.cuh"
#include "ss_synth.cuh"

// one thread per FASTA record ">1\nKMER\n"
__global__ void ss_synth_db_kernel(ss_synth_params p, ss_synth_db_plan plan, uint8_t *__restrict__ text,
                                   uint32_t *__restrict__ node_of_record) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plan.n_records) return;
    unsigned long long q = (plan.perm_a * i + plan.perm_c) % plan.n_records;
    // node = last v with node_off[v] <= q
    uint32_t lo = 0, hi = plan.n_nodes;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (plan.node_off[mid] <= q) lo = mid; else hi = mid;
    }
    uint32_t v = lo;
    unsigned long long j = q - plan.node_off[v];
    uint32_t strand = (uint32_t)(j & 1ull);
    uint32_t t = (uint32_t)(j >> 1);
    uint32_t d = ss_synth_depth(v);
    uint32_t nb = plan.blk_off[d + 1] - plan.blk_off[d];
    uint32_t blk = plan.blk_list[plan.blk_off[d] + (t % nb)];
    uint32_t pos = blk * p.block_len + (t / nb);
    const uint32_t rec = p.k + 4;
    uint8_t *o = text + i * rec;
    o[0] = '>'; o[1] = '1'; o[2] = '\n';
    for (uint32_t x = 0; x < p.k; x++) {
        uint32_t gp = strand ? pos + p.k - 1 - x : pos + x;
        uint32_t b = ss_synth_node_base(&p, v, gp);
        if (strand) b = 3u - b;
        o[3 + x] = "ACGT"[b];
    }
    o[3 + p.k] = '\n';
    if (node_of_record) node_of_record[i] = v;
}

// one thread per FASTQ record
__global__ void ss_synth_reads_kernel(ss_synth_params p, uint8_t *__restrict__ text, unsigned long long n_reads,
                                      unsigned long long first_read) {
    unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned long long rid = first_read + r;
    const uint32_t L = p.read_len, H = p.header_len;
    const unsigned long long rec = (unsigned long long)H + 1 + L + 1 + 2 + L + 1;
    uint8_t *o = text + r * rec;
    // header: '@' 'S' then zero-padded decimal read id
    o[0] = '@'; o[1] = 'S';
    {
        unsigned long long x = rid;
        for (int i = (int)H - 1; i >= 2; i--) { o[i] = '0' + (uint8_t)(x % 10); x /= 10; }
    }
    o[H] = '\n';
    uint64_t h = ss_h3(p.seed, 0x4EAD5ull, rid);
    bool off = (uint32_t)h < p.p_offtarget;
    uint32_t src = 0;
    uint64_t h2 = ss_h3(p.seed, 0x4EAD6ull, rid);
    if (!off) {
        uint32_t a = (uint32_t)(h >> 32);
        while (src + 1 < p.n_sources && a > p.source_cum[src]) src++;
    }
    uint32_t leaf = p.n_leaves - 1 + p.source_leaf[src];
    uint32_t strain = p.source_strain[src];
    uint32_t pos = (uint32_t)(h2 % (uint64_t)(p.genome_len - L + 1));
    uint32_t strand = (uint32_t)(h2 >> 63);
    uint8_t *s = o + H + 1;
    for (uint32_t i = 0; i < L; i++) {
        uint32_t b;
        if (off) {
            b = (uint32_t)(ss_h3(p.seed ^ 0x0FF7ull, rid, i) >> 9) & 3u;
        } else {
            uint32_t gp = strand ? pos + L - 1 - i : pos + i;
            b = ss_synth_strain_base(&p, leaf, strain, gp);
            if (strand) b = 3u - b;
        }
        uint64_t e = ss_h3(p.seed ^ 0xE440ull, rid, i);
        uint8_t c;
        if ((uint32_t)e < p.p_n) c = 'N';
        else {
            if ((uint32_t)(e >> 32) < p.p_sub) b = (b + 1u + (uint32_t)((e >> 20) % 3u)) & 3u;
            c = "ACGT"[b];
        }
        s[i] = c;
    }
    s[L] = '\n'; s[L + 1] = '+'; s[L + 2] = '\n';
    uint8_t *qv = s + L + 3;
    for (uint32_t i = 0; i < L; i++)   // Phred+33 range '!'..'J': includes '@' and '+' (also as first char)
        qv[i] = '!' + (uint8_t)((ss_h3(p.seed ^ 0x9A1ull, rid, i) >> 13) % 42u);
    qv[L] = '\n';
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
