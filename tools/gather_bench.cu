// tools/gather_bench.cu -- microbenchmark behind DESIGN.md's "random 32-byte sector" roofline:
// how many DRAM bytes does one random 32-byte probe cost on B200, per load flavour and per
// cudaLimitMaxL2FetchGranularity setting?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__host__ __device__ inline uint64_t mix(uint64_t x) {
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; return x;
}
struct __align__(32) S32 { unsigned long long k[4]; };

template <int MODE> __device__ __forceinline__ unsigned long long ld32(const S32 *p) {
    unsigned long long a, b, c, d;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 2) {  // two 16-byte loads
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"((const char *)p + 16));
    }
    if (MODE == 3) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 4) { // 8-byte load only (one slot)
        asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p)); b = c = d = 0;
    }
    return a ^ b ^ c ^ d;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather(const S32 *buf, uint64_t n_sectors, uint64_t n_probes, uint64_t seed, unsigned long long *sink) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i + 3 * stride < n_probes; i += 4 * stride) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = ld32<MODE>(buf + __umul64hi(mix(seed + i + u * stride), n_sectors));
#pragma unroll
        for (int u = 0; u < 4; u++) acc += v[u];
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <int MODE> float run(const S32 *buf, uint64_t ns, uint64_t np, unsigned long long *sink, int ctas) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        cudaEventRecord(a);
        gather<MODE><<<148 * ctas, 256>>>(buf, ns, np, 77 + it, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    size_t gran = argc > 1 ? atoi(argv[1]) : 0;
    uint64_t bytes = (argc > 2 ? atoll(argv[2]) : 1400) * 1000000ull;
    if (argc > 4) {   // sweep mode: L2-resident random access, 8-byte loads, buffer sizes in MB from argv[4..]
        unsigned long long *sink; cudaMalloc(&sink, 8);
        for (int a = 4; a < argc; a++) {
            uint64_t b = atoll(argv[a]) * 1000000ull, ns = b / 32; S32 *buf; cudaMalloc(&buf, ns * 32); cudaMemset(buf, 0x5a, ns * 32);
            float m4 = run<4>(buf, ns, 1ull << 28, sink, atoi(argv[3])), m0 = run<0>(buf, ns, 1ull << 28, sink, atoi(argv[3]));
            printf("sweep buf %6.0f MB: 8B loads %7.1f Gprobes/s   32B loads %7.1f Gprobes/s\n", b / 1e6, (1ull << 28) / m4 / 1e6, (1ull << 28) / m0 / 1e6);
            cudaFree(buf);
        }
        return 0;
    }
    int ctas = argc > 3 ? atoi(argv[3]) : 8;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set limit %zu -> %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    uint64_t ns = bytes / 32, np = 1ull << 28;
    S32 *buf; unsigned long long *sink;
    cudaMalloc(&buf, ns * 32); cudaMemset(buf, 0x5a, ns * 32); cudaMalloc(&sink, 8);
    const char *names[] = {"nc.no_allocate.v4.u64", "plain v4.u64", "2 x nc v2.u64", "evict_first v4.u64", "nc u64 (8 B)"};
    float ms[5] = {run<0>(buf, ns, np, sink, ctas), run<1>(buf, ns, np, sink, ctas), run<2>(buf, ns, np, sink, ctas),
                   run<3>(buf, ns, np, sink, ctas), run<4>(buf, ns, np, sink, ctas)};
    for (int m = 0; m < 5; m++)
        printf("buf %.2f GB ctas/SM %d  %-24s %8.3f ms  %7.1f Gprobes/s  %8.1f GB/s (32 B/probe)\n", bytes / 1e9, ctas, names[m], ms[m],
               np / ms[m] / 1e6, np * 32.0 / ms[m] / 1e6);
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
