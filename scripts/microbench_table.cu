// microbench_table.cu -- what bounds random count-table updates on a B200?
// Random 16-byte slots in a table of S bytes; per key: optional 256-bit (or
// 128-bit) sector load, optional 64/32-bit RED or ATOM.  Prints G ops/s.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_table microbench_table.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

enum { LD256 = 1, LD128 = 2, RED64 = 4, RED32 = 8, ATOM64 = 16, SAMEWARPLINE = 32 };

template <int OPS, int U>
__global__ void __launch_bounds__(256) k(ulonglong2 *slots, uint64_t mask, uint64_t n, uint64_t *sink) {
    uint64_t acc = 0;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint64_t idx[U];
        ulonglong2 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t key = base + u * stride;
            if (OPS & SAMEWARPLINE) key = (key >> 5);  // a warp hits one slot pair: coalesced reference
            idx[u] = mix(key) & mask & ~1ULL;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (OPS & LD256)
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u].x), "=l"(a[u].y), "=l"(b[u].x), "=l"(b[u].y) : "l"(slots + idx[u]));
            if (OPS & LD128) a[u] = __ldcg(slots + idx[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (OPS & LD256) acc += a[u].x ^ b[u].x;
            if (OPS & LD128) acc += a[u].x;
            if (OPS & RED64) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(&slots[idx[u]].y), "l"(1ULL) : "memory");
            if (OPS & RED32) asm volatile("red.global.add.u32 [%0], %1;" ::"l"(&slots[idx[u]].y), "r"(1u) : "memory");
            if (OPS & ATOM64) acc += atomicAdd((unsigned long long *)&slots[idx[u]].y, 1ULL);
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int OPS, int U>
void run(const char *name, ulonglong2 *slots, uint64_t nslots, uint64_t n, uint64_t *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        k<OPS, U><<<sms * 8, 256>>>(slots, nslots - 1, n, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    printf("  %-34s U=%d  %8.2f G keys/s\n", name, U, n / best / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint64_t *sink; cudaMalloc(&sink, 8);
    const uint64_t n = 1ull << 29;
    for (uint64_t mib : {4ull, 32ull, 128ull, 1024ull, 8192ull}) {
        const uint64_t nslots = mib * (1ull << 20) / 16;
        ulonglong2 *slots; cudaMalloc(&slots, nslots * 16); cudaMemset(slots, 0, nslots * 16);
        printf("table %llu MiB (%llu slots), %llu keys per run, %d SMs\n", (unsigned long long)mib, (unsigned long long)nslots, (unsigned long long)n, sms);
        run<LD256, 8>("LDG.256 only", slots, nslots, n, sink, sms);
        run<LD128, 8>("LDG.128 only", slots, nslots, n, sink, sms);
        run<RED64, 8>("RED.64 only", slots, nslots, n, sink, sms);
        run<RED32, 8>("RED.32 only", slots, nslots, n, sink, sms);
        run<ATOM64, 8>("ATOM.64 (returning) only", slots, nslots, n, sink, sms);
        run<LD256 | RED64, 8>("LDG.256 + RED.64", slots, nslots, n, sink, sms);
        run<LD256 | RED64, 4>("LDG.256 + RED.64", slots, nslots, n, sink, sms);
        run<LD256 | RED64, 16>("LDG.256 + RED.64", slots, nslots, n, sink, sms);
        run<LD128 | RED64, 8>("LDG.128 + RED.64", slots, nslots, n, sink, sms);
        run<LD256 | RED64 | SAMEWARPLINE, 8>("warp-uniform LDG.256 + RED.64", slots, nslots, n, sink, sms);
        cudaFree(slots);
    }
    return 0;
}
