// microbench_q8.cu -- how fast are random count updates when the table fits the L2?
// Same irreducible work as microbench_table.cu (one 32-byte sector load + one RED.64 per key),
// swept over table sizes from 32 to 160 MiB, to locate where the B200's L2 stops holding a
// randomly updated table: the curve a slot layout of 8 bytes (4 candidate slots per sector,
// 67 MB for the C2 workload) would move along instead of today's 16-byte slots at 134 MB.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench_q8 scripts/microbench_q8.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

// STREAM: also read `stream_bytes_per_key` bytes of a big buffer per key, like the reads that
// flow through L2 next to the table in the real kernel (1.26 B per k-mer).
template <int U, bool STREAM, bool RETURNING = false>
__global__ void __launch_bounds__(256) k(uint64_t *slots, uint64_t nbuckets, uint64_t n, const uint4 *stream,
                                        uint64_t stream_vecs, uint64_t *sink) {
    uint64_t acc = 0;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint64_t idx[U], a[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t h = mix(base + u * stride);
            idx[u] = __umul64hi(h, nbuckets) * 4 + (h & 3);  // slot inside a 4-slot, 32-byte bucket
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u][0]), "=l"(a[u][1]), "=l"(a[u][2]), "=l"(a[u][3]) : "l"(slots + (idx[u] & ~3ULL)));
        if (STREAM) {  // one 16-byte vector per 8 keys of this thread ~ 2 B per key
            const uint4 v = __ldcs(stream + (base / U) % stream_vecs);
            acc += v.x;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc += a[u][0] ^ a[u][1] ^ a[u][2] ^ a[u][3];
            if (RETURNING)  // what a compact slot needs to see its count field wrap: the old value comes back
                acc += atomicAdd((unsigned long long *)(slots + idx[u]), 1ULL << 44) >> 63;
            else
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(slots + idx[u]), "l"(1ULL) : "memory");
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int U, bool STREAM, bool RETURNING = false>
float run(uint64_t *slots, uint64_t nbuckets, uint64_t n, const uint4 *stream, uint64_t stream_vecs, uint64_t *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        k<U, STREAM, RETURNING><<<sms * 8, 256>>>(slots, nbuckets, n, stream, stream_vecs, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    return n / best / 1e6f;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint64_t *sink; cudaMalloc(&sink, 8);
    const uint64_t n = 1ull << 29;
    const uint64_t stream_bytes = 2ull << 30;
    uint4 *stream; cudaMalloc(&stream, stream_bytes); cudaMemset(stream, 1, stream_bytes);
    printf("random 32-byte bucket load + RED.64 per key, %llu keys per run, %d SMs, L2 %d MB\n",
           (unsigned long long)n, sms, p.l2CacheSize >> 20);
    printf("%10s %14s %14s %22s %22s\n", "table MiB", "U=4 G keys/s", "U=8 G keys/s", "U=4 + streamed reads", "U=4, returning ATOM");
    for (uint64_t mib : {16ull, 32ull, 48ull, 64ull, 72ull, 80ull, 96ull, 112ull, 128ull, 160ull, 256ull}) {
        const uint64_t nbuckets = mib * (1ull << 20) / 32;
        uint64_t *slots; cudaMalloc(&slots, nbuckets * 32); cudaMemset(slots, 0, nbuckets * 32);
        const float a = run<4, false>(slots, nbuckets, n, stream, stream_bytes / 16, sink, sms);
        const float b = run<8, false>(slots, nbuckets, n, stream, stream_bytes / 16, sink, sms);
        const float c = run<4, true>(slots, nbuckets, n, stream, stream_bytes / 16, sink, sms);
        const float d = run<4, false, true>(slots, nbuckets, n, stream, stream_bytes / 16, sink, sms);
        printf("%10llu %14.2f %14.2f %22.2f %22.2f\n", (unsigned long long)mib, a, b, c, d);
        cudaFree(slots);
    }
    return 0;
}
