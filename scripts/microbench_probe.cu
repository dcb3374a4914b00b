// microbench_probe.cu -- how fast can a real probe-and-count loop go?  D distinct
// 64-bit keys, each hit many times in random order, counted into an open-addressing
// table of 16-byte slots (two-slot home buckets, linear probing), at several load
// factors.  Variants isolate the cost of the slow path.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__host__ __device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
constexpr uint64_t EMPTY = ~0ULL, PHI = 0x9E3779B97F4A7C15ULL;
struct T { ulonglong2 *s; uint64_t cap; uint32_t shift; };
__device__ __forceinline__ uint64_t home(const T &t, uint64_t k) { return ((k * PHI) >> t.shift) & ~1ULL; }
__device__ __forceinline__ void ld2(const ulonglong2 *p, ulonglong2 &a, ulonglong2 &b) {
    asm("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.x), "=l"(a.y), "=l"(b.x), "=l"(b.y) : "l"(p));
}
__device__ __forceinline__ void red(unsigned long long *p) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(1ULL) : "memory"); }
__device__ __forceinline__ uint64_t key_of(uint64_t i, uint64_t D) { return mix(mix(i) % D) | 1; }

__global__ void init(T t) { for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < t.cap; i += gridDim.x * (uint64_t)blockDim.x) t.s[i] = make_ulonglong2(EMPTY, 0); }
__global__ void build(T t, uint64_t D) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < D; i += gridDim.x * (uint64_t)blockDim.x) {
        uint64_t k = mix(i) | 1, j = home(t, k);
        for (;;) {
            uint64_t old = atomicCAS((unsigned long long *)&t.s[j].x, EMPTY, k);
            if (old == EMPTY || old == k) break;
            j = (j + 1) & (t.cap - 1);
        }
    }
}
// MODE 0: fast path only (misses ignored); 1: per-lane bucket-wise slow path inline;
// 2: blind: RED on the home slot without looking (upper bound of LDG+RED);
// 3: slow path probes 2 buckets (64 B) per round
template <int MODE, int U>
__global__ void __launch_bounds__(256) count(T t, uint64_t n, uint64_t D, uint64_t *misses) {
    uint64_t miss = 0;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint64_t k[U], idx[U]; ulonglong2 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { k[u] = mix(mix(base + u * stride) % D) | 1; idx[u] = home(t, k[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u) ld2(t.s + idx[u], a[u], b[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (MODE == 2) { red(&t.s[idx[u]].y); miss += a[u].x ^ b[u].x; continue; }
            if (a[u].x == k[u]) red(&t.s[idx[u]].y);
            else if (b[u].x == k[u]) red(&t.s[idx[u] + 1].y);
            else if (MODE == 0) ++miss;
            else {
                ++miss;
                uint64_t j = (idx[u] + 2) & (t.cap - 1);
                for (;;) {
                    ulonglong2 c, d; ld2(t.s + j, c, d);
                    if (c.x == k[u]) { red(&t.s[j].y); break; }
                    if (d.x == k[u]) { red(&t.s[j + 1].y); break; }
                    if (MODE == 3) {
                        ulonglong2 e, f; ld2(t.s + ((j + 2) & (t.cap - 1)), e, f);
                        if (e.x == k[u]) { red(&t.s[(j + 2) & (t.cap - 1)].y); break; }
                        if (f.x == k[u]) { red(&t.s[((j + 2) & (t.cap - 1)) + 1].y); break; }
                        j = (j + 4) & (t.cap - 1);
                    } else j = (j + 2) & (t.cap - 1);
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) miss += __shfl_xor_sync(~0u, miss, o);
    if ((threadIdx.x & 31) == 0 && miss) atomicAdd((unsigned long long *)misses, (unsigned long long)miss);
}
template <int MODE, int U> void run(const char *name, T t, uint64_t n, uint64_t D, uint64_t *misses, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e9f; uint64_t m = 0;
    for (int it = 0; it < 3; ++it) {
        cudaMemset(misses, 0, 8); cudaEventRecord(e0);
        count<MODE, U><<<sms * 8, 256>>>(t, n, D, misses);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        cudaMemcpy(&m, misses, 8, cudaMemcpyDeviceToHost);
    }
    printf("    %-44s %7.2f G keys/s   home-bucket misses %.1f%%\n", name, n / best / 1e6, MODE == 2 ? 0.0 : 100.0 * m / n);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint64_t *misses; cudaMalloc(&misses, 8);
    const uint64_t n = 1ull << 29;
    for (uint64_t D : {600000ull, 5000000ull}) for (int lg : {0, 1, 2}) {
        uint64_t cap = 1024; while (cap < D + D / 2) cap <<= 1; cap <<= lg;
        T t; t.cap = cap; uint32_t l = 0; while ((1ull << l) < cap) ++l; t.shift = 64 - l;
        cudaMalloc(&t.s, cap * 16); init<<<sms * 8, 256>>>(t); build<<<sms * 8, 256>>>(t, D); cudaDeviceSynchronize();
        printf("D=%llu keys, %llu slots (%.0f MiB), load %.2f\n", (unsigned long long)D, (unsigned long long)cap, cap * 16.0 / (1 << 20), (double)D / cap);
        run<2, 8>("blind LDG.256 + RED (no compare)", t, n, D, misses, sms);
        run<0, 8>("compare, fast path only (misses dropped)", t, n, D, misses, sms);
        run<1, 8>("full: inline per-lane slow path", t, n, D, misses, sms);
        run<3, 8>("full: slow path reads 2 buckets per round", t, n, D, misses, sms);
        run<1, 4>("full, 4 keys in flight per thread", t, n, D, misses, sms);
        cudaFree(t.s);
    }
}
