// microbench_soa.cu -- would a split layout lift the DRAM-bound update rate?
// Layout A (current): 16-byte slots {key,count}, 2-slot home bucket = one 32-byte sector.
// Layout S: keys[cap] (8 B) probed 4 per 32-byte sector + cnt32[cap] updated with RED.32;
//           the counts array is a quarter of the AoS table and stays L2-resident.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__host__ __device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
constexpr uint64_t EMPTY = ~0ULL, PHI = 0x9E3779B97F4A7C15ULL;
__global__ void fill(uint64_t *p, uint64_t n, uint64_t v) { for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (uint64_t)blockDim.x) p[i] = v; }
__global__ void build(uint64_t *keys, uint64_t cap, uint32_t shift, uint64_t D) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < D; i += gridDim.x * (uint64_t)blockDim.x) {
        uint64_t k = mix(i) | 1, j = ((k * PHI) >> shift) & ~3ULL;
        for (;;) { uint64_t old = atomicCAS((unsigned long long *)&keys[j], EMPTY, k); if (old == EMPTY || old == k) break; j = (j + 1) & (cap - 1); }
    }
}
template <int U>
__global__ void __launch_bounds__(256) count_soa(const uint64_t *keys, uint32_t *cnt, uint64_t cap, uint32_t shift, uint64_t n, uint64_t D, uint64_t *misses) {
    uint64_t miss = 0;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint64_t k[U], idx[U], q[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) { k[u] = mix(mix(base + u * stride) % D) | 1; idx[u] = ((k[u] * PHI) >> shift) & ~3ULL; }
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q[u][0]), "=l"(q[u][1]), "=l"(q[u][2]), "=l"(q[u][3]) : "l"(keys + idx[u]));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int hit = -1;
#pragma unroll
            for (int s = 0; s < 4; ++s) if (q[u][s] == k[u]) hit = s;
            if (hit >= 0) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(cnt + idx[u] + hit), "r"(1u) : "memory"); continue; }
            ++miss;
            uint64_t j = (idx[u] + 4) & (cap - 1);
            for (;;) {
                uint64_t r[4];
                asm("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(keys + j));
                int h2 = -1;
#pragma unroll
                for (int s = 0; s < 4; ++s) if (r[s] == k[u]) h2 = s;
                if (h2 >= 0) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(cnt + j + h2), "r"(1u) : "memory"); break; }
                j = (j + 4) & (cap - 1);
            }
        }
    }
    for (int o = 16; o; o >>= 1) miss += __shfl_xor_sync(~0u, miss, o);
    if ((threadIdx.x & 31) == 0 && miss) atomicAdd((unsigned long long *)misses, (unsigned long long)miss);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint64_t *misses; cudaMalloc(&misses, 8);
    const uint64_t n = 1ull << 29;
    for (uint64_t D : {600000ull, 5000000ull, 40000000ull}) {
        uint64_t cap = 1024; while (cap < D + D / 2) cap <<= 1;
        uint32_t l = 0; while ((1ull << l) < cap) ++l; uint32_t shift = 64 - l;
        uint64_t *keys; uint32_t *cnt; cudaMalloc(&keys, cap * 8); cudaMalloc(&cnt, cap * 4);
        fill<<<sms * 8, 256>>>(keys, cap, EMPTY); cudaMemset(cnt, 0, cap * 4);
        build<<<sms * 8, 256>>>(keys, cap, shift, D); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int U : {4, 8}) {
            float best = 1e9f; uint64_t m = 0;
            for (int it = 0; it < 3; ++it) {
                cudaMemset(misses, 0, 8); cudaEventRecord(e0);
                if (U == 4) count_soa<4><<<sms * 8, 256>>>(keys, cnt, cap, shift, n, D, misses);
                else count_soa<8><<<sms * 8, 256>>>(keys, cnt, cap, shift, n, D, misses);
                cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                cudaMemcpy(&m, misses, 8, cudaMemcpyDeviceToHost);
            }
            printf("SoA D=%llu cap=%llu (keys %.0f MiB + cnt32 %.0f MiB, load %.2f) U=%d: %7.2f G keys/s, home-bucket misses %.1f%%\n",
                   (unsigned long long)D, (unsigned long long)cap, cap * 8.0 / (1 << 20), cap * 4.0 / (1 << 20), (double)D / cap, U, n / best / 1e6, 100.0 * m / n);
        }
        cudaFree(keys); cudaFree(cnt);
    }
}
