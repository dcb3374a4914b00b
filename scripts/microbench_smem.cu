// microbench_smem.cu -- what one aggregation step costs on the SM's shared-memory path, and
// what a scattered 8/32/128-byte store costs, so that the partitioned pipeline's budget can
// be written down from measurements instead of guesses.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_smem microbench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

constexpr int kSlots = 8192;

// mode 0: ATOMS.ADD u32 no return, random slot      mode 1: same with return
// mode 2: LDS.128 random + ATOMS.ADD               mode 3: LDS.128 random only
// mode 4: tag write-and-check (STS lane, LDS back, LDS+STS count), no atomics
// mode 5: ATOMS.ADD all lanes same address         mode 6: ATOMS.CAS.64 random
template <int MODE>
__global__ void __launch_bounds__(512, 2) smem_kernel(uint64_t iters, uint64_t *out) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(sm);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sm + kSlots * 8);
    for (int i = threadIdx.x; i < kSlots; i += blockDim.x) { keys[i] = i; cnt[i] = 0; }
    __syncthreads();
    uint64_t x = mix(blockIdx.x * 1024ull + threadIdx.x), acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t s = (uint32_t)(x >> 40) & (kSlots - 1);
        if (MODE == 0) atomicAdd(&cnt[s], 1u);
        if (MODE == 1) acc += atomicAdd(&cnt[s], 1u);
        if (MODE == 2 || MODE == 3) {
            const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(keys + (s & ~1u));
            acc += kk.x ^ kk.y;
            if (MODE == 2) atomicAdd(&cnt[s], 1u);
        }
        if (MODE == 4) {
            volatile uint32_t *tag = reinterpret_cast<volatile uint32_t *>(keys);
            tag[s] = threadIdx.x;
            __syncwarp();
            if (tag[s] == threadIdx.x) cnt[s] = cnt[s] + 1;
            __syncwarp();
        }
        if (MODE == 5) atomicAdd(&cnt[0], 1u);
        if (MODE == 6) acc += atomicCAS((unsigned long long *)&keys[s], ~0ULL, (unsigned long long)x);
    }
    if (acc == 0x1234567) out[0] = acc + cnt[threadIdx.x];
}

// scattered stores: each thread appends to one of n_front fronts chosen at random; width bytes
template <int WIDTH>
__global__ void __launch_bounds__(256) scatter_kernel(uint8_t *buf, uint32_t n_front, uint32_t front_bytes, uint64_t iters) {
    uint64_t x = mix(blockIdx.x * 1024ull + threadIdx.x);
    uint32_t round = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t f = (uint32_t)((x >> 33) % n_front);
        // position inside the front advances with the iteration so a front's sectors fill in order
        const uint32_t pos = (uint32_t)((it * WIDTH + (threadIdx.x & 3) * 0) % front_bytes) & ~(WIDTH - 1);
        uint8_t *p = buf + (uint64_t)f * front_bytes + pos;
        if (WIDTH == 8) *reinterpret_cast<uint64_t *>(p) = x;
        if (WIDTH == 16) *reinterpret_cast<uint4 *>(p) = make_uint4((uint32_t)x, 1, 2, 3);
        if (WIDTH == 32) asm volatile("st.global.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(p), "l"(x) : "memory");
        ++round;
    }
}

// per-thread fronts: thread owns its fronts (like per-CTA fragments): 8-byte stores walk each front sequentially
__global__ void __launch_bounds__(256) frag_kernel(uint64_t *buf, uint32_t n_dest, uint32_t cap, uint64_t iters) {
    extern __shared__ uint32_t fill[];
    for (uint32_t i = threadIdx.x; i < n_dest; i += blockDim.x) fill[i] = 0;
    __syncthreads();
    uint64_t x = mix(blockIdx.x * 1024ull + threadIdx.x);
    uint64_t *mine = buf + (uint64_t)blockIdx.x * cap;
    const uint64_t row = (uint64_t)gridDim.x * cap;
    for (uint64_t it = 0; it < iters; ++it) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t d = (uint32_t)((x >> 33) % n_dest);
        const uint32_t pos = atomicAdd(&fill[d], 1u) % cap;
        mine[d * row + pos] = x;
    }
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t *out; CK(cudaMalloc(&out, 64));
    const size_t smem = kSlots * 12;
    const uint64_t iters = 4096;
    const double ops = (double)sms * 2 * 512 * iters;
#define RUN(M, name) { CK(cudaFuncSetAttribute(smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        float ms = timeit([&] { smem_kernel<M><<<sms * 2, 512, smem>>>(iters, out); }); \
        printf("%-44s %8.3f ms  %8.1f G ops/s  %6.2f cyc/op/SM @1.965GHz\n", name, ms, ops / ms / 1e6, ms * 1e-3 * 1.965e9 * sms / ops); }
    RUN(0, "ATOMS.ADD u32 random, no return");
    RUN(1, "ATOMS.ADD u32 random, return");
    RUN(2, "LDS.128 random + ATOMS.ADD");
    RUN(3, "LDS.128 random only");
    RUN(4, "tag write-and-check + plain increment");
    RUN(5, "ATOMS.ADD same address");
    RUN(6, "ATOMS.CAS.64 random");

    uint8_t *buf; const uint64_t bytes = 2ull << 30; CK(cudaMalloc(&buf, bytes));
    const uint64_t sit = 2048; const double sops = (double)sms * 8 * 256 * sit;
    for (uint32_t n_front : {1u << 12, 1u << 16, 1u << 18, 1u << 20, 1u << 21}) {
        const uint32_t fb = (uint32_t)(bytes / n_front) & ~127u;
        float m8 = timeit([&] { scatter_kernel<8><<<sms * 8, 256>>>(buf, n_front, fb, sit); });
        float m16 = timeit([&] { scatter_kernel<16><<<sms * 8, 256>>>(buf, n_front, fb, sit); });
        float m32 = timeit([&] { scatter_kernel<32><<<sms * 8, 256>>>(buf, n_front, fb, sit); });
        printf("scatter to %8u fronts: 8B %7.1f G st/s (%6.1f GB/s)  16B %7.1f G st/s  32B %7.1f G st/s (%6.1f GB/s)\n", n_front,
               sops / m8 / 1e6, sops * 8 / m8 / 1e6, sops / m16 / 1e6, sops / m32 / 1e6, sops * 32 / m32 / 1e6);
    }
    for (uint32_t n_dest : {512u, 1024u, 2048u, 4096u}) {
        const int grid = sms * 3;
        const uint32_t cap = (uint32_t)((bytes / 8) / ((uint64_t)n_dest * grid)) & ~3u;
        const uint64_t fit = 2048; const double fops = (double)grid * 256 * fit;
        float ms = timeit([&] { frag_kernel<<<grid, 256, n_dest * 4>>>((uint64_t *)buf, n_dest, cap, fit); });
        printf("per-CTA fragments, %4u dests x %d CTAs (cap %u): %7.1f G st/s  %6.2f cyc/op/SM\n", n_dest, grid, cap, fops / ms / 1e6, ms * 1e-3 * 1.965e9 * sms / fops);
    }
    return 0;
}
