// q8_kernel.cu -- standalone GPU microbenchmark of the compact slot layout sketched in q8_model.c.
// NOT product code; compiled and run by hand:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/q8_kernel scripts/proto_q8/q8_kernel.cu
//   /tmp/q8_kernel [distinct_keys=5000000] [log2_slots=23] [n_hashes=2^29]
// Counts a list of n hashes drawn from `distinct` keys (plus four very hot ones) into
//   (a) the compact table: 8-byte slots, 4-slot buckets, spill table for full buckets / promoted keys,
//   (b) a 16-byte-slot table with 2-slot home buckets and linear probing (the current layout, simplified),
// checks every key's count against the host, and prints G keys/s for both.
// (Written at the end of round 1 without GPU minutes left: it compiles, it has not been run.)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr uint64_t PHI = 0x9E3779B97F4A7C15ULL;
constexpr uint64_t EMPTY16 = ~0ULL;

__host__ __device__ inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
// the i-th hash of the stream: one in 64 is one of four hot keys, the rest uniform over `distinct`
__host__ __device__ inline uint64_t stream_key(uint64_t i, uint64_t distinct) {
    const uint64_t x = mix(i);
    const uint64_t id = (x & 63) == 0 ? (x >> 6 & 3) : (x >> 6) % distinct;
    return mix(id ^ 0xabcdef) | 1;  // never 0, never 2^64-1's neighbour issues: odd keys only
}

__device__ __forceinline__ void red64(uint64_t *p, uint64_t v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld256(const uint64_t *p, uint64_t (&w)[4]) {
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
}

// ---- exact 16-byte-slot table (spill table of (a), whole table of (b)) ----
struct T16 { ulonglong2 *slots; uint64_t cap; int shift; };
__device__ __forceinline__ void t16_add(const T16 &t, uint64_t key, uint64_t inc) {
    uint64_t i = ((key * PHI) >> t.shift) & ~1ULL;
    for (;;) {
        uint64_t w[4];
        ld256(reinterpret_cast<const uint64_t *>(t.slots + i), w);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint64_t k = w[2 * q];
            if (k == EMPTY16) {
                const uint64_t old = atomicCAS((unsigned long long *)&t.slots[i + q].x, EMPTY16, key);
                k = old == EMPTY16 ? key : old;
            }
            if (k == key) { red64(reinterpret_cast<uint64_t *>(&t.slots[i + q].y), inc); return; }
        }
        i = (i + 2) & (t.cap - 1);
    }
}
__device__ __forceinline__ uint64_t t16_get(const T16 &t, uint64_t key) {
    uint64_t i = ((key * PHI) >> t.shift) & ~1ULL;
    for (;;) {
        for (int q = 0; q < 2; ++q) {
            const ulonglong2 s = t.slots[i + q];
            if (s.x == key) return s.y;
            if (s.x == EMPTY16) return 0;
        }
        i = (i + 2) & (t.cap - 1);
    }
}

// ---- compact table ----
struct Q8 { uint64_t *slots; int c, rbits; uint64_t rmask, one; T16 spill; };
__device__ __forceinline__ void q8_resolve(const Q8 &t, uint64_t key, uint64_t bucket, uint64_t r, uint64_t (&w)[4]) {
    // after the bucket's four words are in registers
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (w[q] != 0 && (w[q] & t.rmask) == r) {
            if (w[q] >> 63) t16_add(t.spill, key, 1);           // promoted: exact counter elsewhere
            else red64(t.slots + bucket + q, t.one);             // blind add into the count field
            return;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (w[q] == 0) {
            const uint64_t old = atomicCAS((unsigned long long *)(t.slots + bucket + q), 0ULL, r + t.one);
            if (old == 0) return;                                // claimed with count 1
            if ((old & t.rmask) == r) {                          // someone else just created this key here
                if (old >> 63) t16_add(t.spill, key, 1); else red64(t.slots + bucket + q, t.one);
                return;
            }
            w[q] = old;                                          // taken by another key: keep looking
        }
    }
    t16_add(t.spill, key, 1);                                    // bucket full of other keys
}
template <int U>
__global__ void __launch_bounds__(256) q8_count(Q8 t, uint64_t n, uint64_t distinct) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint64_t key[U], bucket[U], r[U], w[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + u * stride;
            key[u] = i < n ? stream_key(i, distinct) : 0;
            const uint64_t kp = key[u] * PHI;
            bucket[u] = (kp >> t.rbits) * 4; r[u] = kp & t.rmask;
            ld256(t.slots + bucket[u], w[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * stride < n) q8_resolve(t, key[u], bucket[u], r[u], w[u]);
    }
}
__global__ void q8_get(Q8 t, uint64_t distinct, uint64_t *out) {
    const uint64_t id = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (id >= distinct) return;
    const uint64_t key = mix(id ^ 0xabcdef) | 1, kp = key * PHI;
    const uint64_t *b = t.slots + (kp >> t.rbits) * 4, r = kp & t.rmask;
    uint64_t total = 0; bool found = false, full = true, plain = false;
    for (int q = 0; q < 4; ++q) {
        if (b[q] == 0) { full = false; continue; }
        if ((b[q] & t.rmask) == r) { total = b[q] >> t.rbits; found = true; plain = !(b[q] >> 63); }
    }
    if (!plain && (found || full)) total += t16_get(t.spill, key);
    out[id] = total;
}
template <int U>
__global__ void __launch_bounds__(256) t16_count(T16 t, uint64_t n, uint64_t distinct) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += stride * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + u * stride;
            if (i < n) t16_add(t, stream_key(i, distinct), 1);
        }
    }
}
__global__ void t16_get_all(T16 t, uint64_t distinct, uint64_t *out) {
    const uint64_t id = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (id < distinct) out[id] = t16_get(t, mix(id ^ 0xabcdef) | 1);
}
__global__ void fill16(ulonglong2 *s, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (uint64_t)blockDim.x) s[i] = make_ulonglong2(EMPTY16, 0);
}

static T16 make16(uint64_t cap) {
    T16 t; t.cap = cap; t.shift = 64 - (int)__builtin_ctzll(cap);
    CK(cudaMalloc(&t.slots, cap * 16));
    fill16<<<1024, 256>>>(t.slots, cap);
    return t;
}

int main(int argc, char **argv) {
    const uint64_t distinct = argc > 1 ? strtoull(argv[1], nullptr, 10) : 5000000;
    const int c = argc > 2 ? atoi(argv[2]) : 23;
    const uint64_t n = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1ull << 29;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 3;   // 444 CTAs x 256 threads x 4 keys in flight: the bound q8_model.c argues with
    // expected counts on the host
    std::vector<uint64_t> want(distinct, 0);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t x = mix(i);
        want[(x & 63) == 0 ? (x >> 6 & 3) : (x >> 6) % distinct]++;
    }
    uint64_t *d_out; CK(cudaMalloc(&d_out, distinct * 8));
    std::vector<uint64_t> got(distinct);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto check = [&](const char *name, float ms) {
        CK(cudaMemcpy(got.data(), d_out, distinct * 8, cudaMemcpyDeviceToHost));
        uint64_t bad = 0;
        for (uint64_t i = 0; i < distinct; ++i) bad += got[i] != want[i];
        printf("%-44s %7.2f ms  %6.2f G keys/s  %s\n", name, ms, n / ms / 1e6, bad ? "MISMATCH" : "exact");
    };
    {   // (a) compact
        Q8 t; t.c = c; t.rbits = 66 - c; t.rmask = (1ULL << t.rbits) - 1; t.one = 1ULL << t.rbits;
        CK(cudaMalloc(&t.slots, (8ull << c))); CK(cudaMemset(t.slots, 0, 8ull << c));
        t.spill = make16(1ull << (c - 2));
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {   // second pass: every key present (the steady state of the bench)
            CK(cudaEventRecord(e0));
            q8_count<4><<<grid, 256>>>(t, n, distinct);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("compact pass %d: %.2f ms\n", rep, ms);
        }
        for (auto &w : want) w *= 2;
        q8_get<<<(unsigned)((distinct + 255) / 256), 256>>>(t, distinct, d_out);
        char name[96]; snprintf(name, sizeof name, "compact 8-byte slots, %llu MiB + spill", (unsigned long long)((8ull << c) >> 20));
        check(name, ms);
        for (auto &w : want) w /= 2;
        CK(cudaFree(t.slots)); CK(cudaFree(t.spill.slots));
    }
    {   // (b) today's layout
        T16 t = make16(1ull << c);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            t16_count<4><<<grid, 256>>>(t, n, distinct);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("16-byte pass %d: %.2f ms\n", rep, ms);
        }
        for (auto &w : want) w *= 2;
        t16_get_all<<<(unsigned)((distinct + 255) / 256), 256>>>(t, distinct, d_out);
        char name[96]; snprintf(name, sizeof name, "16-byte slots, %llu MiB", (unsigned long long)((16ull << c) >> 20));
        check(name, ms);
        CK(cudaFree(t.slots));
    }
    return 0;
}
