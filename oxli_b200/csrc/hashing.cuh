// hashing.cuh -- device-side canonical k-mer hashing.
//
// Replaces, for the GPU, what the reference obtains from sourmash 0.23.0
// `SeqToHashes` + `_hash_murmur` (call sites /root/reference/src/lib.rs:69-76,
// 576-584): hash = h1 of MurmurHash3_x64_128(min(kmer, revcomp(kmer)) as
// upper-case ASCII, seed 42).  Integer-only; results are bit-exact.
#pragma once
#include <cstdint>

namespace oxg {

constexpr uint32_t kSeed = 42;
constexpr uint64_t kC1 = 0x87c37b91114253d5ULL;
constexpr uint64_t kC2 = 0x4cf5ad432745937fULL;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

__device__ __forceinline__ void mm_block(uint64_t &h1, uint64_t &h2, uint64_t k1, uint64_t k2) {
    k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
}

__device__ __forceinline__ uint64_t mm_finish(uint64_t h1, uint64_t h2, uint64_t len) {
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    return h1 + h2;
}

// Hash of the K bytes held little-endian in w[0..ceil(K/8)); bytes past K in the
// last word may hold anything.
template <int K>
__device__ __forceinline__ uint64_t murmur_words(const uint64_t *w) {
    constexpr int NB = K / 16, REM = K % 16;
    uint64_t h1 = kSeed, h2 = kSeed;
#pragma unroll
    for (int b = 0; b < NB; ++b) mm_block(h1, h2, w[2 * b], w[2 * b + 1]);
    if constexpr (REM > 8) {
        uint64_t k2 = w[2 * NB + 1] & (~0ULL >> (8 * (16 - REM)));
        k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; h2 ^= k2;
    }
    if constexpr (REM > 0) {
        uint64_t k1 = w[2 * NB];
        if constexpr (REM < 8) k1 &= ~0ULL >> (8 * (8 - REM));
        k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; h1 ^= k1;
    }
    return mm_finish(h1, h2, (uint64_t)K);
}

// Runtime-length variant reading bytes through `at(i)` (i in [0,len)).
template <class ByteAt>
__device__ __forceinline__ uint64_t murmur_bytes(ByteAt at, int len) {
    uint64_t h1 = kSeed, h2 = kSeed;
    const int nb = len / 16, rem = len % 16;
    for (int b = 0; b < nb; ++b) {
        uint64_t k1 = 0, k2 = 0;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            k1 = (k1 << 8) | at(16 * b + i);
            k2 = (k2 << 8) | at(16 * b + 8 + i);
        }
        mm_block(h1, h2, k1, k2);
    }
    uint64_t k1 = 0, k2 = 0;
    for (int i = rem; i > 8; --i) k2 = (k2 << 8) | at(16 * nb + i - 1);
    if (rem > 8) { k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; h2 ^= k2; }
    for (int i = rem < 8 ? rem : 8; i > 0; --i) k1 = (k1 << 8) | at(16 * nb + i - 1);
    if (rem > 0) { k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; h1 ^= k1; }
    return mm_finish(h1, h2, (uint64_t)len);
}

// ---- base classification, four bytes at a time -----------------------------

// 0x80 in every byte of x that is zero (exact, no cross-byte carries).
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

// Upper-case four ASCII bytes the way `to_ascii_uppercase` matters here: only
// A/C/G/T/a/c/g/t can become valid, and x & 0xDF maps exactly those onto ACGT.
__device__ __forceinline__ uint32_t upper4(uint32_t w) { return w & 0xdfdfdfdfu; }

// 4-bit mask, bit i set when byte i of u (already upper-cased) is one of ACGT.
__device__ __forceinline__ uint32_t acgt_mask4(uint32_t u) {
    uint32_t v = zero_bytes((u | 0x02020202u) ^ 0x43434343u)  // A or C
                 | zero_bytes(u ^ 0x47474747u)                // G
                 | zero_bytes(u ^ 0x54545454u);               // T
    return (((v >> 7) * 0x01020408u) >> 24) & 0xfu;
}

// Complement of four upper-case bases (garbage for non-ACGT bytes, which are
// never hashed): A<->T is ^0x15, C<->G is ^0x04; bit 1 tells the pairs apart.
__device__ __forceinline__ uint32_t complement4(uint32_t u) {
    uint32_t m = (u >> 1) & 0x01010101u;
    return u ^ 0x15151515u ^ (m * 0x11u);
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    return ((uint64_t)bswap32((uint32_t)x) << 32) | bswap32((uint32_t)(x >> 32));
}

// bytes [off, off+8) of the little-endian byte string held in x[0..n)
template <int OFF, int N>
__device__ __forceinline__ uint64_t word_at(const uint64_t (&x)[N]) {
    constexpr int A = OFF / 8, S = (OFF % 8) * 8;
    if constexpr (S == 0) {
        return x[A];
    } else {
        uint64_t lo = x[A] >> S;
        if constexpr (A + 1 < N) lo |= x[A + 1] << (64 - S);
        return lo;
    }
}

}  // namespace oxg
