// shard.cuh -- device side of the multi-GPU exchange protocol.
//
// The table is sharded by the high bits of the hash (owner(h) = h >> (64 - log2 N)); the
// reference's primitive for combining partial tables is KmerCountTable::add
// (/root/reference/src/lib.rs:778-837), here replaced by routing every hash to its owner
// before it is counted.  The exchange is a PULL: pass A (consume_kernel<K, kModePart>) leaves
// each rank's hashes in its OWN HBM, in fragments addressed by (owner, partition); pass B
// (aggregate_kernel) of the owner reads the fragments of all ranks through peer-mapped
// pointers, so the NVLink transfer is the aggregation kernel's own load stream.  What crosses
// the ranks besides those loads is a handful of 8-byte flags in each rank's exchange header,
// written by peers with system-scope release stores and awaited by one-thread kernels on the
// owner's stream -- no host in the loop, no collective library.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace oxg {

constexpr int kShardMaxRanks = 16;
constexpr uint32_t kGatherBytes = 1u << 20;  // per-source slot of the small all-gather

// At the start of every rank's exchange area.  Every field is written by exactly one rank.
struct ShardHeader {
    // ready[s]:    rounds rank s has published for me to read          (written by rank s)
    // consumed[d]: rounds of MINE rank d has finished reading           (written by rank d)
    // gathered[s]: all-gather sequence number of rank s's blob          (written by rank s)
    unsigned long long ready[kShardMaxRanks];
    unsigned long long consumed[kShardMaxRanks];
    unsigned long long gathered[kShardMaxRanks];
    unsigned long long gather_len[kShardMaxRanks];
    unsigned long long error;       // a wait gave up (peer gone): set locally
    unsigned long long pad[63];
};
static_assert(sizeof(ShardHeader) == 1024, "ShardHeader layout is part of the inter-rank protocol");

struct FlagList {
    unsigned long long *ptr[kShardMaxRanks];
    int n;
};

// *ptr[i] = value for every i, after everything this GPU wrote before is visible system-wide
static __global__ void shard_signal_kernel(FlagList f, unsigned long long value) {
    __threadfence_system();
    if ((int)threadIdx.x < f.n) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.ptr[threadIdx.x]), "l"(value) : "memory");
    }
}

// spin until *ptr[i] >= value for every i (flags only grow); gives up after timeout_ns
static __global__ void shard_wait_kernel(FlagList f, unsigned long long value, unsigned long long timeout_ns,
                                         unsigned long long *error) {
    if ((int)threadIdx.x >= f.n) return;
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f.ptr[threadIdx.x]) : "memory");
        if (v >= value) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) { atomicExch(error, 1ULL); break; }
        __nanosleep(200);
    }
}

// copy `bytes` (multiple of 16) from src to dst[i] for every i: the blob of the small all-gather
struct BlobDests {
    uint8_t *ptr[kShardMaxRanks];
    int n;
};
static __global__ void shard_scatter_blob_kernel(BlobDests d, const uint8_t *__restrict__ src, uint32_t bytes) {
    const uint32_t n16 = bytes / 16;
    for (int i = blockIdx.y; i < d.n; i += gridDim.y)
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n16; j += gridDim.x * blockDim.x)
            reinterpret_cast<uint4 *>(d.ptr[i])[j] = reinterpret_cast<const uint4 *>(src)[j];
}

}  // namespace oxg
