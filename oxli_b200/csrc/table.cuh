// table.cuh -- open-addressing u64 -> u64 count table in HBM.
//
// Stands in for the reference's `counts: HashMap<u64,u64>`
// (/root/reference/src/lib.rs:33; entry/insert/get at 100-104, 178, 187, 679).
//
// Layout (split, chosen by measurement -- profiles/r1_microbench_soa_layout.txt):
//   keys[cap]  u64   probed four at a time: a 32-byte sector is one home bucket
//   lo[cap]    u32   the hot part of the count, bumped with RED.ADD.U32
//   hi[cap]    u64   the cold part; count = hi + lo
// `cap` is a power of two; home(key) = ((key * phi64) >> (64 - log2 cap)) & ~3, linear
// probing from there.  Compared with 16-byte {key,count} slots the probed array is half
// the size, twice as many candidates arrive per sector (home-bucket misses drop from 16 %
// to 6.5 % at load 0.6) and the array that takes the atomics (4 B per slot) stays in L2.
//
// Exactness of the split count: consume launches add 1 per k-mer with a RED that cannot
// see a 32-bit wrap, so the host keeps lo below 2^31 before such a launch (normalize
// kernel: moves bit 31 into hi) and never lets one launch add 2^31 to a slot; every other
// update goes through count_add(), which detects the wrap and repairs it.
//
// An empty slot holds kEmpty as key; that one key value is kept outside the arrays
// (Ctrl::side_*), so every u64 -- including 0 and 2^64-1 -- is a legal key and 0 a legal
// count.  No tombstones: single-key erase shifts the probe run back, bulk cuts rebuild.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace oxg {

constexpr uint64_t kEmpty = ~0ULL;
constexpr uint64_t kPhi = 0x9E3779B97F4A7C15ULL;
constexpr int kMaxProbe = 1024;
constexpr int kBucket = 4;  // keys per home bucket (one 32-byte sector)

struct Ctrl {              // lives in device memory, mirrored to pinned host memory
    // line 0: table state, read by every warp (size) -- kept apart from the hot atomics
    uint64_t size;         // live keys in the arrays
    uint64_t side_present; // key kEmpty is present
    uint64_t side_count;   // its count
    uint64_t first_bad;    // error-mode scan: smallest bad window start
    uint64_t pad0[12];
    // line 1: per-launch counters (zeroed together before every consume launch)
    uint64_t counted;      // k-mers counted by the running consume launch
    uint64_t overflow;     // entries appended to the overflow list
    uint64_t absorbed;     // received hashes counted by the running launch
    uint64_t pad1[13];
    // lines 2 and 3: the two work counters every warp hits with atomics
    uint64_t tile_counter;   // dynamic tile scheduler of the running consume launch
    uint64_t pad2[15];
    uint64_t absorb_counter; // same, for blocks of received hashes (sharded route launches)
    uint64_t pad3[15];
    uint64_t scratch[16];  // per-op outputs (stats, set sizes, ...)
};
static_assert(offsetof(Ctrl, counted) == 128 && offsetof(Ctrl, tile_counter) == 256 &&
              offsetof(Ctrl, absorb_counter) == 384 && offsetof(Ctrl, scratch) == 512, "Ctrl layout");

struct Slots {  // one allocation: keys | hi | lo
    uint64_t *keys;
    uint64_t *hi;
    uint32_t *lo;
};

struct TableView {
    Slots s;
    uint64_t cap;      // power of two, >= kBucket
    uint32_t shift;    // 64 - log2(cap)
    uint64_t limit;    // stop creating keys once size reaches this
    Ctrl *ctrl;
    uint64_t *overflow;   // deferred hashes (table too full), may be null
    uint64_t overflow_cap;
    // home slot: first of the four slots of a 32-byte key sector
    __device__ __forceinline__ uint64_t home(uint64_t key) const {
        return ((key * kPhi) >> shift) & ~(uint64_t)(kBucket - 1);
    }
};

// the four keys of a bucket with one 256-bit load (sm_100: LDG.E.256)
__device__ __forceinline__ void load_keys4(const uint64_t *p, uint64_t (&k)[kBucket]) {
    asm("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
}

// +1 without a return value.  Only legal while the host's headroom accounting holds
// (see the header comment).
__device__ __forceinline__ void count_red1(const TableView &t, uint64_t slot) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(t.s.lo + slot), "r"(1u) : "memory");
}

// count += inc, exact for any inc and any concurrency; returns the part of the old value
// that lived in lo (the caller adds hi when it needs the full count).
__device__ __forceinline__ uint32_t count_add(const TableView &t, uint64_t slot, uint64_t inc) {
    const uint32_t inc_lo = (uint32_t)(inc & 0x7fffffffu);
    const uint64_t inc_hi = inc - inc_lo;
    if (inc_hi) atomicAdd((unsigned long long *)(t.s.hi + slot), (unsigned long long)inc_hi);
    const uint32_t old = atomicAdd(t.s.lo + slot, inc_lo);
    if ((uint64_t)old + inc_lo > 0xffffffffull)  // lo wrapped: the lost 2^32 goes to hi
        atomicAdd((unsigned long long *)(t.s.hi + slot), 1ull << 32);
    return old;
}

__device__ __forceinline__ uint64_t count_value(const TableView &t, uint64_t slot) {
    return __ldcg(t.s.hi + slot) + __ldcg(t.s.lo + slot);
}

__device__ __forceinline__ void count_store(const TableView &t, uint64_t slot, uint64_t v) {
    const uint32_t lo = (uint32_t)(v & 0x7fffffffu);
    t.s.lo[slot] = lo;
    t.s.hi[slot] = v - lo;
}

__device__ __forceinline__ void push_overflow(const TableView &t, uint64_t key) {
    uint64_t at = atomicAdd((unsigned long long *)&t.ctrl->overflow, 1ULL);
    if (t.overflow && at < t.overflow_cap) t.overflow[at] = key;
}

// Slot of `key`, creating it if absent.  Returns the slot, or ~0 when the key was
// deferred to the overflow list (`full`: the table reached its load limit and the key is
// not present; or a deferral list exists and the probe run got absurdly long).
// *created is bumped when a new key was claimed.  Probing starts at bucket-aligned slot
// `i`, which the caller guarantees is not past the key's position.
__device__ __forceinline__ uint64_t table_slot_for(const TableView &t, uint64_t key, bool full, uint64_t i,
                                                   uint32_t *created) {
    for (int probe = 0; t.overflow == nullptr || probe < kMaxProbe; probe += kBucket) {
        uint64_t k[kBucket];
        load_keys4(t.s.keys + i, k);
#pragma unroll
        for (int q = 0; q < kBucket; ++q) {
            if (k[q] == key) return i + q;
            if (k[q] == kEmpty) {
                if (full) { push_overflow(t, key); return ~0ULL; }
                const uint64_t old = atomicCAS((unsigned long long *)(t.s.keys + i + q), kEmpty, key);
                if (old == kEmpty) { *created += 1; return i + q; }
                if (old == key) return i + q;
            }
        }
        i = (i + kBucket) & (t.cap - 1);
    }
    push_overflow(t, key);
    return ~0ULL;
}

// counts[key] += inc (exact path).  Returns 1 when a new key was created.
__device__ __forceinline__ uint32_t table_add(const TableView &t, uint64_t key, uint64_t inc, bool full) {
    if (key == kEmpty) {
        atomicAdd((unsigned long long *)&t.ctrl->side_count, (unsigned long long)inc);
        t.ctrl->side_present = 1;
        return 0;
    }
    uint32_t created = 0;
    const uint64_t slot = table_slot_for(t, key, full, t.home(key), &created);
    if (slot != ~0ULL) count_add(t, slot, inc);
    return created;
}

// +1 through the RED path (consume launches; headroom guaranteed by the host), probing
// from bucket-aligned slot `start`.
__device__ __forceinline__ uint32_t table_inc1_from(const TableView &t, uint64_t key, bool full, uint64_t start) {
    if (key == kEmpty) {
        atomicAdd((unsigned long long *)&t.ctrl->side_count, 1ULL);
        t.ctrl->side_present = 1;
        return 0;
    }
    uint32_t created = 0;
    const uint64_t slot = table_slot_for(t, key, full, start, &created);
    if (slot != ~0ULL) count_red1(t, slot);
    return created;
}

// counts[key] += inc, returning the count after the increment (count_hash,
// src/lib.rs:100-104); never defers: the caller reserved room.
__device__ __forceinline__ uint64_t table_add_fetch(const TableView &t, uint64_t key, uint64_t inc,
                                                    uint32_t *created) {
    if (key == kEmpty) {
        t.ctrl->side_present = 1;
        return atomicAdd((unsigned long long *)&t.ctrl->side_count, (unsigned long long)inc) + inc;
    }
    TableView u = t;
    u.overflow = nullptr;  // unbounded probing
    const uint64_t slot = table_slot_for(u, key, false, t.home(key), created);
    const uint64_t hi_before = __ldcg(t.s.hi + slot);
    const uint32_t old_lo = count_add(t, slot, inc);
    return hi_before + old_lo + inc;
}

// slot index of key, or -1
__device__ __forceinline__ int64_t table_find(const TableView &t, uint64_t key) {
    uint64_t i = t.home(key);
    for (uint64_t probe = 0; probe < t.cap; probe += kBucket) {
        uint64_t k[kBucket];
        load_keys4(t.s.keys + i, k);
#pragma unroll
        for (int q = 0; q < kBucket; ++q) {
            if (k[q] == key) return (int64_t)(i + q);
            if (k[q] == kEmpty) return -1;
        }
        i = (i + kBucket) & (t.cap - 1);
    }
    return -1;
}

__device__ __forceinline__ bool table_contains(const TableView &t, uint64_t key) {
    if (key == kEmpty) return t.ctrl->side_present != 0;
    return table_find(t, key) >= 0;
}

__device__ __forceinline__ uint64_t table_get(const TableView &t, uint64_t key) {
    if (key == kEmpty) return t.ctrl->side_present ? t.ctrl->side_count : 0;
    int64_t i = table_find(t, key);
    return i < 0 ? 0 : count_value(t, (uint64_t)i);
}

}  // namespace oxg
