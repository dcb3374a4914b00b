// aggregate.cuh -- pass B of the partitioned counting pipeline.
//
// Together with consume_kernel<K, kModePart> (pass A) this replaces the per-window
// `count_hash` of KmerCountTable::consume (/root/reference/src/lib.rs:586-600, 100-104)
// by "scatter, pre-reduce duplicates, then update":
//
//   pass A   hashes the reads and scatters every hash into the fragment of its destination,
//            a destination being (owning rank, contiguous segment of that rank's table) --
//            the partition index is the top bits of h * phi, the same product whose top bits
//            are the slot index.
//   pass B   (this file) one CTA at a time takes a partition, streams its fragments -- from
//            every source rank: the pointers may be peer memory, then these loads are the
//            multi-GPU exchange -- through a shared-memory table that adds up duplicates
//            (key -> occurrences in this launch), and merges the distinct keys into the table
//            in HBM with one update each, in slot order.
//
// Why: a random 16-byte read-modify-write per k-mer is bound by L1TEX wavefronts and L2
// atomics at ~45 G k-mers/s on a hot table (profiles/r1_microbench_table_updates.txt).  After
// partitioning, all occurrences of a key inside a launch meet in one CTA: on high-coverage
// input (C2: every key ~10x per 64 M-window launch) nine of ten updates never leave shared
// memory, and what does reach the table walks one small segment in ascending slot order.
//
// The shared-memory table is a cache, not a container: a hash that finds its neighbourhood
// full (kLocalProbe buckets) bypasses it and updates the table in HBM directly, so a
// partition with more distinct keys than the cache holds (singleton-heavy input) degrades
// to the direct path without any mode switch, and skew is harmless: the spill list of pass A
// (a k-mer flooding its partition, e.g. poly-A) is cut into slices that any CTA aggregates.
#pragma once
#include "table.cuh"

namespace oxg {

constexpr int kAggThreads = 512;
constexpr int kAggWarps = kAggThreads / 32;
constexpr int kLocalBits = 13;
constexpr uint32_t kLocalSlots = 1u << kLocalBits;  // 8192 x (8-byte key + 4-byte count) = 96 KB
constexpr int kLocalProbe = 4;                      // buckets of two slots examined before bypassing
constexpr uint32_t kSpillSlice = 1u << 15;          // spill-list entries per work item
constexpr int kMaxSources = 16;

struct AggSource {
    const uint64_t *frag;                 // [n_dest_total][n_ctas][frag_cap]
    const uint32_t *frag_cnt;             // [n_dest_total][n_ctas]
    const uint64_t *spill;
    const unsigned long long *spill_n;
};

struct AggParams {
    TableView table;
    AggSource src[kMaxSources];
    int n_src;
    uint32_t n_parts;     // partitions of THIS rank's table
    uint32_t dest0;       // self_rank * n_parts: first destination index that is ours
    uint32_t n_ctas;      // grid of pass A (fragments per destination and source)
    uint32_t frag_cap;
    uint32_t part_bits;   // log2(n_parts)
    uint32_t groups;      // work items per partition: item (part, g) takes the fragments f with f % groups == g
    uint64_t spill_cap;
    int owner_shift, self_rank, n_ranks;  // owner(h) = h >> owner_shift, for the spill lists
    unsigned long long *work_counter;     // zeroed before the launch
};

inline size_t aggregate_smem_bytes() { return (size_t)kLocalSlots * 12; }

// One occurrence of h into the shared-memory table.  false = neighbourhood full, bypass.
// Slots only ever go from empty to a key, so a stale "empty" is caught by the CAS and a stale
// "other key" cannot happen; equal keys racing for a slot agree on it through the CAS result.
__device__ __forceinline__ bool local_count(uint64_t *lk, uint32_t *ld, uint64_t h, uint32_t idx) {
    uint32_t b = idx & ~1u;
#pragma unroll 1
    for (int pr = 0; pr < kLocalProbe; ++pr) {
        const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(lk + b);
        if (kk.x == h) { atomicAdd(ld + b, 1u); return true; }
        if (kk.y == h) { atomicAdd(ld + b + 1, 1u); return true; }
        if (kk.x == kEmpty) {
            const uint64_t old = atomicCAS((unsigned long long *)(lk + b), (unsigned long long)kEmpty, (unsigned long long)h);
            if (old == kEmpty || old == h) { atomicAdd(ld + b, 1u); return true; }
        }
        if (kk.y == kEmpty) {
            const uint64_t old = atomicCAS((unsigned long long *)(lk + b + 1), (unsigned long long)kEmpty, (unsigned long long)h);
            if (old == kEmpty || old == h) { atomicAdd(ld + b + 1, 1u); return true; }
        }
        b = (b + 2) & (kLocalSlots - 1);
    }
    return false;
}

__global__ void __launch_bounds__(kAggThreads, 2) aggregate_kernel(const AggParams p) {
    extern __shared__ __align__(16) uint8_t agg_smem[];
    uint64_t *lk = reinterpret_cast<uint64_t *>(agg_smem);                    // keys
    uint32_t *ld = reinterpret_cast<uint32_t *>(agg_smem + kLocalSlots * 8);  // occurrences
    __shared__ unsigned long long s_item;
    __shared__ uint64_t s_spill_first[kMaxSources + 1];  // work-item index of each source's first spill slice

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TableView &tv = p.table;

    if (threadIdx.x == 0) {
        uint64_t run = (uint64_t)p.n_parts * p.groups;
        for (int s = 0; s < p.n_src; ++s) {
            s_spill_first[s] = run;
            const uint64_t n = min((uint64_t)*p.src[s].spill_n, p.spill_cap);
            run += (n + kSpillSlice - 1) / kSpillSlice;
        }
        s_spill_first[p.n_src] = run;
    }
    __syncthreads();
    const uint64_t n_items = s_spill_first[p.n_src];

    uint32_t created = 0;
    uint64_t n_taken = 0;  // occurrences this thread fed into the table (cache or direct)
    auto flush_created = [&]() {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, created);
        if (lane == 0 && tot) atomicAdd((unsigned long long *)&tv.ctrl->size, (unsigned long long)tot);
        created = 0;
    };

    // U hashes per lane: through the shared-memory table first; what bypasses it goes to the
    // table in HBM with all home-bucket loads in flight together
    constexpr int U = 4;
    auto take = [&](const uint64_t (&h)[U], uint32_t live, bool full) {
        uint32_t direct = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!((live >> u) & 1u)) continue;
            const uint32_t idx = (uint32_t)(((h[u] * kPhi) << p.part_bits) >> (64 - kLocalBits));
            if (h[u] == kEmpty || !local_count(lk, ld, h[u], idx)) direct |= 1u << u;
        }
        if (direct) {  // two at a time: four sets of bucket registers do not fit 64 registers per thread
            const uint64_t one[2] = {1, 1};
            const uint64_t lo2[2] = {h[0], h[1]}, hi2[2] = {h[2], h[3]};
            if (direct & 3u) created += table_add_many<2>(tv, lo2, one, direct & 3u, full);
            if (direct >> 2) created += table_add_many<2>(tv, hi2, one, direct >> 2, full);
        }
    };
    // a contiguous run of n hashes, consumed by one warp
    auto take_run = [&](const uint64_t *ptr, uint32_t n, bool filter_owner, bool full) {
        for (uint32_t off = 0; off < n; off += 32 * U) {
            uint64_t h[U];
            uint32_t live = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = off + u * 32 + lane;
                h[u] = i < n ? __ldcs(ptr + i) : 0;
                bool ok = i < n && h[u] != 0;  // (0 pads the last line a launch of pass A wrote)
                if (filter_owner && p.n_ranks > 1 && ok) ok = (int)(h[u] >> p.owner_shift) == p.self_rank;
                live |= (ok ? 1u : 0u) << u;
            }
            n_taken += __popc(live);
            take(h, live, full);
        }
    };

    for (;;) {
        __syncthreads();  // previous item fully merged; s_item free
        if (threadIdx.x == 0) s_item = atomicAdd(p.work_counter, 1ULL);
        // empty the cache
        for (uint32_t i = threadIdx.x; i < kLocalSlots / 2; i += kAggThreads)
            reinterpret_cast<ulonglong2 *>(lk)[i] = make_ulonglong2(kEmpty, kEmpty);
        for (uint32_t i = threadIdx.x; i < kLocalSlots / 4; i += kAggThreads)
            reinterpret_cast<uint4 *>(ld)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const uint64_t item = s_item;
        if (item >= n_items) break;

        if (item < (uint64_t)p.n_parts * p.groups) {
            // a partition of this rank's table: its fragments from every source and pass-A CTA
            // (with groups > 1, one share of them: neighbouring CTAs then work on the same
            // segment of the table at the same time)
            const uint64_t dest = p.dest0 + item / p.groups;
            const uint32_t g = (uint32_t)(item % p.groups);
            const uint32_t n_frag = p.n_ctas * (uint32_t)p.n_src;
            const bool full = tv.overflow != nullptr && __ldcg(&tv.ctrl->size) >= tv.limit;
            // warp w takes the fragments g + (w + kAggWarps * i) * groups; their fill counts are
            // fetched 32 at a time, one per lane, so that a fragment costs no latency of its own
            for (uint32_t f0 = g + warp * p.groups; f0 < n_frag; f0 += 32 * kAggWarps * p.groups) {
                const uint32_t mine = f0 + lane * kAggWarps * p.groups;
                uint32_t my_n = 0;
                const uint64_t *my_ptr = nullptr;
                if (mine < n_frag) {
                    const int s = (int)(mine / p.n_ctas);
                    const uint64_t row = dest * p.n_ctas + (mine - (uint32_t)s * p.n_ctas);
                    my_n = min(__ldg(p.src[s].frag_cnt + row), p.frag_cap);
                    my_ptr = p.src[s].frag + row * p.frag_cap;
                }
                for (int l = 0; l < 32; ++l) {
                    const uint32_t n = __shfl_sync(0xffffffffu, my_n, l);
                    const uint64_t *ptr = reinterpret_cast<const uint64_t *>(__shfl_sync(0xffffffffu, (unsigned long long)my_ptr, l));
                    if (f0 + (uint32_t)l * kAggWarps * p.groups >= n_frag) break;
                    if (n) take_run(ptr, n, false, full);
                }
            }
            flush_created();
        } else {
            // a slice of some source's spill list (hashes of any partition, any owner)
            int s = 0;
            while (item >= s_spill_first[s + 1]) ++s;
            const uint64_t n_sp = min((uint64_t)*p.src[s].spill_n, p.spill_cap);
            const uint64_t lo = (item - s_spill_first[s]) * kSpillSlice;
            const uint64_t hi = min(n_sp, lo + kSpillSlice);
            constexpr uint32_t kRun = kSpillSlice / kAggWarps;
            const uint64_t a = lo + (uint64_t)warp * kRun;
            const bool full = tv.overflow != nullptr && __ldcg(&tv.ctrl->size) >= tv.limit;
            if (a < hi) take_run(p.src[s].spill + a, (uint32_t)min((uint64_t)kRun, hi - a), true, full);
            flush_created();
        }
        __syncthreads();  // the cache holds every occurrence that did not bypass it

        // merge: distinct keys of this item, in slot order of the table (the cache index is the
        // next-lower bits of the same product h * phi)
        {
            const bool full = tv.overflow != nullptr && __ldcg(&tv.ctrl->size) >= tv.limit;
            constexpr int M = 2;
            for (uint32_t base = 0; base < kLocalSlots; base += kAggThreads * M) {
                uint64_t key[M], inc[M];
                uint32_t live = 0;
#pragma unroll
                for (int u = 0; u < M; ++u) {
                    const uint32_t i = base + u * kAggThreads + threadIdx.x;
                    key[u] = lk[i];
                    inc[u] = ld[i];
                    live |= (key[u] != kEmpty ? 1u : 0u) << u;
                }
                if (live) created += table_add_many<M>(tv, key, inc, live, full);
            }
            flush_created();
        }
    }
    for (int o = 16; o; o >>= 1) n_taken += __shfl_xor_sync(0xffffffffu, n_taken, o);
    if (lane == 0 && n_taken) atomicAdd((unsigned long long *)&tv.ctrl->absorbed, (unsigned long long)n_taken);
}

}  // namespace oxg
