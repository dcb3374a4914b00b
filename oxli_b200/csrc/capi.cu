// capi.cu -- host side of the C ABI declared in include/oxli_b200.h.
//
// One DeviceCtx per GPU (streams, staging buffers, scratch), one oxg_table per
// count table.  Everything is plain CUDA runtime: no torch, no Thrust/CUB.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "../../include/oxli_b200.h"
#include "aggregate.cuh"
#include "consume.cuh"
#include "scatter.cuh"
#include "klist.h"
#include "shard.cuh"
#include "sortexport.cuh"
#include "tableops.cuh"

using namespace oxg;

namespace {

// Shards of one process wait for each other inside kernels on different streams; streams that
// share a hardware work queue would serialise a wait in front of the kernel it waits for.  The
// driver reads this when the context is created, so it is set when the library is loaded
// (never overriding the user's choice).
const int g_connections_set = setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

oxg_status fail(oxg_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return st;
}

#define CU(expr)                                                                             \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(e__ == cudaErrorMemoryAllocation ? OXG_ERR_NOMEM : OXG_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                        __LINE__);                                                           \
    } while (0)
#define TRY(expr)                          \
    do {                                   \
        oxg_status s__ = (expr);           \
        if (s__ != OXG_OK) return s__;     \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// NVTX ranges around the stages of an ingest call (copy / scatter / aggregate / fused consume /
// replay / shard round), so that a timeline of the end-to-end step can be read; free without a
// profiler attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// host->device streaming granule: OXLI_B200_CHUNK_MB in 1..1024, anything else means the default
static const uint64_t kChunkBytes = [] {
    const char *e = getenv("OXLI_B200_CHUNK_MB");
    long mb = 64;
    if (e && *e) {
        char *end = nullptr;
        const long v = strtol(e, &end, 10);
        if (end && *end == '\0' && v >= 1 && v <= 1024) mb = v;
    }
    return (uint64_t)mb << 20;
}();
static const uint64_t kLaunchWindows = [] {  // windows per consume launch (bounds the overflow list); OXLI_B200_LAUNCH_MW for experiments
    const char *e = getenv("OXLI_B200_LAUNCH_MW");
    const long v = e && *e ? atol(e) : 64;
    return (uint64_t)(v >= 1 && v <= 1024 ? v : 64) << 20;
}();
constexpr uint64_t kSmallBatch = 1ull << 20;      // below this, reserve for the worst case up front
constexpr uint64_t kMinCap = 1024;
// Launches of at least this many windows go through the partitioned pipeline (pass A scatter,
// pass B aggregate + merge); smaller ones keep the fused kernel, whose fixed costs are lower.
// OXLI_B200_PIPELINE=fused|part overrides the choice (part still needs a specialised k).
constexpr uint64_t kPartMinWindows = 8ull << 20;
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; }
// 0 = by launch size, 1 = fused kernel only, 2 = partitioned whenever the k is specialised
static std::atomic<int> g_pipeline{[] {
    const char *e = getenv("OXLI_B200_PIPELINE");
    return e && !strcmp(e, "fused") ? 1 : e && !strcmp(e, "part") ? 2 : 0;
}()};
static std::atomic<uint32_t> g_parts_override{(uint32_t)env_int("OXLI_B200_PARTS", 0)};    // 0 = from the table size
// pass A launches whose fragments pass B takes together (more duplicates per key meet in one
// aggregation, one merge instead of several); 1 = pass B after every pass A
static std::atomic<uint32_t> g_accumulate{(uint32_t)std::max(1, std::min(16, env_int("OXLI_B200_ACCUMULATE", 8)))};
static std::atomic<uint32_t> g_groups_override{(uint32_t)env_int("OXLI_B200_GROUPS", 0)};  // 0 = default
constexpr int kStageBufs = 4;  // host batches: chunks in flight between the copy stream and the kernels

struct DeviceCtx {
    int dev = -1;
    int sms = 0;
    cudaStream_t stream = nullptr, copy = nullptr;
    std::mutex mu;
    // ring of staging buffers for host batches
    uint8_t *d_stage[kStageBufs] = {};
    uint8_t *h_stage[kStageBufs] = {};
    uint64_t d_stage_cap[kStageBufs] = {}, h_stage_cap[kStageBufs] = {};
    uint64_t *d_offs[kStageBufs] = {};
    uint64_t *h_offs[kStageBufs] = {};
    uint64_t offs_cap[kStageBufs] = {};
    cudaEvent_t ev_ready[kStageBufs] = {};
    cudaEvent_t ev_stage_free[kStageBufs] = {};  // the kernels reading staging buffer b have finished
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_mid = nullptr;
    cudaEvent_t ev_user0 = nullptr, ev_user1 = nullptr;
    // scratch
    uint64_t *d_tile_first = nullptr; uint64_t tile_first_cap = 0;
    ulonglong2 *d_overflow = nullptr;   uint64_t overflow_cap = 0;   // deferred (key, increment) pairs
    ulonglong2 *d_overflow2 = nullptr;  uint64_t overflow2_cap = 0;  // replay target when a replay defers again
    uint64_t *d_dense = nullptr;      // kHistDense bins
    uint64_t *d_big = nullptr;        uint64_t big_cap = 0;
    uint64_t *d_io = nullptr;         uint64_t io_cap = 0;   // generic u64 in/out scratch (device)
    uint64_t *h_io = nullptr;         uint64_t h_io_cap = 0; // pinned mirror
    double *d_f64 = nullptr;          // 4 doubles
    uint8_t *d_seq = nullptr;         uint64_t seq_cap = 0;  // hash_windows input
    // partitioned pipeline (pass A output): fragments, their fill counts, the spill list
    uint64_t *d_frag = nullptr;       uint64_t frag_cap = 0;
    uint32_t *d_frag_cnt = nullptr;   uint64_t frag_cnt_cap = 0;
    uint64_t *d_spill = nullptr;      uint64_t spill_cap = 0;
    unsigned long long *d_spill_n = nullptr;
    bool agg_attr_done = false;
    // export: compacted pairs, the second pair of arrays of the radix sort, its histogram, pinned bounce buffers
    uint64_t *d_exp_k[2] = {}, *d_exp_v[2] = {};
    uint64_t exp_cap[2] = {};
    uint64_t *d_exp_counts = nullptr; uint64_t exp_counts_cap = 0;
    uint64_t *d_sort_hist = nullptr;  uint64_t sort_hist_cap = 0;
    uint8_t *h_bounce[2] = {};
    cudaEvent_t ev_bounce[2] = {};
};

std::mutex g_ctx_mu;
std::vector<std::unique_ptr<DeviceCtx>> g_ctx;

oxg_status get_ctx(int dev, DeviceCtx **out) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(OXG_ERR_CUDA, "no CUDA device available (oxli_b200 has no CPU fallback)");
    }
    if (dev < 0 || dev >= n) return fail(OXG_ERR_INVALID, "device %d out of range (have %d)", dev, n);
    if ((int)g_ctx.size() < n) g_ctx.resize(n);
    if (!g_ctx[dev]) {
        auto c = std::make_unique<DeviceCtx>();
        c->dev = dev;
        CU(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            return fail(OXG_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                        dev, prop.major, prop.minor);
        c->sms = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        for (int i = 0; i < kStageBufs; ++i) {
            CU(cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_stage_free[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreate(&c->ev_t0));
        CU(cudaEventCreate(&c->ev_t1));
        CU(cudaEventCreate(&c->ev_mid));
        CU(cudaEventCreate(&c->ev_user0));
        CU(cudaEventCreate(&c->ev_user1));
        CU(cudaMalloc(&c->d_dense, kHistDense * sizeof(uint64_t)));
        CU(cudaMalloc(&c->d_f64, 4 * sizeof(double)));
        g_ctx[dev] = std::move(c);
    }
    *out = g_ctx[dev].get();
    return OXG_OK;
}

template <class T>
oxg_status ensure_dev(T **p, uint64_t *cap, uint64_t want) {
    if (*cap >= want && *p) return OXG_OK;
    if (*p) CU(cudaFree(*p));
    *p = nullptr; *cap = 0;
    uint64_t ncap = std::max<uint64_t>(want, 1024);
    CU(cudaMalloc(p, ncap * sizeof(T)));
    *cap = ncap;
    return OXG_OK;
}

oxg_status ensure_io(DeviceCtx *c, uint64_t n) {
    TRY(ensure_dev(&c->d_io, &c->io_cap, n));
    if (c->h_io_cap < n) {
        if (c->h_io) CU(cudaFreeHost(c->h_io));
        c->h_io = nullptr; c->h_io_cap = 0;
        uint64_t ncap = std::max<uint64_t>(n, 1024);
        CU(cudaMallocHost(&c->h_io, ncap * sizeof(uint64_t)));
        c->h_io_cap = ncap;
    }
    return OXG_OK;
}

int grid_for(const DeviceCtx *c, uint64_t items, int threads, int per_sm) {
    uint64_t want = (items + threads - 1) / threads;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)c->sms * per_sm));
}

// Capacity policy.  Displaced keys are what slows the probe loop down
// (profiles/r1_microbench_probe_slow_path.txt: 16 % home-bucket misses at load
// 0.6 cut a 62 G keys/s loop to 23-35 G keys/s; at load 0.15 it runs at 66), so
// a table that is small anyway is kept sparse: up to kSparseBytes the target
// load is <= 0.3, beyond that memory and L2 footprint win and it may fill to 0.7.
constexpr uint64_t kSparseBytes = 128ull << 20;
uint64_t capacity_for_keys(uint64_t keys);
bool over_loaded(uint64_t size, uint64_t cap) {
    if (cap * 16 < kSparseBytes) return size * 10 > cap * 3;
    return size * 10 > cap * 7;
}

uint64_t pow2_at_least(uint64_t x) {
    uint64_t c = kMinCap;
    while (c < x) c <<= 1;
    return c;
}

uint64_t capacity_for_keys(uint64_t keys) {
    uint64_t cap = pow2_at_least(keys + keys / 2 + 1);                  // <= 67 % load
    while (cap * 16 < kSparseBytes && keys * 10 > cap * 3) cap <<= 1;   // sparse while it is cheap
    return cap;
}


}  // namespace

namespace {
struct PartPlan {
    uint32_t n_parts = 0, part_bits = 0;   // partitions of one rank's table
    int n_ranks = 1, self_rank = 0, owner_shift = 64;
    uint32_t grid_a = 0;                   // CTAs of pass A = fragments per destination
    uint32_t frag_cap = 0;                 // entries per fragment (multiple of the line)
    uint32_t line_shift = 4;               // entries staged per destination before a line is written: 2^line_shift
    uint64_t spill_cap = 0;
    uint32_t groups = 1;
    uint32_t n_dest() const { return n_parts * (uint32_t)n_ranks; }
    uint64_t frag_entries() const { return (uint64_t)n_dest() * grid_a * frag_cap; }
};

}  // namespace

struct oxg_table {
    DeviceCtx *ctx = nullptr;
    uint32_t k = 0;
    ulonglong2 *slots = nullptr;
    uint64_t cap = 0;
    Ctrl *d_ctrl = nullptr;
    Ctrl *h_ctrl = nullptr;  // pinned
    uint64_t size = 0;       // host mirror of ctrl->size as of the last sync
    bool hinted = false;     // the caller said how many distinct keys to expect
    uint64_t hint_keys = 0;  // ... and this many
    bool pooled = false;     // slot arrays come from the stream-ordered allocator (shards: no call of
                             // theirs may synchronise the device while a peer's flag wait is running)
    float last_ms_a = 0.f, last_ms_b = 0.f;  // partitioned pipeline: pass A / pass B share of last_ms
    // partitioned pipeline: pass A launches whose fragments are waiting for their pass B
    struct {
        bool active = false;
        PartPlan pl;
        uint32_t launches = 0;
        uint64_t windows = 0, planned = 0;
    } pend;
    uint64_t part_budget = 0;  // windows the running consume call still has to go (0 = unknown)
    uint32_t group_cap = 0;    // != 0: at most this many pass A launches per pass B in the running call
    // HyperLogLog sketch of the keys that came in through the partitioned pipeline (aggregate.cuh)
    uint32_t *d_sketch = nullptr;
    uint32_t *h_sketch = nullptr;  // pinned
    uint64_t sketch_covers = 0;    // keys of the table the sketch has seen
    uint64_t est_keys = 0;         // the sketch's last word on how many keys the table is about to hold
    uint64_t last_made = 0;        // keys the previous group of launches created
    uint64_t last_new = 0;   // keys created by the previous consume launch (growth look-ahead)
    float last_ms = 0.f;
    uint64_t last_launches = 0;
};

namespace {

TableView view_of(const oxg_table *t, bool with_overflow) {
    TableView v;
    v.slots = t->slots;
    v.cap = t->cap;
    uint32_t lg = 0;
    while ((1ull << lg) < t->cap) ++lg;
    v.shift = 64 - lg;
    v.limit = t->cap - t->cap / 5;  // stop creating keys at 80 % load
    v.ctrl = t->d_ctrl;
    v.overflow = with_overflow ? t->ctx->d_overflow : nullptr;
    v.overflow_cap = with_overflow ? t->ctx->overflow_cap : 0;
    return v;
}

oxg_status pull_ctrl(oxg_table *t) {
    DeviceCtx *c = t->ctx;
    CU(cudaMemcpyAsync(t->h_ctrl, t->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    t->size = t->h_ctrl->size;
    return OXG_OK;
}

// zero the ctrl fields [first, first+n) (in uint64 units)
oxg_status zero_ctrl_fields(oxg_table *t, size_t first_field, size_t n_fields) {
    CU(cudaMemsetAsync(reinterpret_cast<uint64_t *>(t->d_ctrl) + first_field, 0, n_fields * 8, t->ctx->stream));
    return OXG_OK;
}
constexpr size_t kFieldCounted = offsetof(Ctrl, counted) / 8;
constexpr size_t kFieldTile = offsetof(Ctrl, tile_counter) / 8;
constexpr size_t kLaunchFields = (offsetof(Ctrl, scratch) - offsetof(Ctrl, counted)) / 8;
constexpr size_t kFieldScratch = offsetof(Ctrl, scratch) / 8;

oxg_status alloc_slots(DeviceCtx *c, uint64_t cap, ulonglong2 **out, bool pooled = false) {
    if (pooled) CU(cudaMallocAsync(out, cap * sizeof(ulonglong2), c->stream));
    else CU(cudaMalloc(out, cap * sizeof(ulonglong2)));  // (cudaMallocAsync pools were 2x slower for growing multi-GB tables)
    init_slots_kernel<<<grid_for(c, cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(*out, cap);
    LAUNCHED();
    CU(cudaGetLastError());
    return OXG_OK;
}

// grow (or rebuild at the same size) so that `keys` distinct keys sit at <= 50 % load
oxg_status grow_to_fit(oxg_table *t, uint64_t keys) {
    uint64_t want = std::max(capacity_for_keys(keys), pow2_at_least(keys * 2));
    if (want <= t->cap) return OXG_OK;
    DeviceCtx *c = t->ctx;
    ulonglong2 *fresh = nullptr;
    TRY(alloc_slots(c, want, &fresh, t->pooled));
    ulonglong2 *old = t->slots;
    const uint64_t old_cap = t->cap;
    t->slots = fresh;
    t->cap = want;
    if (t->size) {
        rehash_kernel<<<grid_for(c, old_cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(old, old_cap, view_of(t, false));
        LAUNCHED();
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(c->stream));
    if (t->pooled) CU(cudaFreeAsync(old, c->stream)); else CU(cudaFree(old));
    return OXG_OK;
}

oxg_status reserve_keys(oxg_table *t, uint64_t extra) { return grow_to_fit(t, t->size + extra); }

// A launch ran into the load limit and deferred `ov` hash OCCURRENCES (not distinct keys: on
// high-coverage input every new key is deferred many times).  Grow geometrically -- room for
// at most as many new keys as the table already holds -- and replay; a replay that runs into
// the limit again defers into the second list and the loop goes round once more.
oxg_status drain_deferred(oxg_table *t, uint64_t ov, bool may_allocate = true) {
    NvtxRange range("oxg:replay deferred");
    DeviceCtx *c = t->ctx;
    if (ov > c->overflow_cap) return fail(OXG_ERR_CUDA, "internal: overflow list overrun");
    uint64_t prev = ~0ULL;
    while (ov) {
        const uint64_t room = std::min<uint64_t>(ov, std::max<uint64_t>(t->size, 1ull << 20));
        const uint64_t cap_before = t->cap;
        TRY(grow_to_fit(t, t->size + room));
        // deferred for a probe run that would not end rather than for load: double regardless
        if (t->cap == cap_before && ov >= prev) TRY(grow_to_fit(t, t->cap));
        prev = ov;
        if (may_allocate) TRY(ensure_dev(&c->d_overflow2, &c->overflow2_cap, ov));
        else if (c->overflow2_cap < ov) return fail(OXG_ERR_NOMEM, "deferral list too small for the replay");
        TRY(zero_ctrl_fields(t, offsetof(Ctrl, overflow) / 8, 1));
        TableView v = view_of(t, false);
        v.overflow = c->d_overflow2; v.overflow_cap = c->overflow2_cap;
        replay_pairs_kernel<<<grid_for(c, (ov + 3) / 4, kOpThreads, 8), kOpThreads, 0, c->stream>>>(v, c->d_overflow, ov);
        LAUNCHED();
        CU(cudaGetLastError());
        TRY(pull_ctrl(t));
        ov = t->h_ctrl->overflow;
        std::swap(c->d_overflow, c->d_overflow2);
        std::swap(c->overflow_cap, c->overflow2_cap);
    }
    return OXG_OK;
}

// k values with a compile-time specialised consume kernel: klist.h.  Each lives in its own
// translation unit (consume_inst.cu) and is reached through its entry function.
#define OXG_DECLARE_ENTRY(KK) extern "C" const void *oxg_consume_entry_##KK(int mode);
OXG_FOR_EACH_K(OXG_DECLARE_ENTRY)
#undef OXG_DECLARE_ENTRY
const void *specialised_entry(uint32_t k, int mode) {
    switch (k) {
#define OXG_K_ENTRY(KK) case KK: return oxg_consume_entry_##KK(mode);
        OXG_FOR_EACH_K(OXG_K_ENTRY)
#undef OXG_K_ENTRY
    default:
        return nullptr;
    }
}
bool specialised_k(uint32_t k) { return specialised_entry(k, kModeCount) != nullptr; }
uint32_t tile_width(uint32_t k) { return specialised_k(k) ? kWarpTile : kTileW; }

template <int MODE>
oxg_status launch_consume(oxg_table *t, const ConsumeParams &p) {
    DeviceCtx *c = t->ctx;
    const int k = (int)t->k;
    auto grid_of = [&](const void *fn, size_t smem) {
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const uint64_t tiles_per_cta = specialised_k(t->k) ? kThreads / 32 : 1;
        const uint64_t work = p.n_tiles;
        return (int)std::max<uint64_t>(1, std::min<uint64_t>((work + tiles_per_cta - 1) / tiles_per_cta, (uint64_t)c->sms * per_sm));
    };
    if (const void *fn = specialised_entry(t->k, MODE)) {
        const size_t dyn = consume_dyn_smem(MODE);
        if (dyn) {  // static + dynamic shared memory may pass 48 KB: opt in once per function and device
            static std::mutex attr_mu;
            static std::vector<std::pair<const void *, int>> attr_done;
            std::lock_guard<std::mutex> lk(attr_mu);
            const std::pair<const void *, int> key{fn, c->dev};
            if (std::find(attr_done.begin(), attr_done.end(), key) == attr_done.end()) {
                CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                attr_done.push_back(key);
            }
        }
        void *args[] = {const_cast<ConsumeParams *>(&p)};
        CU(cudaLaunchKernel(fn, dim3(grid_of(fn, dyn)), dim3(kThreads), args, dyn, c->stream));
    } else {
        const size_t smem = generic_smem_bytes(k);
        consume_generic_kernel<MODE><<<grid_of((const void *)consume_generic_kernel<MODE>, smem), kThreads, smem, c->stream>>>(p);
    }
    LAUNCHED();
    CU(cudaGetLastError());
    return OXG_OK;
}


// ---- partitioned pipeline: pass A (hash + scatter) and pass B (aggregate + merge) -----------

// pass A launches per pass B: eight when the reads are resident (pass B 8.0 instead of 9.2 ms per C2
// step), four when they stream in from the host (the last group's pass B is the tail of the call)
uint32_t group_launches(const oxg_table *t) {
    const uint32_t g = g_accumulate.load();
    return t->group_cap ? std::min(g, t->group_cap) : g;
}

// Which pipeline a counting launch of `span` windows takes.  Measured on one B200
// (profiles/r2_pipeline_choice.txt): the partitioned pipeline matches or beats the fused kernel
// everywhere but on small launches -- hot table (C2: 29.0 vs 31.0 ms per step), low-complexity
// input (every window the same k-mer: 30x before the fused kernel learnt to pre-reduce a warp's
// repeats, on a par since), and, with pass B updating the table directly from a few thousand work
// items, input where nearly every key is new (C3-shaped, 273 M keys in 8 GiB: 65 vs 99 ms) -- there
// the fused kernel's updates are random DRAM accesses issued between hashes, pass B's come
// partition by partition.  Small launches stay fused: two kernels and a pass over fragments do
// not pay below some millions of windows.
bool use_partitioned(const oxg_table *t, uint64_t span) {
    if (!specialised_entry(t->k, kModePart)) return false;
    const int choice = g_pipeline.load();
    if (choice == 1) return false;
    if (choice == 2) return true;
    if (t->pend.active) return true;  // a group in flight is completed the way it began
    return span >= kPartMinWindows;
}

// Partition count: 3072-6144 distinct keys per partition, so that the shared-memory table of
// pass B (16384 slots in buckets of four) holds a partition's keys at load 0.2-0.4 and duplicates
// meet there; at most max_parts, because pass A stages one line per destination in shared memory
// (and keeps longer lines with fewer destinations).
uint32_t choose_parts(const oxg_table *t, uint32_t max_parts) {
    const uint32_t forced = g_parts_override.load();
    if (forced >= 2 && !(forced & (forced - 1))) return std::min(forced, std::max(max_parts, 2u));
    const uint64_t est = std::max(t->size + t->last_new, t->hint_keys);
    if (est == 0) return std::min<uint32_t>(1024, max_parts);  // nothing known: the middle of the range
    uint64_t parts = 64;
    while (parts < max_parts && parts * 6144 < est) parts <<= 1;
    return (uint32_t)std::min<uint64_t>(parts, max_parts);
}

// dynamic shared memory of scatter_kernel<k> (mirrors scatter_smem_bytes<K>)
size_t scatter_smem_rt(uint32_t k, uint32_t n_dest, uint32_t line_shift) {
    const uint32_t q = 8 * ((k + 7 + 7) / 8);
    const uint32_t bl = ((kWarpTile - 8 + q) + 15) / 16 * 16;
    return (((size_t)n_dest * ((8u << line_shift) + 12) + 16 + 15) & ~(size_t)15) + (size_t)kScatWarps * (2 * bl + 128);
}

oxg_status plan_partitioned(oxg_table *t, uint64_t span, uint64_t n_tiles, int n_ranks, int self_rank, PartPlan *out) {
    DeviceCtx *c = t->ctx;
    PartPlan pl;
    // destinations are (rank, partition): 1024 of them keep full 128-byte lines staged in pass A
    pl.n_parts = choose_parts(t, std::max(2u, 2048u / (uint32_t)n_ranks));
    while ((1u << pl.part_bits) < pl.n_parts) ++pl.part_bits;
    pl.n_ranks = n_ranks; pl.self_rank = self_rank;
    int lg = 0;
    while ((1 << lg) < n_ranks) ++lg;
    pl.owner_shift = 64 - lg;
    // the staged line: as long as the shared memory of kScatCtasPerSm CTAs per SM allows
    pl.line_shift = 4;
    while (pl.line_shift > 1 && scatter_smem_rt(t->k, pl.n_dest(), pl.line_shift) * kScatCtasPerSm > 220u * 1024) --pl.line_shift;
    if (scatter_smem_rt(t->k, pl.n_dest(), pl.line_shift) * kScatCtasPerSm > 220u * 1024) return fail(OXG_ERR_INVALID, "internal: too many scatter destinations");
    // kScatCtasPerSm CTAs per SM (their staging fills the SM's shared memory); fewer when the launch is small
    const uint64_t tiles_per_cta = kScatWarps;
    pl.grid_a = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n_tiles + tiles_per_cta - 1) / tiles_per_cta, (uint64_t)c->sms * kScatCtasPerSm));
    // a fragment holds its fair share of the launch's windows plus half again plus four standard
    // deviations; what does not fit (skew) goes to the spill list, which can take the whole launch
    const uint64_t line = 1ull << pl.line_shift;
    const uint64_t fair = span / ((uint64_t)pl.n_dest() * pl.grid_a) + 1;
    uint64_t sq = 1;
    while (sq * sq < fair) ++sq;
    pl.frag_cap = (uint32_t)((fair + fair / 2 + 4 * sq + line + line - 1) & ~(line - 1));
    pl.spill_cap = span;
    // work items of pass B = partitions x groups: enough of them (four per SM) that the CTAs finish
    // together -- 256 partitions of a shard on 148 SMs would otherwise be two uneven waves
    const uint32_t forced_groups = g_groups_override.load();
    pl.groups = forced_groups ? forced_groups : std::max<uint32_t>(1, std::min<uint32_t>(8, (4u * (uint32_t)c->sms + pl.n_parts - 1) / pl.n_parts));
    if (pl.frag_entries() >> 32) return fail(OXG_ERR_INVALID, "internal: fragment buffer too large for one launch");
    *out = pl;
    return OXG_OK;
}

// scatter_kernel needs more than 48 KB of dynamic shared memory: opt in once per k and device
oxg_status scatter_attr(DeviceCtx *c, uint32_t k) {
    static std::mutex mu;
    static std::vector<std::pair<uint32_t, int>> done;
    std::lock_guard<std::mutex> lk(mu);
    const std::pair<uint32_t, int> key{k, c->dev};
    if (std::find(done.begin(), done.end(), key) != done.end()) return OXG_OK;
    CU(cudaFuncSetAttribute(specialised_entry(k, kModePart), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done.push_back(key);
    return OXG_OK;
}

oxg_status ensure_part_buffers(DeviceCtx *c, const PartPlan &pl) {
    TRY(ensure_dev(&c->d_frag, &c->frag_cap, pl.frag_entries()));
    TRY(ensure_dev(&c->d_frag_cnt, &c->frag_cnt_cap, (uint64_t)pl.n_dest() * pl.grid_a));
    TRY(ensure_dev(&c->d_spill, &c->spill_cap, pl.spill_cap));
    if (!c->d_spill_n) CU(cudaMalloc(&c->d_spill_n, 8));
    return OXG_OK;
}

// pass A: p is a filled-in ConsumeParams (table.ctrl is where `counted` goes)
oxg_status launch_part_a(oxg_table *t, ConsumeParams p, const PartPlan &pl, uint64_t *d_frag, uint32_t *d_frag_cnt,
                         uint64_t *d_spill, unsigned long long *d_spill_n, cudaStream_t stream, bool append = false) {
    DeviceCtx *c = t->ctx;
    const void *fn = specialised_entry(t->k, kModePart);
    TRY(scatter_attr(c, t->k));
    const size_t dyn = scatter_smem_rt(t->k, pl.n_dest(), pl.line_shift);
    p.frag = d_frag; p.frag_cnt = d_frag_cnt; p.frag_cap = pl.frag_cap;
    p.n_parts = pl.n_parts; p.part_shift = 64 - pl.part_bits; p.n_dest = pl.n_dest(); p.line_shift = pl.line_shift;
    p.n_ranks = pl.n_ranks; p.self_rank = pl.self_rank; p.owner_shift = pl.owner_shift;
    p.spill = d_spill; p.spill_cap = pl.spill_cap; p.spill_n = d_spill_n;
    p.frag_append = append ? 1u : 0u;
    if (!append) CU(cudaMemsetAsync(d_spill_n, 0, 8, stream));
    void *args[] = {&p};
    CU(cudaLaunchKernel(fn, dim3(pl.grid_a), dim3(kScatThreads), args, dyn, stream));
    LAUNCHED();
    CU(cudaGetLastError());
    return OXG_OK;
}

// the aggregation kernel is built for three CTA sizes (one CTA per SM each); OXLI_B200_AGG_THREADS picks
static const int g_agg_threads = [] { const int v = env_int("OXLI_B200_AGG_THREADS", kAggThreadsDefault); return v == 512 || v == 768 || v == 1024 ? v : kAggThreadsDefault; }();
// ... and OXLI_B200_AGG_THREADS_DIRECT for the variant without the cache, which is bound by the latency
// of its table loads and wants every thread it can get (C3-shaped input, pass B per step:
// 512 / 768 / 1024 threads = 46.2 / 35.3 / 32.4 ms)
static const int g_agg_threads_direct = [] { const int v = env_int("OXLI_B200_AGG_THREADS_DIRECT", 1024); return v == 512 || v == 768 || v == 1024 ? v : 1024; }();
int agg_threads(bool cache = true) { return cache ? g_agg_threads : g_agg_threads_direct; }
const void *agg_fn(bool cache = true) {
    if (!cache) return g_agg_threads_direct == 512 ? (const void *)aggregate_kernel<512, false> : g_agg_threads_direct == 768 ? (const void *)aggregate_kernel<768, false> : (const void *)aggregate_kernel<1024, false>;
    return g_agg_threads == 512 ? (const void *)aggregate_kernel<512, true> : g_agg_threads == 1024 ? (const void *)aggregate_kernel<1024, true> : (const void *)aggregate_kernel<768, true>;
}
// grid: one CTA per SM and item at most (the cached kernel fills an SM's shared memory; the direct
// one is held to one CTA per SM by its registers)
oxg_status launch_aggregate(DeviceCtx *c, const AggParams &a, cudaStream_t stream) {
    const bool cache = a.use_cache != 0;
    const size_t smem = aggregate_smem_bytes(cache);
    if (cache && !c->agg_attr_done) {
        CU(cudaFuncSetAttribute(agg_fn(true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->agg_attr_done = true;
    }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, agg_fn(cache), agg_threads(cache), smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint64_t items = (uint64_t)a.n_parts * a.groups;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(items, (uint64_t)c->sms * per_sm));
    void *args[] = {const_cast<AggParams *>(&a)};
    CU(cudaLaunchKernel(agg_fn(cache), dim3(grid), dim3(agg_threads(cache)), args, smem, stream));
    return OXG_OK;
}

// The shared-memory table of pass B pays when keys repeat inside what it is given -- at least two
// occurrences per key the table holds, is hinted to hold or (sketch) is about to hold -- and a
// partition's keys fit it: 16 Ki slots took 4.9 K keys per partition at 9.2 ms per C2 step and
// 9.8 K at 19.6.  Nothing known at all (no hint, no sketch yet): on.
static const int g_cache_override = env_int("OXLI_B200_AGG_CACHE", -1);  // experiments: 0 / 1 force it off / on
uint32_t use_cache_for(const oxg_table *t, const PartPlan &pl, uint64_t windows) {
    if (g_cache_override >= 0) return g_cache_override ? 1u : 0u;
    const uint64_t keys = std::max(std::max(t->size, t->hint_keys), t->est_keys);
    return keys == 0 || (windows >= 2 * keys && keys <= 8192ull * pl.n_parts) ? 1u : 0u;
}

// Work items per partition.  With the cache on, as few as keep the SMs level (the plan's: every
// item empties and merges a cache).  With it off an item is just a share of the partition's
// hashes: some thousands of items of 40 thousand hashes or more measured best (C3-shaped input,
// 8 GiB table, pass B over 512 Mi hashes at a time: 1024 partitions x 1 / 3 / 12 / 25 / 50 / 148
// items = 60.5 / 40.2 / 39.0 / 41.7 / 48.5 / 74.0 ms per step; 256 x 37 / 148 = 37.9 / 43.9; a shard's
// 1024 partitions over 64 Mi hashes at a time: x 3 / 8 / 16 = 86.7 / 94.5 / 120.5 ms per step).
uint32_t aggregate_groups(const PartPlan &pl, uint32_t use_cache, uint64_t hashes) {
    if (use_cache || g_groups_override.load()) return pl.groups;
    const uint64_t items = std::max<uint64_t>(1, hashes / 40960);
    return (uint32_t)std::max<uint64_t>(pl.groups, std::min<uint64_t>(64, (items + pl.n_parts - 1) / pl.n_parts));
}

// pass B over n_src sources (one, this device's own pass A output, without sharding)
oxg_status launch_part_b(oxg_table *t, const PartPlan &pl, const AggSource *src, int n_src, uint64_t windows) {
    DeviceCtx *c = t->ctx;
    AggParams a{};
    a.table = view_of(t, true);
    for (int s = 0; s < n_src; ++s) a.src[s] = src[s];
    a.n_src = n_src;
    a.n_parts = pl.n_parts; a.dest0 = (uint32_t)pl.self_rank * pl.n_parts; a.n_ctas = pl.grid_a;
    a.frag_cap = pl.frag_cap; a.part_bits = pl.part_bits; a.spill_cap = pl.spill_cap;
    a.owner_shift = pl.owner_shift; a.self_rank = pl.self_rank; a.n_ranks = pl.n_ranks;
    a.work_counter = (unsigned long long *)&t->d_ctrl->absorb_counter;
    a.use_cache = use_cache_for(t, pl, windows);
    a.groups = aggregate_groups(pl, a.use_cache, windows);
    TRY(launch_aggregate(c, a, c->stream));
    LAUNCHED();
    CU(cudaGetLastError());
    return OXG_OK;
}

// HyperLogLog estimate from 2^kSketchBits registers (Flajolet et al.; linear counting below 2.5 m)
double sketch_estimate(const uint32_t *regs) {
    const double m = (double)kSketchRegs;
    double sum = 0.0;
    uint32_t zeros = 0;
    for (uint32_t i = 0; i < kSketchRegs; ++i) { sum += std::ldexp(1.0, -(int)regs[i]); zeros += regs[i] == 0; }
    const double alpha = 0.7213 / (1.0 + 1.079 / m);
    double e = alpha * m * m / sum;
    if (e <= 2.5 * m && zeros) e = m * std::log(m / (double)zeros);
    return e;
}

// Before pass B: fold the group's hashes into the table's sketch and make room for what the
// table will hold afterwards -- unless the caller's hint still covers it.  `src` as for pass B.
oxg_status presize_for_group(oxg_table *t, const PartPlan &pl, const AggSource *src, int n_src, cudaStream_t stream,
                             uint64_t arrivals = 0, uint64_t rounds_left = 1) {
    DeviceCtx *c = t->ctx;
    if (!t->d_sketch) {
        CU(cudaMalloc(&t->d_sketch, kSketchRegs * 4));
        CU(cudaMallocHost(&t->h_sketch, kSketchRegs * 4));
        CU(cudaMemsetAsync(t->d_sketch, 0, kSketchRegs * 4, stream));
    }
    AggParams a{};
    for (int s = 0; s < n_src; ++s) a.src[s] = src[s];
    a.n_src = n_src;
    a.n_parts = pl.n_parts; a.dest0 = (uint32_t)pl.self_rank * pl.n_parts; a.n_ctas = pl.grid_a;
    a.frag_cap = pl.frag_cap; a.part_bits = pl.part_bits; a.spill_cap = pl.spill_cap;
    a.owner_shift = pl.owner_shift; a.self_rank = pl.self_rank; a.n_ranks = pl.n_ranks;
    sketch_kernel<<<c->sms * 2, 512, 0, stream>>>(a, t->d_sketch);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(t->h_sketch, t->d_sketch, kSketchRegs * 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const uint64_t unseen = t->size > t->sketch_covers ? t->size - t->sketch_covers : 0;  // keys that came in another way
    const double est = sketch_estimate(t->h_sketch);
    t->est_keys = (uint64_t)est + unseen;
    // sharded tables call this in the first round only, with more rounds of the same size to come:
    // where nearly everything that arrives is new (a quarter or more), room for half that rate over
    // the whole batch, in one step (the rate falls as coverage builds up); else for the rounds in
    // flight before the host monitor catches up
    double headroom = 1.0;
    if (rounds_left > 1) headroom = est * 4 >= (double)arrivals ? std::max(4.0, 0.5 * (double)rounds_left) : (double)std::min<uint64_t>(4, rounds_left);
    uint64_t expect = (uint64_t)(est * 1.04 * headroom) + unseen + 1024;  // 1.04: three standard errors
    if (headroom > 4.0) {  // a guess that size must not cost more than half of what is free
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        while (headroom > 4.0 && std::max(capacity_for_keys(expect), pow2_at_least(expect * 2)) * 16 > free_b / 2) {
            headroom /= 2;
            expect = (uint64_t)(est * 1.04 * std::max(headroom, 4.0)) + unseen + 1024;
        }
    }
    if (expect * 10 > t->cap * 7) TRY(grow_to_fit(t, expect));
    return OXG_OK;
}

// pass B for the pass A launches accumulated so far, then the bookkeeping of a counting launch
oxg_status flush_pending(oxg_table *t, uint64_t *counted) {
    if (!t->pend.active) return OXG_OK;
    NvtxRange range("oxg:aggregate (pass B)");
    DeviceCtx *c = t->ctx;
    t->pend.active = false;
    CU(cudaEventRecord(c->ev_mid, c->stream));
    const AggSource own{c->d_frag, c->d_frag_cnt, c->d_spill, c->d_spill_n};
    // The sketch costs a pass over the group's hashes (0.8 ms per 209 M): taken when the table's
    // size is anybody's guess -- no hint (or the hint is used up), and either nothing is known yet
    // or the room left would not take twice what the previous group created.
    const bool hint_holds = t->hinted && t->size <= t->hint_keys;
    const bool roomy = t->size != 0 && (t->size + 2 * t->last_made) * 10 <= t->cap * 7;
    const bool sketched = !hint_holds && !roomy;
    if (sketched) TRY(presize_for_group(t, t->pend.pl, &own, 1, c->stream));
    TRY(launch_part_b(t, t->pend.pl, &own, 1, t->pend.windows));
    CU(cudaEventRecord(c->ev_t1, c->stream));
    const uint64_t size_before = t->size;
    TRY(pull_ctrl(t));
    float ms = 0.f, ms_a = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    CU(cudaEventElapsedTime(&ms_a, c->ev_t0, c->ev_mid));
    t->last_ms += ms; t->last_launches += t->pend.launches;
    t->last_ms_a += ms_a; t->last_ms_b += ms - ms_a;
    if (counted) *counted += t->h_ctrl->counted;
    const uint64_t ov = t->h_ctrl->overflow;
    // forecast for the next group: what this one created, unless the caller's hint still covers the table
    const uint64_t made = t->size - size_before;
    t->last_new = ov ? made + ov : (t->hinted && t->size <= t->hint_keys) ? 0 : made;
    if (ov) TRY(drain_deferred(t, ov));
    t->last_made = made;
    if (sketched) t->sketch_covers = t->size;
    if (!hint_holds) {
        // Where keys keep coming (a quarter of the windows or more brought a new key), the rest of
        // the call will bring them at no more than this rate: make the room now, in one step,
        // rather than by doubling under load.  0.6: the rate of a read set falls as coverage builds
        // up (C3-shaped input: 0.39 new keys per window in the first group, 0.22 over the whole set).
        if (made * 4 >= t->pend.windows && t->part_budget && t->size * 20 >= t->cap * 7) {  // (only a table that is filling up)
            const uint64_t more = (uint64_t)((double)made / (double)t->pend.windows * 0.6 * (double)t->part_budget);
            const uint64_t want = t->size + more;
            const uint64_t cap_want = std::max(capacity_for_keys(want), pow2_at_least(want * 2));
            if (cap_want > t->cap) {  // (cudaMemGetInfo is a slow driver call, 10-150 ms at times: only when there is a question)
                size_t free_b = 0, total_b = 0;
                CU(cudaMemGetInfo(&free_b, &total_b));
                if (cap_want * 16 < free_b / 2) TRY(grow_to_fit(t, want));
            }
        }
    }
    return OXG_OK;
}

// Run one mode over window starts [w_lo, w_hi) of a device-resident span.
// `bases` holds global positions [g0, data_end); offsets are global positions.
// kModeCount: loops launches of <= kLaunchWindows windows, growing the table and
// replaying deferred hashes between launches.  *counted accumulates.
oxg_status run_span(oxg_table *t, int mode, const uint8_t *d_bases, uint64_t g0, uint64_t w_lo,
                    uint64_t w_hi, uint64_t data_end, const uint64_t *d_offsets, uint64_t n_off,
                    uint64_t *hashes_out, uint64_t *counted) {
    DeviceCtx *c = t->ctx;
    if (w_hi <= w_lo) return OXG_OK;
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (g0 & 15))
        return fail(OXG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t lo = w_lo;
    while (lo < w_hi) {
        const uint64_t tile_base = std::max<uint64_t>(g0, lo & ~(uint64_t)15);
        uint64_t hi = std::min<uint64_t>(w_hi, tile_base + kLaunchWindows);
        const uint32_t tw = tile_width(t->k);
        const uint64_t n_tiles = (hi - tile_base + tw - 1) / tw;
        TRY(ensure_dev(&c->d_tile_first, &c->tile_first_cap, n_tiles));
        if (mode == kModeCount) {
            const uint64_t span = hi - lo;
            // a launch that will not join the group of pass A launches in flight (too small for the
            // partitioned pipeline) must not share its launch counters: complete the group first
            if (t->pend.active && !use_partitioned(t, span)) TRY(flush_pending(t, counted));
            if (span <= kSmallBatch) TRY(reserve_keys(t, span));
            else {
                // Look ahead instead of running into the load limit mid-launch: a table nobody
                // sized gets at least one slot per 8 windows of a full launch (128 MiB for 64 Mi
                // windows: enough for high-coverage data, cheap if it is not), and if the
                // previous launch created keys at a rate that would cross the limit, grow now.
                if (!t->hinted && t->cap < pow2_at_least(span / 8)) TRY(grow_to_fit(t, span / 16));
                const uint64_t expect = t->size + t->last_new + t->last_new / 4;
                if (expect * 10 > t->cap * 7) TRY(grow_to_fit(t, expect));
                else if (over_loaded(t->size, t->cap)) TRY(grow_to_fit(t, t->size));
            }
            TRY(ensure_dev(&c->d_overflow, &c->overflow_cap, std::max<uint64_t>(span, t->pend.active ? t->pend.planned : 0)));
            // counted, overflow, absorbed, tile and absorb counters -- unless a group of pass A
            // launches is accumulating into them
            if (!t->pend.active) TRY(zero_ctrl_fields(t, kFieldCounted, kLaunchFields));
        } else {
            TRY(zero_ctrl_fields(t, kFieldTile, 1));
        }
        ConsumeParams p{};
        p.bases = d_bases; p.g0 = g0; p.w_lo = lo; p.w_hi = hi; p.data_end = data_end;
        p.tile_base = tile_base; p.n_tiles = n_tiles; p.offsets = d_offsets; p.n_off = n_off;
        p.tile_first = c->d_tile_first; p.table = view_of(t, mode == kModeCount);
        p.hashes_out = hashes_out ? hashes_out + (lo - w_lo) : nullptr; p.ksize = t->k;
        tile_first_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, c->stream>>>(d_offsets, n_off, tile_base, n_tiles, tw, c->d_tile_first);
        LAUNCHED();
        CU(cudaGetLastError());
        const bool part = mode == kModeCount && use_partitioned(t, hi - lo);
        if (part) {
            NvtxRange range("oxg:scatter (pass A)");
            // pass A now; pass B once the group of launches is complete (flush_pending)
            const uint64_t span = hi - lo;
            if (t->pend.active && t->pend.windows + span > t->pend.planned) TRY(flush_pending(t, counted));
            if (!t->pend.active) {
                const uint64_t planned = std::min<uint64_t>(std::max<uint64_t>(t->part_budget, span), (uint64_t)group_launches(t) * kLaunchWindows);
                TRY(plan_partitioned(t, planned, planned / tw + 1, 1, 0, &t->pend.pl));
                TRY(ensure_part_buffers(c, t->pend.pl));
                t->pend.active = true; t->pend.launches = 0; t->pend.windows = 0; t->pend.planned = planned;
                CU(cudaEventRecord(c->ev_t0, c->stream));
            }
            TRY(launch_part_a(t, p, t->pend.pl, c->d_frag, c->d_frag_cnt, c->d_spill, c->d_spill_n, c->stream, t->pend.launches > 0));
            t->pend.launches += 1; t->pend.windows += span;
            t->part_budget = t->part_budget > span ? t->part_budget - span : 0;
            if (t->pend.launches >= group_launches(t) || t->pend.windows >= t->pend.planned) TRY(flush_pending(t, counted));
            lo = hi;
            continue;
        }
        NvtxRange range(mode == kModeCount ? "oxg:consume (fused)" : mode == kModeHash ? "oxg:hash windows" : "oxg:first-bad scan");
        CU(cudaEventRecord(c->ev_t0, c->stream));
        if (mode == kModeCount) TRY(launch_consume<kModeCount>(t, p));
        else if (mode == kModeHash) TRY(launch_consume<kModeHash>(t, p));
        else TRY(launch_consume<kModeFirstBad>(t, p));
        CU(cudaEventRecord(c->ev_t1, c->stream));
        if (mode == kModeCount) {
            const uint64_t size_before = t->size;
            TRY(pull_ctrl(t));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
            t->last_ms += ms; t->last_launches += 1;
            if (counted) *counted += t->h_ctrl->counted;
            uint64_t ov = t->h_ctrl->overflow;
            // what the next launch of this stream will probably create: the rate of the last
            // quarter of this one (high-coverage input finds its keys early and then stops
            // creating them; the whole-launch count would quadruple the table for nothing),
            // or everything it wanted when it ran into the limit
            t->last_new = (hi - lo) <= kSmallBatch ? 0 : ov ? (t->size - size_before) + ov : 4 * t->h_ctrl->late_new;
            if (ov) TRY(drain_deferred(t, ov));  // table hit its load limit: grow, then replay the deferred hashes
        }
        lo = hi;
    }
    return OXG_OK;
}

// error-mode driver on a device-resident batch (src/lib.rs:586-600 semantics)
oxg_status consume_resident(oxg_table *t, const uint8_t *d_bases, const uint64_t *d_offsets,
                            const uint64_t *h_offsets /*nullable*/, uint64_t n_reads,
                            uint64_t total, int skip_bad, uint64_t *total_counted,
                            int64_t *err_read, uint64_t *err_pos) {
    DeviceCtx *c = t->ctx;
    const uint64_t k = t->k;
    uint64_t counted = 0;
    if (err_read) *err_read = -1;
    if (err_pos) *err_pos = 0;
    t->last_ms = 0.f; t->last_ms_a = 0.f; t->last_ms_b = 0.f; t->last_launches = 0;
    const uint64_t n_win = total >= k ? total - k + 1 : 0;
    t->part_budget = n_win;
    t->group_cap = 0;
    oxg_status ret = OXG_OK;
    if (skip_bad || n_win == 0) {
        TRY(run_span(t, kModeCount, d_bases, 0, 0, n_win, total, d_offsets, n_reads + 1, nullptr, &counted));
    } else {
        t->h_ctrl->first_bad = ~0ULL;
        CU(cudaMemcpyAsync(&t->d_ctrl->first_bad, &t->h_ctrl->first_bad, 8, cudaMemcpyHostToDevice, c->stream));
        TRY(run_span(t, kModeFirstBad, d_bases, 0, 0, n_win, total, d_offsets, n_reads + 1, nullptr, nullptr));
        TRY(pull_ctrl(t));
        const uint64_t fb = t->h_ctrl->first_bad;
        if (fb == ~0ULL) {
            TRY(run_span(t, kModeCount, d_bases, 0, 0, n_win, total, d_offsets, n_reads + 1, nullptr, &counted));
        } else {
            // read holding position fb: last r with offsets[r] <= fb
            std::vector<uint64_t> tmp;
            if (!h_offsets) {
                tmp.resize(n_reads + 1);
                CU(cudaMemcpyAsync(tmp.data(), d_offsets, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                h_offsets = tmp.data();
            }
            const uint64_t r = (uint64_t)(std::upper_bound(h_offsets, h_offsets + n_reads + 1, fb) - h_offsets) - 1;
            const uint64_t r_start = h_offsets[r];
            uint64_t in_read = 0;
            // everything before that read, then the clean prefix of the read itself
            TRY(run_span(t, kModeCount, d_bases, 0, 0, r_start >= k ? r_start - k + 1 : 0, r_start, d_offsets, n_reads + 1, nullptr, &counted));
            TRY(flush_pending(t, &counted));  // the position reported below counts the windows of the bad read only
            TRY(run_span(t, kModeCount, d_bases, 0, r_start, fb, fb + k - 1, d_offsets, n_reads + 1, nullptr, &in_read));
            TRY(flush_pending(t, &in_read));
            counted += in_read;
            if (err_read) *err_read = (int64_t)r;
            if (err_pos) *err_pos = in_read;
            ret = fail(OXG_ERR_BAD_KMER, "bad k-mer encountered at position %llu", (unsigned long long)in_read);
        }
    }
    if (ret == OXG_OK) TRY(flush_pending(t, &counted));
    t->part_budget = 0;
    if (total_counted) *total_counted = counted;
    return ret;
}

}  // namespace

// ------------------------------------------------------------------ C ABI ---

extern "C" {

const char *oxg_last_error(void) { return g_err.c_str(); }
const char *oxg_version(void) { return "0.3.0"; }  // tracks the reference's Cargo.toml version
int oxg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
uint64_t oxg_launch_count(void) { return g_launches.load(); }

oxg_status oxg_table_create(int device, uint32_t ksize, uint64_t capacity_hint, oxg_table **out) {
    if (!out) return fail(OXG_ERR_INVALID, "out is null");
    *out = nullptr;
    if (ksize < 1 || ksize > 255) return fail(OXG_ERR_INVALID, "ksize must be in 1..255 (got %u)", ksize);
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->dev));
    auto t = std::make_unique<oxg_table>();
    t->ctx = c; t->k = ksize;
    t->cap = capacity_for_keys(capacity_hint);
    t->hinted = capacity_hint != 0;
    t->hint_keys = capacity_hint;
    CU(cudaMalloc(&t->d_ctrl, sizeof(Ctrl)));
    CU(cudaMemsetAsync(t->d_ctrl, 0, sizeof(Ctrl), c->stream));
    CU(cudaMallocHost(&t->h_ctrl, sizeof(Ctrl)));
    memset(t->h_ctrl, 0, sizeof(Ctrl));
    TRY(alloc_slots(c, t->cap, &t->slots));
    CU(cudaStreamSynchronize(c->stream));
    *out = t.release();
    return OXG_OK;
}

oxg_status oxg_table_destroy(oxg_table *t) {
    if (!t) return OXG_OK;
    std::lock_guard<std::mutex> lk(t->ctx->mu);
    cudaSetDevice(t->ctx->dev);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->pooled) cudaFreeAsync(t->slots, t->ctx->stream); else cudaFree(t->slots);
    cudaFree(t->d_ctrl);
    cudaFreeHost(t->h_ctrl);
    if (t->d_sketch) cudaFree(t->d_sketch);
    if (t->h_sketch) cudaFreeHost(t->h_sketch);
    delete t;
    return OXG_OK;
}

#define ENTER(t)                                               \
    if (!(t)) return fail(OXG_ERR_INVALID, "table is null");   \
    DeviceCtx *c = (t)->ctx;                                   \
    std::lock_guard<std::mutex> lk(c->mu);                     \
    CU(cudaSetDevice(c->dev))

oxg_status oxg_table_clear(oxg_table *t) {
    ENTER(t);
    init_slots_kernel<<<grid_for(c, t->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(t->slots, t->cap);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(t->d_ctrl, 0, sizeof(Ctrl), c->stream));
    if (t->d_sketch) CU(cudaMemsetAsync(t->d_sketch, 0, kSketchRegs * 4, c->stream));
    t->sketch_covers = 0;
    t->est_keys = 0;
    t->last_made = 0;
    t->size = 0;
    t->last_new = 0;
    return OXG_OK;
}

oxg_status oxg_table_reserve(oxg_table *t, uint64_t n_keys) {
    ENTER(t);
    return grow_to_fit(t, std::max(n_keys, t->size));
}

oxg_status oxg_table_ksize(const oxg_table *t, uint32_t *ksize) {
    if (!t || !ksize) return fail(OXG_ERR_INVALID, "null argument");
    *ksize = t->k;
    return OXG_OK;
}

oxg_status oxg_table_capacity(const oxg_table *t, uint64_t *slots) {
    if (!t || !slots) return fail(OXG_ERR_INVALID, "null argument");
    *slots = t->cap;
    return OXG_OK;
}

oxg_status oxg_sync(oxg_table *t) {
    ENTER(t);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy));
    return OXG_OK;
}

oxg_status oxg_timer_start(oxg_table *t) {
    ENTER(t);
    CU(cudaEventRecord(c->ev_user0, c->stream));
    return OXG_OK;
}

oxg_status oxg_timer_stop(oxg_table *t, float *ms) {
    ENTER(t);
    if (!ms) return fail(OXG_ERR_INVALID, "null argument");
    CU(cudaEventRecord(c->ev_user1, c->stream));
    CU(cudaEventSynchronize(c->ev_user1));
    CU(cudaEventElapsedTime(ms, c->ev_user0, c->ev_user1));
    return OXG_OK;
}

oxg_status oxg_last_consume_kernel_ms(oxg_table *t, float *ms, uint64_t *launches) {
    if (!t) return fail(OXG_ERR_INVALID, "table is null");
    if (ms) *ms = t->last_ms;
    if (launches) *launches = t->last_launches;
    return OXG_OK;
}

oxg_status oxg_set_pipeline(int choice, uint32_t n_parts, uint32_t groups) {
    if (choice < 0 || choice > 2) return fail(OXG_ERR_INVALID, "pipeline choice must be 0 (by size), 1 (fused) or 2 (partitioned)");
    if (n_parts && (n_parts < 2 || n_parts > 8192 || (n_parts & (n_parts - 1))))
        return fail(OXG_ERR_INVALID, "n_parts must be 0 or a power of two in 2..8192");
    g_pipeline.store(choice);
    g_parts_override.store(n_parts);
    g_groups_override.store(groups);
    return OXG_OK;
}

oxg_status oxg_last_consume_pass_ms(oxg_table *t, float *ms_scatter, float *ms_aggregate) {
    if (!t) return fail(OXG_ERR_INVALID, "table is null");
    if (ms_scatter) *ms_scatter = t->last_ms_a;
    if (ms_aggregate) *ms_aggregate = t->last_ms_b;
    return OXG_OK;
}

// ---- hashing ---------------------------------------------------------------

oxg_status oxg_hash_windows(oxg_table *t, const uint8_t *seq, uint64_t len, uint64_t *hashes_out) {
    ENTER(t);
    const uint64_t k = t->k;
    if (len < k) return OXG_OK;
    if (!seq || !hashes_out) return fail(OXG_ERR_INVALID, "null argument");
    const uint64_t n_win = len - k + 1;
    // piecewise so the device-side output stays bounded; scratch lives in the device context
    // (hash_kmer / count / get call this once per k-mer)
    const uint64_t piece = 4ull << 20;
    TRY(ensure_dev(&c->d_seq, &c->seq_cap, std::min(len, piece + k - 1) + 32));
    TRY(ensure_io(c, std::min(n_win, piece) + 2));
    uint64_t *d_off = c->d_io + std::min(n_win, piece);
    c->h_io[0] = 0; c->h_io[1] = len;
    CU(cudaMemcpyAsync(d_off, c->h_io, 16, cudaMemcpyHostToDevice, c->stream));
    for (uint64_t lo = 0; lo < n_win; lo += piece) {
        const uint64_t hi = std::min(n_win, lo + piece);
        const uint64_t bytes = hi - lo + k - 1;
        CU(cudaMemcpyAsync(c->d_seq, seq + lo, bytes, cudaMemcpyHostToDevice, c->stream));
        TRY(run_span(t, kModeHash, c->d_seq, lo, lo, hi, lo + bytes, d_off, 2, c->d_io, nullptr));
        CU(cudaMemcpyAsync(hashes_out + lo, c->d_io, (hi - lo) * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return OXG_OK;
}

// ---- consume ---------------------------------------------------------------

oxg_status oxg_consume_batch_device(oxg_table *t, const uint8_t *d_bases, const uint64_t *d_offsets,
                                    uint64_t n_reads, uint64_t total_bases, int skip_bad,
                                    uint64_t *total_counted, int64_t *err_read, uint64_t *err_pos) {
    ENTER(t);
    if (total_counted) *total_counted = 0;
    if (n_reads == 0 || total_bases == 0) { if (err_read) *err_read = -1; if (err_pos) *err_pos = 0; return OXG_OK; }
    if (!d_bases || !d_offsets) return fail(OXG_ERR_INVALID, "null argument");
    return consume_resident(t, d_bases, d_offsets, nullptr, n_reads, total_bases, skip_bad, total_counted, err_read, err_pos);
}

oxg_status oxg_hash_batch_device(oxg_table *t, const uint8_t *d_bases, const uint64_t *d_offsets,
                                 uint64_t n_reads, uint64_t total_bases, uint64_t *d_hashes_out) {
    ENTER(t);
    if (!d_bases || !d_offsets || !d_hashes_out) return fail(OXG_ERR_INVALID, "null argument");
    const uint64_t k = t->k;
    if (total_bases < k) return OXG_OK;
    TRY(run_span(t, kModeHash, d_bases, 0, 0, total_bases - k + 1, total_bases, d_offsets, n_reads + 1, d_hashes_out, nullptr));
    CU(cudaStreamSynchronize(c->stream));
    return OXG_OK;
}

// Host batch: stream [w_lo, w_hi) through a ring of kStageBufs staging buffers.  A producer
// thread slices the read boundaries, stages pageable sources into pinned memory and queues the
// H2D copies back to back on the copy stream, up to kStageBufs chunks ahead of the kernels; the
// calling thread launches each chunk's kernels as its copy lands (run_span blocks until they
// are done, which is also what frees the chunk's buffer).  With only one chunk in flight ahead
// the copy engine idled between chunks and the step was 19 ms longer than its kernels.
static oxg_status stream_span(oxg_table *t, int mode, const uint8_t *bases, const uint64_t *offsets,
                              uint64_t n_reads, uint64_t w_lo, uint64_t w_hi, uint64_t data_end,
                              bool src_pinned, uint64_t *counted, const uint8_t *mapped = nullptr) {
    // mapped != nullptr: device-usable alias of `bases` (pinned, mapped host memory); the kernels
    // then read the bases in place over PCIe and only the read boundaries are copied
    DeviceCtx *c = t->ctx;
    const uint64_t k = t->k;
    if (w_hi <= w_lo) return OXG_OK;
    const uint64_t base0 = offsets[0];
    const uint64_t first = w_lo & ~(uint64_t)(kTileW - 1);
    const uint64_t n_chunks = (w_hi - first + kChunkBytes - 1) / kChunkBytes;
    // staging sized to the batch (pinning full chunks costs a third of a second; a caller
    // that hands over a few thousand reads should not pay it), full chunks once one is needed
    const uint64_t stage_bytes = std::min(pow2_at_least(std::max<uint64_t>(data_end - first, 1 << 20)), kChunkBytes) + 256 + 16;
    for (int b = 0; b < (int)std::min<uint64_t>(n_chunks, kStageBufs) && !mapped; ++b) {
        if (c->d_stage_cap[b] < stage_bytes) {
            CU(cudaStreamSynchronize(c->stream));
            if (c->d_stage[b]) CU(cudaFree(c->d_stage[b]));
            c->d_stage[b] = nullptr; c->d_stage_cap[b] = 0;
            CU(cudaMalloc(&c->d_stage[b], stage_bytes));
            c->d_stage_cap[b] = stage_bytes;
        }
        if (!src_pinned && c->h_stage_cap[b] < stage_bytes) {
            CU(cudaStreamSynchronize(c->copy));
            if (c->h_stage[b]) CU(cudaFreeHost(c->h_stage[b]));
            c->h_stage[b] = nullptr; c->h_stage_cap[b] = 0;
            CU(cudaMallocHost(&c->h_stage[b], stage_bytes));
            c->h_stage_cap[b] = stage_bytes;
        }
    }
    struct Slice { uint64_t lo, hi; const uint64_t *ob; uint64_t n_off; };
    auto slice_of = [&](uint64_t ci) {
        Slice s;
        s.lo = first + ci * kChunkBytes;
        s.hi = std::min(data_end, s.lo + kChunkBytes + k - 1);
        // reads whose boundaries can fall inside [lo, hi + 1]
        s.ob = std::lower_bound(offsets, offsets + n_reads + 1, base0 + s.lo + 1);
        s.n_off = (uint64_t)(std::upper_bound(offsets, offsets + n_reads + 1, base0 + s.hi + 1) - s.ob);
        return s;
    };
    // the previous use of buffer b (chunk ci - kStageBufs) is over: its kernels were waited for
    auto issue_copy = [&](uint64_t ci) -> oxg_status {
        NvtxRange range("oxg:stage + H2D copy");
        const int b = (int)(ci % kStageBufs);
        // every chunk also validates its share of the offsets (all pairs are seen once per call),
        // so the O(n_reads) pass runs on the producer thread next to the copies instead of in
        // front of the first one
        for (uint64_t r = n_reads * ci / n_chunks, e = n_reads * (ci + 1) / n_chunks; r < e; ++r)
            if (offsets[r + 1] < offsets[r]) return fail(OXG_ERR_INVALID, "offsets must be non-decreasing");
        const Slice s = slice_of(ci);
        if (ci >= (uint64_t)kStageBufs) {
            // the buffer's previous chunk: its copy out of the pinned staging has finished (host
            // side may be refilled) and its kernels have read the device side (a partitioned
            // launch is only queued when run_span returns)
            CU(cudaEventSynchronize(c->ev_ready[b]));
            CU(cudaStreamWaitEvent(c->copy, c->ev_stage_free[b], 0));
        }
        if (c->offs_cap[b] < s.n_off + 1) {
            CU(cudaStreamSynchronize(c->copy));
            if (c->d_offs[b]) CU(cudaFree(c->d_offs[b]));
            if (c->h_offs[b]) CU(cudaFreeHost(c->h_offs[b]));
            c->d_offs[b] = nullptr; c->h_offs[b] = nullptr; c->offs_cap[b] = 0;
            const uint64_t ncap = std::max<uint64_t>(s.n_off + 1, 1 << 16);
            CU(cudaMalloc(&c->d_offs[b], ncap * 8));
            CU(cudaMallocHost(&c->h_offs[b], ncap * 8));
            c->offs_cap[b] = ncap;
        }
        for (uint64_t i = 0; i < s.n_off; ++i) c->h_offs[b][i] = s.ob[i] - base0;  // batch-relative positions
        c->h_offs[b][s.n_off] = ~0ULL >> 1;  // sentinel keeps n_off >= 1
        const uint8_t *src = bases + base0 + s.lo;
        if (!src_pinned) { memcpy(c->h_stage[b], src, s.hi - s.lo); src = c->h_stage[b]; }
        if (!mapped) CU(cudaMemcpyAsync(c->d_stage[b], src, s.hi - s.lo, cudaMemcpyHostToDevice, c->copy));
        CU(cudaMemcpyAsync(c->d_offs[b], c->h_offs[b], (s.n_off + 1) * 8, cudaMemcpyHostToDevice, c->copy));
        CU(cudaEventRecord(c->ev_ready[b], c->copy));
        return OXG_OK;
    };
    auto run_chunk = [&](uint64_t ci) -> oxg_status {
        const int b = (int)(ci % kStageBufs);
        const Slice s = slice_of(ci);
        CU(cudaStreamWaitEvent(c->stream, c->ev_ready[b], 0));
        TRY(run_span(t, mode, mapped ? mapped + base0 + s.lo : c->d_stage[b], s.lo, std::max(s.lo, w_lo),
                     std::min(w_hi, s.lo + kChunkBytes), s.hi, c->d_offs[b], s.n_off + 1, nullptr, counted));
        CU(cudaEventRecord(c->ev_stage_free[b], c->stream));
        return OXG_OK;
    };
    if (n_chunks == 1) {
        TRY(issue_copy(0));
        TRY(run_chunk(0));
        if (mode != kModeCount) TRY(pull_ctrl(t));  // see below: the buffer is free again when this returns
        return OXG_OK;
    }
    std::mutex mu;
    std::condition_variable cv;
    uint64_t issued = 0, done = 0;  // chunks whose copies are queued / whose kernels have finished
    bool stop = false;
    oxg_status producer_status = OXG_OK;
    std::string producer_err;  // g_err is thread-local: carry the producer's message across
    auto produce = [&] {
        cudaSetDevice(c->dev);
        for (uint64_t ci = 0; ci < n_chunks; ++ci) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || ci < done + kStageBufs; });
                if (stop) return;
            }
            const oxg_status st = issue_copy(ci);
            std::lock_guard<std::mutex> lk(mu);
            if (st != OXG_OK) { producer_status = st; producer_err = g_err; issued = n_chunks; cv.notify_all(); return; }
            issued = ci + 1;
            cv.notify_all();
        }
    };
    std::thread producer;
    try {
        producer = std::thread(produce);
    } catch (const std::system_error &) {
        // no helper thread to be had: stage and run the chunks in turn
        for (uint64_t ci = 0; ci < n_chunks; ++ci) {
            TRY(issue_copy(ci));
            TRY(run_chunk(ci));
            if (mode != kModeCount) {
                TRY(pull_ctrl(t));
                if (mode == kModeFirstBad && t->h_ctrl->first_bad != ~0ULL) break;
            }
        }
        return OXG_OK;
    }
    oxg_status st = OXG_OK;
    for (uint64_t ci = 0; ci < n_chunks && st == OXG_OK; ++ci) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return issued > ci; });
            if (producer_status != OXG_OK) { st = fail(producer_status, "%s", producer_err.c_str()); break; }
        }
        st = run_chunk(ci);
        bool found = false;
        if (st == OXG_OK && mode != kModeCount) {
            // only counting launches are waited for inside run_span; the buffer must not be
            // refilled before this chunk's kernel has read it.  A pre-scan that has found its
            // bad window needs no later chunk.
            st = pull_ctrl(t);
            found = st == OXG_OK && mode == kModeFirstBad && t->h_ctrl->first_bad != ~0ULL;
        }
        std::lock_guard<std::mutex> lk(mu);
        done = ci + 1;
        cv.notify_all();
        if (found) break;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
    producer.join();
    // a call that stopped early (error, or pre-scan hit) may have copies queued that nobody will
    // wait for: drain them before the staging buffers are used again
    if (st != OXG_OK || mode != kModeCount) cudaStreamSynchronize(c->copy);
    return st;
}

oxg_status oxg_consume_batch(oxg_table *t, const uint8_t *bases, const uint64_t *offsets,
                             uint64_t n_reads, int skip_bad, uint64_t *total_counted,
                             int64_t *err_read, uint64_t *err_pos) {
    ENTER(t);
    if (total_counted) *total_counted = 0;
    if (err_read) *err_read = -1;
    if (err_pos) *err_pos = 0;
    if (n_reads == 0) return OXG_OK;
    if (!bases || !offsets) return fail(OXG_ERR_INVALID, "null argument");
    // (monotonicity of the offsets is checked chunk by chunk while they are staged: stream_span)
    if (offsets[n_reads] < offsets[0]) return fail(OXG_ERR_INVALID, "offsets must be non-decreasing");
    const uint64_t k = t->k;
    const uint64_t total = offsets[n_reads] - offsets[0];
    const uint64_t n_win = total >= k ? total - k + 1 : 0;
    t->last_ms = 0.f; t->last_ms_a = 0.f; t->last_ms_b = 0.f; t->last_launches = 0;
    if (n_win == 0) return OXG_OK;
    t->part_budget = n_win;
    t->group_cap = 4;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, bases) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    static const bool zero_copy = [] { const char *e = getenv("OXLI_B200_ZEROCOPY"); return e && *e && strcmp(e, "0") != 0; }();
    const uint8_t *mapped = pinned && zero_copy && attr.devicePointer ? static_cast<const uint8_t *>(attr.devicePointer) : nullptr;
    if (mapped && ((reinterpret_cast<uintptr_t>(mapped) + offsets[0]) & 15)) mapped = nullptr;  // the kernels copy 16-byte units
    uint64_t counted = 0;
    oxg_status ret = OXG_OK;
    if (skip_bad) {
        TRY(stream_span(t, kModeCount, bases, offsets, n_reads, 0, n_win, total, pinned, &counted, mapped));
    } else {
        t->h_ctrl->first_bad = ~0ULL;
        CU(cudaMemcpyAsync(&t->d_ctrl->first_bad, &t->h_ctrl->first_bad, 8, cudaMemcpyHostToDevice, c->stream));
        TRY(stream_span(t, kModeFirstBad, bases, offsets, n_reads, 0, n_win, total, pinned, nullptr));
        TRY(pull_ctrl(t));
        const uint64_t fb = t->h_ctrl->first_bad;
        if (fb == ~0ULL) {
            TRY(stream_span(t, kModeCount, bases, offsets, n_reads, 0, n_win, total, pinned, &counted));
        } else {
            const uint64_t base0 = offsets[0];
            const uint64_t r = (uint64_t)(std::upper_bound(offsets, offsets + n_reads + 1, base0 + fb) - offsets) - 1;
            const uint64_t r_start = offsets[r] - base0;
            uint64_t in_read = 0;
            TRY(stream_span(t, kModeCount, bases, offsets, n_reads, 0, r_start >= k ? r_start - k + 1 : 0, r_start, pinned, &counted));
            TRY(flush_pending(t, &counted));  // the position reported below counts the windows of the bad read only
            TRY(stream_span(t, kModeCount, bases, offsets, n_reads, r_start, fb, fb + k - 1, pinned, &in_read));
            TRY(flush_pending(t, &in_read));
            counted += in_read;
            if (err_read) *err_read = (int64_t)r;
            if (err_pos) *err_pos = in_read;
            ret = fail(OXG_ERR_BAD_KMER, "bad k-mer encountered at position %llu", (unsigned long long)in_read);
        }
    }
    if (ret == OXG_OK) TRY(flush_pending(t, &counted));
    t->part_budget = 0;
    t->group_cap = 0;
    if (total_counted) *total_counted = counted;
    return ret;
}

// ---- by-hash operations ----------------------------------------------------

// counts a device-resident hash list in launches of <= kLaunchWindows entries; large
// lists go in optimistically (load-limit deferral + growth + replay) like consume
static oxg_status count_list_device(oxg_table *t, const uint64_t *d_hashes, uint64_t n, int skip_zero, uint64_t *counted) {
    DeviceCtx *c = t->ctx;
    for (uint64_t lo = 0; lo < n; lo += kLaunchWindows) {
        const uint64_t m = std::min<uint64_t>(kLaunchWindows, n - lo);
        const bool optimistic = m > kSmallBatch;
        if (!optimistic) TRY(reserve_keys(t, m));
        else {
            if (over_loaded(t->size, t->cap)) TRY(grow_to_fit(t, t->size));
            TRY(ensure_dev(&c->d_overflow, &c->overflow_cap, m));
        }
        TRY(zero_ctrl_fields(t, kFieldCounted, kLaunchFields));
        CU(cudaEventRecord(c->ev_t0, c->stream));
        count_hashes_kernel<<<grid_for(c, (m + 7) / 8, kOpThreads, 8), kOpThreads, 0, c->stream>>>(view_of(t, optimistic), d_hashes + lo, m, nullptr, skip_zero);
        LAUNCHED();
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev_t1, c->stream));
        TRY(pull_ctrl(t));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        t->last_ms += ms; t->last_launches += 1;
        if (counted) *counted += t->h_ctrl->counted;
        const uint64_t ov = t->h_ctrl->overflow;
        if (ov) TRY(drain_deferred(t, ov));
    }
    return OXG_OK;
}

oxg_status oxg_count_hashes_device(oxg_table *t, const uint64_t *d_hashes, uint64_t n, int skip_zero, uint64_t *n_counted) {
    ENTER(t);
    if (n_counted) *n_counted = 0;
    if (n == 0) return OXG_OK;
    if (!d_hashes) return fail(OXG_ERR_INVALID, "null argument");
    t->last_ms = 0.f; t->last_ms_a = 0.f; t->last_ms_b = 0.f; t->last_launches = 0;
    return count_list_device(t, d_hashes, n, skip_zero, n_counted);
}

oxg_status oxg_count_hashes(oxg_table *t, const uint64_t *hashes, uint64_t n, uint64_t *new_counts) {
    ENTER(t);
    if (n == 0) return OXG_OK;
    if (!hashes) return fail(OXG_ERR_INVALID, "null argument");
    TRY(reserve_keys(t, n));
    TRY(ensure_io(c, 2 * n));
    memcpy(c->h_io, hashes, n * 8);
    CU(cudaMemcpyAsync(c->d_io, c->h_io, n * 8, cudaMemcpyHostToDevice, c->stream));
    count_hashes_kernel<<<grid_for(c, n, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(t, false), c->d_io, n, new_counts ? c->d_io + n : nullptr, 0);
    LAUNCHED();
    CU(cudaGetLastError());
    if (new_counts) CU(cudaMemcpyAsync(c->h_io + n, c->d_io + n, n * 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(pull_ctrl(t));
    if (new_counts) memcpy(new_counts, c->h_io + n, n * 8);
    return OXG_OK;
}

oxg_status oxg_add_pairs(oxg_table *t, const uint64_t *keys, const uint64_t *vals, uint64_t n) {
    ENTER(t);
    if (n == 0) return OXG_OK;
    if (!keys || !vals) return fail(OXG_ERR_INVALID, "null argument");
    TRY(reserve_keys(t, n));
    TRY(ensure_io(c, 2 * n));
    memcpy(c->h_io, keys, n * 8);
    memcpy(c->h_io + n, vals, n * 8);
    CU(cudaMemcpyAsync(c->d_io, c->h_io, 2 * n * 8, cudaMemcpyHostToDevice, c->stream));
    add_pairs_kernel<<<grid_for(c, n, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(t, false), c->d_io, c->d_io + n, n);
    LAUNCHED();
    CU(cudaGetLastError());
    // the out-of-band key's side entry must exist even when its value is 0
    for (uint64_t i = 0; i < n; ++i)
        if (keys[i] == kEmpty) { const uint64_t one = 1; CU(cudaMemcpyAsync(&t->d_ctrl->side_present, &one, 8, cudaMemcpyHostToDevice, c->stream)); break; }
    return pull_ctrl(t);
}

oxg_status oxg_get_hashes(oxg_table *t, const uint64_t *hashes, uint64_t n, uint64_t *counts_out) {
    ENTER(t);
    if (n == 0) return OXG_OK;
    if (!hashes || !counts_out) return fail(OXG_ERR_INVALID, "null argument");
    TRY(ensure_io(c, 2 * n));
    memcpy(c->h_io, hashes, n * 8);
    CU(cudaMemcpyAsync(c->d_io, c->h_io, n * 8, cudaMemcpyHostToDevice, c->stream));
    get_hashes_kernel<<<grid_for(c, n, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(t, false), c->d_io, n, c->d_io + n);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_io + n, c->d_io + n, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(counts_out, c->h_io + n, n * 8);
    return OXG_OK;
}

oxg_status oxg_set_hash(oxg_table *t, uint64_t hash, uint64_t value) {
    ENTER(t);
    TRY(reserve_keys(t, 1));
    set_hash_kernel<<<1, 32, 0, c->stream>>>(view_of(t, false), hash, value);
    LAUNCHED();
    CU(cudaGetLastError());
    return pull_ctrl(t);
}

oxg_status oxg_erase_hashes(oxg_table *t, const uint64_t *hashes, uint64_t n, uint64_t *n_removed) {
    ENTER(t);
    if (n_removed) *n_removed = 0;
    if (n == 0) return OXG_OK;
    if (!hashes) return fail(OXG_ERR_INVALID, "null argument");
    TRY(ensure_io(c, n));
    memcpy(c->h_io, hashes, n * 8);
    CU(cudaMemcpyAsync(c->d_io, c->h_io, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (n < 256) {
        // the reference's usage: one key per call (drop / drop_hash, src/lib.rs:197-224)
        erase_hashes_kernel<<<1, 32, 0, c->stream>>>(view_of(t, false), c->d_io, n);
        LAUNCHED();
        CU(cudaGetLastError());
        TRY(pull_ctrl(t));
        if (n_removed) *n_removed = t->h_ctrl->scratch[0];
        return OXG_OK;
    }
    // a list: mark in parallel, rebuild once
    TRY(pull_ctrl(t));
    uint32_t *d_doomed = nullptr;
    const uint64_t words = (t->cap + 31) / 32;
    CU(cudaMalloc(&d_doomed, words * 4));
    CU(cudaMemsetAsync(d_doomed, 0, words * 4, c->stream));
    TRY(zero_ctrl_fields(t, kFieldScratch, 1));
    erase_mark_kernel<<<grid_for(c, n, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(t, false), c->d_io, n, d_doomed);
    LAUNCHED();
    ulonglong2 *fresh = nullptr;
    TRY(alloc_slots(c, t->cap, &fresh, t->pooled));
    ulonglong2 *old = t->slots;
    t->slots = fresh;
    erase_rebuild_kernel<<<grid_for(c, t->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(old, t->cap, view_of(t, false), d_doomed);
    LAUNCHED();
    CU(cudaGetLastError());
    TRY(pull_ctrl(t));
    if (t->pooled) CU(cudaFreeAsync(old, c->stream)); else CU(cudaFree(old));
    CU(cudaFree(d_doomed));
    uint64_t removed = t->h_ctrl->scratch[0];
    t->h_ctrl->size -= removed;
    if (t->h_ctrl->side_present && std::find(hashes, hashes + n, kEmpty) != hashes + n) {
        t->h_ctrl->side_present = 0; t->h_ctrl->side_count = 0; ++removed;
    }
    t->size = t->h_ctrl->size;
    CU(cudaMemcpyAsync(t->d_ctrl, t->h_ctrl, offsetof(Ctrl, counted), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_removed) *n_removed = removed;
    return OXG_OK;
}

oxg_status oxg_cut(oxg_table *t, int mode, uint64_t thresh, uint64_t *n_removed) {
    ENTER(t);
    if (mode != 0 && mode != 1) return fail(OXG_ERR_INVALID, "mode must be 0 (mincut) or 1 (maxcut)");
    TRY(pull_ctrl(t));
    uint64_t removed = 0;
    ulonglong2 *fresh = nullptr;
    TRY(alloc_slots(c, t->cap, &fresh, t->pooled));
    ulonglong2 *old = t->slots;
    t->slots = fresh;
    TRY(zero_ctrl_fields(t, kFieldScratch, 1));
    cut_kernel<<<grid_for(c, t->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(old, t->cap, view_of(t, false), mode, thresh);
    LAUNCHED();
    CU(cudaGetLastError());
    TRY(pull_ctrl(t));
    if (t->pooled) CU(cudaFreeAsync(old, c->stream)); else CU(cudaFree(old));
    removed = t->h_ctrl->scratch[0];
    t->h_ctrl->size -= removed;
    if (t->h_ctrl->side_present) {
        const uint64_t v = t->h_ctrl->side_count;
        if (mode == 0 ? v < thresh : v > thresh) { t->h_ctrl->side_present = 0; t->h_ctrl->side_count = 0; ++removed; }
    }
    t->size = t->h_ctrl->size;
    CU(cudaMemcpyAsync(t->d_ctrl, t->h_ctrl, offsetof(Ctrl, counted), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_removed) *n_removed = removed;
    return OXG_OK;
}

// ---- scans -------------------------------------------------------------------

static oxg_status run_stats(oxg_table *t, bool histo, oxg_stats *st, uint64_t *n_big) {
    DeviceCtx *c = t->ctx;
    if (histo) {
        CU(cudaMemsetAsync(c->d_dense, 0, kHistDense * 8, c->stream));
        if (!c->d_big) TRY(ensure_dev(&c->d_big, &c->big_cap, 1 << 16));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        Ctrl init{};
        init.scratch[2] = ~0ULL;
        memcpy(t->h_ctrl->scratch, init.scratch, sizeof init.scratch);
        CU(cudaMemcpyAsync(t->d_ctrl->scratch, t->h_ctrl->scratch, sizeof init.scratch, cudaMemcpyHostToDevice, c->stream));
        stats_kernel<<<grid_for(c, t->cap, kOpThreads, 8), kOpThreads, 0, c->stream>>>(view_of(t, false), histo ? c->d_dense : nullptr, c->d_big, c->big_cap);
        LAUNCHED();
        CU(cudaGetLastError());
        TRY(pull_ctrl(t));
        if (!histo || t->h_ctrl->scratch[4] <= c->big_cap) break;
        // more huge counts than the list holds: enlarge and redo
        TRY(ensure_dev(&c->d_big, &c->big_cap, t->h_ctrl->scratch[4]));
        CU(cudaMemsetAsync(c->d_dense, 0, kHistDense * 8, c->stream));
    }
    const Ctrl *h = t->h_ctrl;
    st->len = h->scratch[0]; st->sum = h->scratch[1];
    st->min = h->scratch[0] ? h->scratch[2] : ~0ULL; st->max = h->scratch[3];
    if (h->side_present) {
        st->len += 1; st->sum += h->side_count;
        st->min = std::min(st->min, h->side_count); st->max = std::max(st->max, h->side_count);
    }
    if (st->len == 0) { st->min = 0; st->max = 0; }
    if (n_big) *n_big = h->scratch[4];
    return OXG_OK;
}

oxg_status oxg_table_len(oxg_table *t, uint64_t *len) {
    ENTER(t);
    if (!len) return fail(OXG_ERR_INVALID, "null argument");
    TRY(pull_ctrl(t));
    *len = t->h_ctrl->size + (t->h_ctrl->side_present ? 1 : 0);
    return OXG_OK;
}

oxg_status oxg_table_stats(oxg_table *t, oxg_stats *out) {
    ENTER(t);
    if (!out) return fail(OXG_ERR_INVALID, "null argument");
    return run_stats(t, false, out, nullptr);
}

oxg_status oxg_histo(oxg_table *t, uint64_t *freq, uint64_t *n, uint64_t cap, uint64_t *n_out) {
    ENTER(t);
    if (!n_out) return fail(OXG_ERR_INVALID, "null argument");
    oxg_stats st;
    uint64_t n_big = 0;
    TRY(run_stats(t, true, &st, &n_big));
    std::vector<uint64_t> dense(kHistDense), big(n_big);
    CU(cudaMemcpyAsync(dense.data(), c->d_dense, kHistDense * 8, cudaMemcpyDeviceToHost, c->stream));
    if (n_big) CU(cudaMemcpyAsync(big.data(), c->d_big, n_big * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (t->h_ctrl->side_present) {
        const uint64_t v = t->h_ctrl->side_count;
        if (v < kHistDense) dense[v] += 1; else big.push_back(v);
    }
    std::sort(big.begin(), big.end());
    uint64_t w = 0, total = 0;
    auto emit = [&](uint64_t f, uint64_t cnt) {
        if (w < cap && freq && n) { freq[w] = f; n[w] = cnt; ++w; }
        ++total;
    };
    for (uint64_t f = 0; f < kHistDense; ++f) if (dense[f]) emit(f, dense[f]);
    for (size_t i = 0; i < big.size();) {
        size_t j = i;
        while (j < big.size() && big[j] == big[i]) ++j;
        emit(big[i], j - i);
        i = j;
    }
    *n_out = total;
    return OXG_OK;
}

oxg_status oxg_table_digest(oxg_table *t, int n_ranks, int rank, uint64_t out[5]) {
    ENTER(t);
    if (!out) return fail(OXG_ERR_INVALID, "null argument");
    if (n_ranks < 1 || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks) return fail(OXG_ERR_INVALID, "bad rank / n_ranks");
    int lg = 0;
    while ((1 << lg) < n_ranks) ++lg;
    CU(cudaMemsetAsync(t->d_ctrl->scratch, 0, 5 * 8, c->stream));
    digest_kernel<<<grid_for(c, t->cap, kOpThreads, 8), kOpThreads, 0, c->stream>>>(view_of(t, false), 64 - lg, (uint64_t)rank);
    LAUNCHED();
    CU(cudaGetLastError());
    TRY(pull_ctrl(t));
    const Ctrl *h = t->h_ctrl;
    for (int i = 0; i < 5; ++i) out[i] = h->scratch[i];
    if (h->side_present) {  // the out-of-band key 2^64-1 (owner: the last rank)
        out[0] += 1; out[1] += h->side_count; out[2] ^= kEmpty; out[3] += kEmpty * h->side_count;
        if (n_ranks > 1 && rank != n_ranks - 1) out[4] += 1;
    }
    return OXG_OK;
}

// ---- export --------------------------------------------------------------------

constexpr uint64_t kBounceBytes = 32ull << 20;

// ordered compaction of the live slots into the context's export arrays [0] (slot order: stable
// between calls while the table is not modified, so dump() == list(iter), src/lib.rs:330-381,658)
static oxg_status export_device(oxg_table *t, uint64_t *n_live) {
    DeviceCtx *c = t->ctx;
    const uint64_t n_chunks = (t->cap + kExportChunk - 1) / kExportChunk;
    TRY(ensure_dev(&c->d_exp_counts, &c->exp_counts_cap, n_chunks + 1));
    export_count_kernel<<<(unsigned)n_chunks, kOpThreads, 0, c->stream>>>(t->slots, t->cap, c->d_exp_counts);
    LAUNCHED();
    export_scan_kernel<<<1, 1024, 0, c->stream>>>(c->d_exp_counts, n_chunks, c->d_exp_counts + n_chunks);
    LAUNCHED();
    CU(cudaGetLastError());
    uint64_t live = 0;
    CU(cudaMemcpyAsync(&live, c->d_exp_counts + n_chunks, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->exp_cap[0] < live + 1) {
        if (c->d_exp_k[0]) { CU(cudaFree(c->d_exp_k[0])); CU(cudaFree(c->d_exp_v[0])); }
        c->d_exp_k[0] = c->d_exp_v[0] = nullptr; c->exp_cap[0] = 0;
        CU(cudaMalloc(&c->d_exp_k[0], (live + 1) * 8));
        CU(cudaMalloc(&c->d_exp_v[0], (live + 1) * 8));
        c->exp_cap[0] = live + 1;
    }
    export_write_kernel<<<(unsigned)n_chunks, kOpThreads, 0, c->stream>>>(t->slots, t->cap, c->d_exp_counts, c->d_exp_k[0], c->d_exp_v[0], live);
    LAUNCHED();
    CU(cudaGetLastError());
    *n_live = live;
    return OXG_OK;
}

// stable LSD radix sort of the n pairs in export arrays [0]; *which = the pair of arrays holding the result
static oxg_status sort_pairs_device(DeviceCtx *c, uint64_t n, int sort_mode, uint64_t max_count, int *which) {
    *which = 0;
    if (n < 2) return OXG_OK;
    if (c->exp_cap[1] < n) {
        if (c->d_exp_k[1]) { CU(cudaFree(c->d_exp_k[1])); CU(cudaFree(c->d_exp_v[1])); }
        c->d_exp_k[1] = c->d_exp_v[1] = nullptr; c->exp_cap[1] = 0;
        CU(cudaMalloc(&c->d_exp_k[1], n * 8));
        CU(cudaMalloc(&c->d_exp_v[1], n * 8));
        c->exp_cap[1] = n;
    }
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    TRY(ensure_dev(&c->d_sort_hist, &c->sort_hist_cap, 256 * n_tiles));
    int cur = 0;
    auto pass = [&](int shift, int by_count) -> oxg_status {
        radix_hist_kernel<<<(unsigned)n_tiles, kSortThreads, 0, c->stream>>>(c->d_exp_k[cur], c->d_exp_v[cur], n, shift, by_count, n_tiles, c->d_sort_hist);
        LAUNCHED();
        radix_scan_kernel<<<1, 1024, 0, c->stream>>>(c->d_sort_hist, 256 * n_tiles);
        LAUNCHED();
        radix_scatter_kernel<<<(unsigned)n_tiles, kSortThreads, 0, c->stream>>>(c->d_exp_k[cur], c->d_exp_v[cur], c->d_exp_k[cur ^ 1], c->d_exp_v[cur ^ 1],
                                                                                n, shift, by_count, n_tiles, c->d_sort_hist);
        LAUNCHED();
        CU(cudaGetLastError());
        cur ^= 1;
        return OXG_OK;
    };
    for (int shift = 0; shift < 64; shift += 8) TRY(pass(shift, 0));           // by hash
    if (sort_mode == 2)                                                        // then, stably, by count
        for (int shift = 0; shift < 64 && (max_count >> shift) != 0; shift += 8) TRY(pass(shift, 1));
    *which = cur;
    return OXG_OK;
}

// device array -> caller's host array through two pinned bounce buffers (the caller's memory is
// usually pageable: a plain cudaMemcpy would stage it the same way, but one chunk at a time)
static oxg_status download(DeviceCtx *c, uint64_t *dst, const uint64_t *d_src, uint64_t n) {
    if (!dst || n == 0) return OXG_OK;
    for (int b = 0; b < 2; ++b)
        if (!c->h_bounce[b]) {
            CU(cudaMallocHost(&c->h_bounce[b], kBounceBytes));
            CU(cudaEventCreateWithFlags(&c->ev_bounce[b], cudaEventDisableTiming));
        }
    const uint64_t per = kBounceBytes / 8;
    const uint64_t n_chunks = (n + per - 1) / per;
    for (uint64_t i = 0; i <= n_chunks; ++i) {
        if (i < n_chunks) {
            const uint64_t lo = i * per, cnt = std::min(per, n - lo);
            CU(cudaMemcpyAsync(c->h_bounce[i & 1], d_src + lo, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaEventRecord(c->ev_bounce[i & 1], c->stream));
        }
        if (i > 0) {  // chunk i-1 has landed while chunk i is on its way
            const uint64_t lo = (i - 1) * per, cnt = std::min(per, n - lo);
            CU(cudaEventSynchronize(c->ev_bounce[(i - 1) & 1]));
            memcpy(dst + lo, c->h_bounce[(i - 1) & 1], cnt * 8);
        }
    }
    return OXG_OK;
}

oxg_status oxg_export(oxg_table *t, uint64_t *keys, uint64_t *vals, uint64_t cap, int sort_mode, uint64_t *n_out) {
    ENTER(t);
    if (!n_out) return fail(OXG_ERR_INVALID, "null argument");
    if (sort_mode < 0 || sort_mode > 2) return fail(OXG_ERR_INVALID, "sort_mode must be 0, 1 or 2");
    TRY(pull_ctrl(t));
    const bool side = t->h_ctrl->side_present != 0;
    const uint64_t side_count = t->h_ctrl->side_count;
    const uint64_t total = t->h_ctrl->size + (side ? 1 : 0);
    *n_out = total;
    if (cap == 0 || total == 0) return OXG_OK;
    uint64_t max_count = 0;
    if (sort_mode == 2) {
        oxg_stats st;
        TRY(run_stats(t, false, &st, nullptr));
        max_count = st.max;
    }
    uint64_t live = 0;
    TRY(export_device(t, &live));
    uint64_t n = live;
    if (side) {  // the out-of-band key 2^64-1 lives outside the slot array: it joins the pairs here
        const uint64_t kv[2] = {kEmpty, side_count};
        CU(cudaMemcpyAsync(c->d_exp_k[0] + n, &kv[0], 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_exp_v[0] + n, &kv[1], 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        ++n;
    }
    int which = 0;
    if (sort_mode != 0) TRY(sort_pairs_device(c, n, sort_mode, max_count, &which));
    TRY(download(c, keys, c->d_exp_k[which], std::min(n, cap)));
    TRY(download(c, vals, c->d_exp_v[which], std::min(n, cap)));
    CU(cudaStreamSynchronize(c->stream));
    return OXG_OK;
}

// ---- set comparisons -------------------------------------------------------------

static oxg_status two_tables(oxg_table *a, oxg_table *b) {
    if (!a || !b) return fail(OXG_ERR_INVALID, "table is null");
    if (a->ctx != b->ctx) return fail(OXG_ERR_INVALID, "tables live on different devices");
    return OXG_OK;
}

static oxg_status setop_sizes_locked(oxg_table *a, oxg_table *b, uint64_t *inter, uint64_t *uni) {
    DeviceCtx *c = a->ctx;
    TRY(pull_ctrl(b));
    TRY(zero_ctrl_fields(a, kFieldScratch, 1));
    setop_count_kernel<<<grid_for(c, a->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(a, false), view_of(b, false));
    LAUNCHED();
    CU(cudaGetLastError());
    TRY(pull_ctrl(a));
    uint64_t both = a->h_ctrl->scratch[0];
    if (a->h_ctrl->side_present && b->h_ctrl->side_present) both += 1;
    const uint64_t na = a->h_ctrl->size + (a->h_ctrl->side_present ? 1 : 0);
    const uint64_t nb = b->h_ctrl->size + (b->h_ctrl->side_present ? 1 : 0);
    if (inter) *inter = both;
    if (uni) *uni = na + nb - both;
    return OXG_OK;
}

oxg_status oxg_setop_sizes(oxg_table *a, oxg_table *b, uint64_t *inter, uint64_t *uni) {
    TRY(two_tables(a, b));
    ENTER(a);
    return setop_sizes_locked(a, b, inter, uni);
}

oxg_status oxg_jaccard(oxg_table *a, oxg_table *b, double *out) {
    TRY(two_tables(a, b));
    ENTER(a);
    if (!out) return fail(OXG_ERR_INVALID, "null argument");
    uint64_t inter = 0, uni = 0;
    TRY(setop_sizes_locked(a, b, &inter, &uni));
    *out = uni == 0 ? 1.0 : (double)inter / (double)uni;  // src/lib.rs:716-721
    return OXG_OK;
}

oxg_status oxg_setop_export(oxg_table *a, oxg_table *b, int op, uint64_t *keys_out, uint64_t cap, uint64_t *n_out) {
    TRY(two_tables(a, b));
    ENTER(a);
    if (!n_out) return fail(OXG_ERR_INVALID, "null argument");
    if (op < 0 || op > 3) return fail(OXG_ERR_INVALID, "unknown set operation %d", op);
    TRY(pull_ctrl(a));
    TRY(pull_ctrl(b));
    const uint64_t room = a->h_ctrl->size + b->h_ctrl->size + 2;
    uint64_t *d_out = nullptr, *d_cnt = nullptr;
    CU(cudaMalloc(&d_out, room * 8));
    CU(cudaMalloc(&d_cnt, 8));
    CU(cudaMemsetAsync(d_cnt, 0, 8, c->stream));
    auto pass = [&](oxg_table *x, oxg_table *y, int want) -> oxg_status {
        setop_export_kernel<<<grid_for(c, x->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(x, false), view_of(y, false), want, d_out, room, d_cnt);
        LAUNCHED();
        CU(cudaGetLastError());
        return OXG_OK;
    };
    switch (op) {
    case OXG_UNION: TRY(pass(a, b, 2)); TRY(pass(b, a, 0)); break;
    case OXG_INTERSECTION: TRY(pass(a, b, 1)); break;
    case OXG_DIFFERENCE: TRY(pass(a, b, 0)); break;
    default: TRY(pass(a, b, 0)); TRY(pass(b, a, 0)); break;
    }
    uint64_t n = 0;
    CU(cudaMemcpyAsync(&n, d_cnt, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> host(n + 1);
    if (n) CU(cudaMemcpy(host.data(), d_out, n * 8, cudaMemcpyDeviceToHost));
    CU(cudaFree(d_out));
    CU(cudaFree(d_cnt));
    // the out-of-band key
    const bool sa = a->h_ctrl->side_present, sb = b->h_ctrl->side_present;
    bool side = false;
    switch (op) {
    case OXG_UNION: side = sa || sb; break;
    case OXG_INTERSECTION: side = sa && sb; break;
    case OXG_DIFFERENCE: side = sa && !sb; break;
    default: side = sa != sb; break;
    }
    if (side) host[n++] = kEmpty;
    *n_out = n;
    if (keys_out) memcpy(keys_out, host.data(), std::min(n, cap) * 8);
    return OXG_OK;
}

oxg_status oxg_cosine(oxg_table *a, oxg_table *b, double *out) {
    TRY(two_tables(a, b));
    ENTER(a);
    if (!out) return fail(OXG_ERR_INVALID, "null argument");
    TRY(pull_ctrl(a));
    TRY(pull_ctrl(b));
    const uint64_t na = a->h_ctrl->size + (a->h_ctrl->side_present ? 1 : 0);
    const uint64_t nb = b->h_ctrl->size + (b->h_ctrl->side_present ? 1 : 0);
    if (na == 0 || nb == 0) { *out = 0.0; return OXG_OK; }  // src/lib.rs:729-731
    CU(cudaMemsetAsync(c->d_f64, 0, 4 * sizeof(double), c->stream));
    TRY(zero_ctrl_fields(a, kFieldScratch, 1));
    cosine_kernel<<<grid_for(c, a->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(a, false), view_of(b, false), c->d_f64);
    LAUNCHED();
    TableView none = view_of(a, false);
    none.slots = nullptr;  // norm-only pass over b: adds nothing to any dot product
    cosine_kernel<<<grid_for(c, b->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(b, false), none, c->d_f64 + 1);
    LAUNCHED();
    CU(cudaGetLastError());
    double sq[2];
    CU(cudaMemcpyAsync(sq, c->d_f64, 16, cudaMemcpyDeviceToHost, c->stream));
    TRY(pull_ctrl(a));
    uint64_t dot = a->h_ctrl->scratch[0];
    if (a->h_ctrl->side_present) {
        sq[0] += (double)a->h_ctrl->side_count * (double)a->h_ctrl->side_count;
        if (b->h_ctrl->side_present) dot += a->h_ctrl->side_count * b->h_ctrl->side_count;
    }
    if (b->h_ctrl->side_present) sq[1] += (double)b->h_ctrl->side_count * (double)b->h_ctrl->side_count;
    const double ma = sqrt(sq[0]), mb = sqrt(sq[1]);
    *out = (ma == 0.0 || mb == 0.0) ? 0.0 : (double)dot / (ma * mb);
    return OXG_OK;
}

oxg_status oxg_merge(oxg_table *dst, oxg_table *src, uint64_t *counts_added, uint64_t *new_keys) {
    TRY(two_tables(dst, src));
    if (dst == src) return fail(OXG_ERR_INVALID, "cannot merge a table into itself");
    if (dst->k != src->k) return fail(OXG_ERR_WRONG_KSIZE, "KmerCountTables must have the same ksize");
    ENTER(dst);
    TRY(pull_ctrl(dst));
    TRY(pull_ctrl(src));
    TRY(reserve_keys(dst, src->h_ctrl->size));
    TRY(zero_ctrl_fields(dst, kFieldScratch, 2));
    merge_kernel<<<grid_for(c, src->cap, kOpThreads, 16), kOpThreads, 0, c->stream>>>(view_of(dst, false), view_of(src, false));
    LAUNCHED();
    CU(cudaGetLastError());
    TRY(pull_ctrl(dst));
    uint64_t added = dst->h_ctrl->scratch[0], fresh = dst->h_ctrl->scratch[1];
    if (src->h_ctrl->side_present) {
        Ctrl *h = dst->h_ctrl;
        if (!h->side_present || h->side_count == 0) ++fresh;
        h->side_present = 1;
        h->side_count += src->h_ctrl->side_count;
        added += src->h_ctrl->side_count;
        CU(cudaMemcpyAsync(&dst->d_ctrl->side_present, &h->side_present, 16, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (counts_added) *counts_added = added;
    if (new_keys) *new_keys = fresh;
    return OXG_OK;
}

// ---- synthetic reads, memory helpers ------------------------------------------------

oxg_status oxg_synth_reads_device(int device, uint8_t *d_bases, uint64_t n_reads, uint32_t read_len,
                                  uint64_t genome_len, uint64_t seed, uint64_t first_read,
                                  uint32_t sub_ppm, uint32_t n_ppm) {
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->dev));
    if (!d_bases || read_len == 0 || genome_len < read_len) return fail(OXG_ERR_INVALID, "bad synth arguments");
    synth_reads_kernel<<<c->sms * 16, 256, 0, c->stream>>>(d_bases, n_reads, read_len, genome_len, seed, first_read, sub_ppm, n_ppm);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return OXG_OK;
}

oxg_status oxg_pinned_alloc(uint64_t bytes, void **out) {
    if (!out) return fail(OXG_ERR_INVALID, "null argument");
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return OXG_OK;
}
oxg_status oxg_pinned_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return OXG_OK;
}
oxg_status oxg_device_alloc(int device, uint64_t bytes, void **d_out) {
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    CU(cudaSetDevice(c->dev));
    CU(cudaMalloc(d_out, bytes ? bytes : 16));
    return OXG_OK;
}
oxg_status oxg_device_free(int device, void *d_ptr) {
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    CU(cudaSetDevice(c->dev));
    if (d_ptr) CU(cudaFree(d_ptr));
    return OXG_OK;
}
oxg_status oxg_memcpy_h2d(int device, void *d_dst, const void *src, uint64_t bytes) {
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    CU(cudaSetDevice(c->dev));
    CU(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice));
    return OXG_OK;
}
oxg_status oxg_memcpy_d2h(int device, void *dst, const void *d_src, uint64_t bytes) {
    DeviceCtx *c;
    TRY(get_ctx(device, &c));
    CU(cudaSetDevice(c->dev));
    CU(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return OXG_OK;
}

}  // extern "C"

#include "capi_shard.inc"
