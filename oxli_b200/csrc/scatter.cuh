// scatter.cuh -- pass A of the partitioned pipeline: hash the reads, scatter every hash into
// the fragment of its destination (see aggregate.cuh for pass B and for why).
//
// Replaces the first half of the hot loop of KmerCountTable::consume
// (/root/reference/src/lib.rs:576-600: the SeqToHashes step); the `count_hash` half happens in
// pass B.  Validity, canonical strand and hashing are the device functions of consume.cuh.
//
// What shapes this kernel is the store side.  A scattered store -- one lane, one sector --
// costs a B200 SM about seven cycles whatever its width (8, 16 or 32 bytes: 40 G stores/s
// chip-wide, profiles/r2_microbench_smem_scatter.txt), so "one 8-byte store per hash straight
// into the destination's fragment" runs at a third of the hashing rate.  Shared-memory atomics
// and stores, in contrast, cost 0.13-0.45 cycles.  So every destination has a line of
// 2^line_shift entries (128 bytes at 16) staged in shared memory; hashes are ranked into it
// with one shared-memory atomic each and a line leaves the SM as ONE full-width coalesced
// store when it is complete.
//
// A CTA of 12 warps (two per SM) works in rounds: every warp hashes one tile of 256 window starts
// (8 per lane), then the CTA
//   1. ranks:   q = atomicAdd(arrivals[dest], 1) is the hash's position in the fragment | barrier
//   2. places:  the line that was being filled is completed in shared memory (whoever takes its
//               last slot lists the destination), entries of further complete lines (rare) go
//               straight out, what lies beyond the last complete line waits       | barrier
//   3. flushes: listed lines, 32 bytes per lane                                   | barrier
//   4. moves the waiting entries in at the front of their line; whoever took a destination's
//      last position of the round brings its book up to date       (no barrier: step 1 of the
//      next round touches other words, and its barrier precedes every reader)
// There is no per-destination loop anywhere: all bookkeeping is done by the threads that hold
// the hashes.
// At the end of a launch the last line of every destination goes out padded with zeros (0 is
// never a hash here, src/lib.rs:589), so fill counts stay multiples of the line and a later
// launch can continue the same fragments (frag_append): pass B then runs once per several
// launches and sees more duplicates per key.
// A fragment is [dest][cta][frag_cap] as pass B expects; what does not fit a fragment (skew: one
// k-mer flooding its partition) goes to the spill list exactly as before.
#pragma once
#include "consume.cuh"

namespace oxg {

// Two CTAs of 12 warps per SM rather than one of 24: the phases between barriers are short and
// serial in nature, and while one CTA is in them the other one hashes.
#ifndef OXG_SCAT_THREADS
#define OXG_SCAT_THREADS 384
#endif
#ifndef OXG_SCAT_CTAS
#define OXG_SCAT_CTAS 2
#endif
constexpr int kScatThreads = OXG_SCAT_THREADS;
constexpr int kScatCtasPerSm = OXG_SCAT_CTAS;
constexpr int kScatWarps = kScatThreads / 32;

// per-warp tile buffers: forward bytes, mirrored complement, bad bits, end bits
template <int K>
struct ScatWarpBuf {
    static constexpr int BL = TileGeom<K>::BL;
    static constexpr int kBytes = 2 * BL + 64 + 64;
};

template <int K>
inline size_t scatter_smem_bytes(uint32_t n_dest, uint32_t line_shift) {
    return (((size_t)n_dest * ((8u << line_shift) + 12) + 16 + 15) & ~(size_t)15) + (size_t)kScatWarps * ScatWarpBuf<K>::kBytes;
}

template <int K>
__global__ void __launch_bounds__(kScatThreads, kScatCtasPerSm) scatter_kernel(const ConsumeParams p) {
    using G = TileGeom<K>;
    constexpr int BL = G::BL, NV = G::NV, NE = G::NE;
    static_assert(NV + 8 <= 32 && NE <= 16, "per-warp mask buffers are 64 bytes each");
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t nd = p.n_dest, ls = p.line_shift, line = 1u << ls;
    uint64_t *stage = reinterpret_cast<uint64_t *>(sm);                       // [nd][line]
    uint2 *book = reinterpret_cast<uint2 *>(sm + ((size_t)nd << (ls + 3)));  // [nd]
    uint32_t *flist = reinterpret_cast<uint32_t *>(book + nd);                // [nd] destinations to flush this round
    uint32_t *s_nflush = flist + nd;                                          // [4]
    uint8_t *wbuf = sm + ((((size_t)nd * ((8u << ls) + 12) + 16) + 15) & ~(size_t)15) + (size_t)(threadIdx.x >> 5) * ScatWarpBuf<K>::kBytes;
    uint8_t *s_fw = wbuf, *s_rc = wbuf + BL;
    uint16_t *s_bad = reinterpret_cast<uint16_t *>(wbuf + 2 * BL);
    uint32_t *s_end = reinterpret_cast<uint32_t *>(wbuf + 2 * BL + 64);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t frag_row = (uint64_t)gridDim.x * p.frag_cap;              // entries between destinations
    uint64_t *const my_frag = p.frag + (uint64_t)blockIdx.x * p.frag_cap;    // + dest * frag_row
    // book[d] = {x: entries of the fragment accounted for at the start of the round,
    //            y: arrivals ever (never reset: an arrival's rank IS its position in the fragment)}
    // A launch may continue fragments an earlier launch began (frag_append): their fill counts are
    // multiples of the line, because every launch pads its last lines with zeros (pass B skips them).
    for (uint32_t i = threadIdx.x; i < nd; i += kScatThreads) {
        const uint32_t at = p.frag_append ? p.frag_cnt[(uint64_t)i * gridDim.x + blockIdx.x] : 0u;
        book[i] = make_uint2(at, at);
    }
    if (threadIdx.x == 0) *s_nflush = 0;
    __syncthreads();

    auto dest_of = [&](uint64_t h) {
        uint32_t d = (uint32_t)((h * kPhi) >> p.part_shift);
        if (p.n_ranks > 1) d += (uint32_t)(h >> p.owner_shift) * p.n_parts;
        return d;
    };

    uint64_t n_counted = 0, first_bad_unused = ~0ULL;
    const uint64_t tiles_per_round = (uint64_t)gridDim.x * kScatWarps;
    const uint64_t n_rounds = (p.n_tiles + tiles_per_round - 1) / tiles_per_round;
    auto tile_of = [&](uint64_t r) { return (r * gridDim.x + blockIdx.x) * kScatWarps + warp; };

    // software pipeline over rounds: the bytes and read boundaries of the next tile are fetched
    // into registers while this one is hashed and scattered
    uint4 raw = make_uint4(0, 0, 0, 0);
    uint64_t tf = 0, off = ~0ULL;
    auto fetch = [&](uint64_t t, uint4 &raw_, uint64_t &tf_) {
        if (t < p.n_tiles) {
            if (lane < NV) raw_ = load_bases16(p, p.tile_base + t * kWarpTile + 16ull * lane);
            tf_ = __ldg(p.tile_first + t);
        }
    };
    auto fetch_off = [&](uint64_t t, uint64_t tf_) {
        const uint64_t rr = tf_ + lane;
        return (t < p.n_tiles && rr < p.n_off) ? __ldg(p.offsets + rr) : ~0ULL;
    };
    fetch(tile_of(0), raw, tf);
    off = fetch_off(tile_of(0), tf);

    // a line leaves as 2^(ls-2) lanes x 4 entries: two 128-bit shared loads, one 256-bit store
    const uint32_t lanes_per_line = ls >= 2 ? (line >> 2) : 1u;

    for (uint64_t r = 0; r < n_rounds; ++r) {
        const uint64_t t = tile_of(r);
        uint4 raw_next = make_uint4(0, 0, 0, 0);
        uint64_t tf_next = 0;
        fetch(tile_of(r + 1), raw_next, tf_next);
        uint64_t h[kWPT] = {};
        if (t < p.n_tiles) {
            const uint64_t w0 = p.tile_base + t * kWarpTile;
            if (lane < NE) s_end[lane] = 0;
            __syncwarp();
            if (lane < NV) stage16_raw<BL>(p, w0, lane, raw, s_fw, s_rc, s_bad);
            {
                const uint64_t e = off - 1 - w0;  // last base of a read, tile-relative (sentinel: huge)
                if (e < (uint64_t)BL) atomicOr(&s_end[e >> 5], 1u << (e & 31));
                if (__shfl_sync(0xffffffffu, e < (uint64_t)BL, 31)) {
                    // more than 32 read boundaries inside one tile (tiny or empty reads): walk the rest
                    for (uint64_t q = tf + 32 + lane; q < p.n_off; q += 32) {
                        const uint64_t e2 = p.offsets[q] - 1 - w0;
                        if (e2 >= (uint64_t)BL) break;
                        atomicOr(&s_end[e2 >> 5], 1u << (e2 & 31));
                    }
                }
            }
            __syncwarp();
            const uint32_t valid = lane_valid_mask<K, kModePart>(p, s_bad, s_end, w0, lane, first_bad_unused);
            if (valid) lane_hashes<K>(s_fw, s_rc, lane * kWPT, valid, h);
        }
        const uint64_t off_next = fetch_off(tile_of(r + 1), tf_next);  // its tile_first entry has arrived by now

        // 1. every hash takes its position q in its destination's fragment
        uint32_t dst[kWPT], q[kWPT];
#pragma unroll
        for (int j = 0; j < kWPT; ++j) {
            dst[j] = 0; q[j] = 0;
            if (h[j] != 0) { ++n_counted; dst[j] = dest_of(h[j]); q[j] = atomicAdd(&book[dst[j]].y, 1u); }
        }
        __syncthreads();

        // 2. place.  With b.x = entries accounted for before this round and total = what the
        //    fragment holds after it: positions inside the line that was being filled go into the
        //    staged line (whoever takes its last slot lists the destination for the flush);
        //    positions in further complete lines (rare) go straight out; positions beyond the last
        //    complete line wait for the flush; positions beyond the fragment spill.
        uint32_t later = 0, spilled = 0, last = 0;
        uint2 bk[kWPT];  // all books first: the loads are independent, the stores below are not
#pragma unroll
        for (int j = 0; j < kWPT; ++j) bk[j] = book[dst[j]];
#pragma unroll
        for (int j = 0; j < kWPT; ++j) {
            if (h[j] == 0) continue;
            const uint2 b = bk[j];
            const uint32_t total = min(b.y, p.frag_cap), line_end = (b.x & ~(line - 1)) + line;
            if (q[j] >= total) { spilled |= 1u << j; continue; }
            if (q[j] + 1 == total) last |= 1u << j;
            if (q[j] < line_end) {
                stage[((size_t)dst[j] << ls) + (q[j] & (line - 1))] = h[j];
                if (q[j] + 1 == line_end) flist[atomicAdd(s_nflush, 1u)] = dst[j];
            } else if (q[j] < (total & ~(line - 1))) my_frag[dst[j] * frag_row + q[j]] = h[j];
            else later |= 1u << j;
        }
        if (__any_sync(0xffffffffu, spilled != 0)) {
            // skewed input (one k-mer flooding its partition): one reservation per warp tile
            const uint32_t mine = __popc(spilled);
            uint32_t incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.spill_n, (unsigned long long)tot);
            base = __shfl_sync(0xffffffffu, base, 0) + (incl - mine);
#pragma unroll
            for (int j = 0; j < kWPT; ++j)
                if ((spilled >> j) & 1u) { if (base < p.spill_cap) p.spill[base] = h[j]; ++base; }
        }
        __syncthreads();

        // 3. flush the completed lines
        {
            const uint32_t n_flush = *s_nflush;
            if (ls >= 2) {
                const uint32_t per_warp = 32u / lanes_per_line, sub = lane / lanes_per_line, li = (lane % lanes_per_line) * 4;
                for (uint32_t e = warp * per_warp + sub; e < n_flush; e += kScatWarps * per_warp) {
                    const uint32_t d = flist[e];
                    const uint64_t *src = stage + ((size_t)d << ls) + li;
                    const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(src), v1 = *reinterpret_cast<const ulonglong2 *>(src + 2);
                    uint64_t *dstp = my_frag + d * frag_row + (book[d].x & ~(line - 1)) + li;
                    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(dstp), "l"(v0.x), "l"(v0.y), "l"(v1.x), "l"(v1.y) : "memory");
                }
            } else {
                const uint32_t per_warp = 32u >> ls, sub = lane >> ls, li = lane & (line - 1);
                for (uint32_t e = warp * per_warp + sub; e < n_flush; e += kScatWarps * per_warp) {
                    const uint32_t d = flist[e];
                    my_frag[d * frag_row + (book[d].x & ~(line - 1)) + li] = stage[((size_t)d << ls) + li];
                }
            }
        }
        __syncthreads();

        // 4. what lies beyond the flushed lines moves in at the front of its line; whoever took a
        //    destination's last position of the round brings its book up to date
#pragma unroll
        for (int j = 0; j < kWPT; ++j) {
            if ((later >> j) & 1u) stage[((size_t)dst[j] << ls) + (q[j] & (line - 1))] = h[j];
            if ((last >> j) & 1u) book[dst[j]].x = q[j] + 1;
        }
        if (threadIdx.x == 0) *s_nflush = 0;
        // (no barrier here: the next round's step 1 touches only the arrival counters, and its
        // barrier comes before anything reads what step 4 wrote)
        raw = raw_next; tf = tf_next; off = off_next;
    }

    __syncthreads();
    // what is still staged: the last line of every destination goes out whole, its unused slots
    // holding 0 (never a hash: pass A drops it, src/lib.rs:589) so that fill counts stay multiples
    // of the line and a later launch can continue the fragment
    {
        const uint32_t per_warp = 32u >> ls, sub = lane >> ls, li = lane & (line - 1);
        for (uint32_t d = warp * per_warp + sub; d < nd; d += kScatWarps * per_warp) {
            const uint32_t pos = book[d].x, part = pos & (line - 1);
            if (part) my_frag[d * frag_row + (pos - part) + li] = li < part ? stage[((size_t)d << ls) + li] : 0ull;
        }
        for (uint32_t d = threadIdx.x; d < nd; d += kScatThreads)
            p.frag_cnt[(uint64_t)d * gridDim.x + blockIdx.x] = min((book[d].x + line - 1) & ~(line - 1), p.frag_cap);
    }
    for (int o = 16; o; o >>= 1) n_counted += __shfl_xor_sync(0xffffffffu, n_counted, o);
    if (lane == 0 && n_counted) atomicAdd((unsigned long long *)&p.table.ctrl->counted, (unsigned long long)n_counted);
}

}  // namespace oxg
