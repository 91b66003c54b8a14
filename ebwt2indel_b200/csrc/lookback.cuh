// lookback.cuh -- single-pass ordered compaction across CTAs (decoupled look-back).
//
// Every frontier sweep appends a variable number of children per tile to NCH output queues and
// must keep them in tile order, so that each queue stays sorted by suffix-array position (the
// property that turns the traversal's gathers into near-sequential sweeps over the index).
// Tiles are handed out by an atomic ticket, so a tile's predecessors have always started.
//
// Descriptor word (one per tile and channel, self-contained so no fence is needed):
//   [63:48] launch epoch   [47:46] status (1 = tile aggregate, 2 = inclusive prefix)   [45:0] value
// A stale epoch reads as "not yet published"; the host bumps the epoch per launch and clears the
// array only when the 16-bit epoch wraps.
#pragma once

#include <cstdint>

namespace e2i {

constexpr unsigned long long kLbValueMask = (1ull << 46) - 1;
constexpr unsigned kLbAgg = 1, kLbIncl = 2;

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, unsigned status, unsigned long long v) {
    return ((unsigned long long)(epoch & 0xffffu) << 48) | ((unsigned long long)status << 46) | (v & kLbValueMask);
}

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// Called by ONE full warp of the CTA that owns `tile`.  agg[c] is the tile's own count for channel
// c (uniform across the warp).  Returns in excl[c] the sum of agg over all tiles < tile.
template <int NCH>
__device__ __forceinline__ void lookback_exclusive(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                                   const unsigned long long (&agg)[NCH],
                                                   unsigned long long (&excl)[NCH]) {
    const int lane = threadIdx.x & 31;
    if (lane < NCH) {
        unsigned long long mine = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) if (lane == c) mine = agg[c];
        lb_store(desc + (size_t)tile * NCH + lane, lb_pack(epoch, tile == 0 ? kLbIncl : kLbAgg, mine));
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) excl[c] = 0;
    if (tile == 0) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        unsigned long long running = 0;
        long long look = (long long)tile - 1;  // lane l inspects tile look - l
        while (true) {
            const long long t = look - lane;
            unsigned status = kLbIncl;
            unsigned long long val = 0;
            if (t >= 0) {
                const unsigned long long d = lb_load(desc + (size_t)t * NCH + c);
                status = ((uint32_t)(d >> 48) == (epoch & 0xffffu)) ? (unsigned)((d >> 46) & 3u) : 0u;
                val = d & kLbValueMask;
            }
            const unsigned incl = __ballot_sync(0xffffffffu, status == kLbIncl);
            const unsigned inval = __ballot_sync(0xffffffffu, status == 0u);
            const int first_incl = incl ? (__ffs(incl) - 1) : 32;
            const unsigned need = first_incl >= 31 ? 0xffffffffu : ((2u << first_incl) - 1u);
            if (inval & need) continue;  // a predecessor we depend on has not published yet
            unsigned long long part = ((need >> lane) & 1u) ? val : 0ull;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
            running += part;
            if (first_incl < 32) break;
            look -= 32;
        }
        excl[c] = running;
    }
    if (lane < NCH) {
        unsigned long long mine = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) if (lane == c) mine = excl[c] + agg[c];
        lb_store(desc + (size_t)tile * NCH + lane, lb_pack(epoch, kLbIncl, mine));
    }
}

}  // namespace e2i

namespace e2i {

// ---------------------------------------------------------------------------------------------
// Four channels in ONE look-back round (tiles of <= 2047 items per channel).
// Descriptor = 8 words per tile:
//   [0]    [63:48] epoch  [47:46] status (1 = aggregates valid, 2 = inclusive prefixes valid)
//          [43:0]  the tile's four 11-bit aggregates
//   [1..4] inclusive prefix of channel c: [63:48] epoch, [47:0] value (written before status 2)
// A reader that sees status 2 re-checks the epoch of the prefix words, so no acquire is needed.
// Called by the full warp 0 of the CTA that owns `tile`; agg[] is uniform across the warp.
// ---------------------------------------------------------------------------------------------
constexpr int kLb4Words = 8;

__device__ __forceinline__ void lookback4(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                          const uint32_t (&agg)[4], unsigned long long (&excl)[4]) {
    const int lane = threadIdx.x & 31;
    const unsigned long long ep = (unsigned long long)(epoch & 0xffffu) << 48;
    unsigned long long *mine = desc + (size_t)tile * kLb4Words;
    const unsigned long long packed = (unsigned long long)agg[0] | ((unsigned long long)agg[1] << 11) |
                                      ((unsigned long long)agg[2] << 22) | ((unsigned long long)agg[3] << 33);
#pragma unroll
    for (int c = 0; c < 4; ++c) excl[c] = 0;
    if (tile != 0) {
        if (lane == 0) lb_store(mine, ep | (1ull << 46) | packed);
        unsigned long long run[4] = {0, 0, 0, 0};
        long long look = (long long)tile - 1;   // lane l inspects tile look - l
        while (true) {
            const long long t = look - lane;
            unsigned status = kLbIncl;          // tiles before 0: an inclusive prefix of zero
            unsigned long long a = 0;
            if (t >= 0) {
                const unsigned long long d = lb_load(desc + (size_t)t * kLb4Words);
                status = ((uint32_t)(d >> 48) == (epoch & 0xffffu)) ? (unsigned)((d >> 46) & 3u) : 0u;
                a = d;
            }
            const unsigned incl = __ballot_sync(0xffffffffu, status == kLbIncl);
            const unsigned inval = __ballot_sync(0xffffffffu, status == 0u);
            const int first_incl = incl ? (__ffs(incl) - 1) : 32;
            const unsigned need = first_incl >= 32 ? 0xffffffffu : ((1u << first_incl) - 1u);
            if (inval & need) continue;         // a predecessor we depend on has not published yet
            const bool contrib = (need >> lane) & 1u;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                run[c] += __reduce_add_sync(0xffffffffu, contrib ? (uint32_t)((a >> (11 * c)) & 0x7ffu) : 0u);
            if (first_incl < 32) {
                const long long ti = look - first_incl;
                unsigned long long v = 0;
                if (ti >= 0 && lane < 4) {
                    const unsigned long long *p = desc + (size_t)ti * kLb4Words + 1 + lane;
                    do { v = lb_load(p); } while ((v >> 48) != (ep >> 48));
                    v &= (1ull << 48) - 1;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) excl[c] = __shfl_sync(0xffffffffu, v, c) + run[c];
                break;
            }
            look -= 32;
        }
    }
    if (lane < 4) {
        unsigned long long mine_incl = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (lane == c) mine_incl = excl[c] + agg[c];
        lb_store(mine + 1 + lane, ep | mine_incl);
        __threadfence();
    }
    __syncwarp();
    if (lane == 0) lb_store(mine, ep | (2ull << 46) | packed);
}

}  // namespace e2i
