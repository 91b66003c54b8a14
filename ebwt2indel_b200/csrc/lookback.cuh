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
