// lookback.cuh -- single-pass prefix sums across CTA tiles (decoupled look-back), for kernels that need "the total of
// every tile before mine" in the same launch that moves the data (split_packed, slice gathers): the three-launch scan of
// scan.cuh reads every count twice and parks the offsets in HBM between launches; here a tile's counts are read once.
//
// One 64-bit descriptor per tile and channel: bits 63:62 = state (0 not yet, 1 the tile's own aggregate, 2 the inclusive
// prefix up to and including the tile), bits 61:0 = the value.  State and value travel in ONE word, so a plain 64-bit
// store publishes them together and no fence is needed; the descriptors must be zero before the launch.  Tiles are
// numbered by a ticket counter in the order CTAs start running, so every tile a CTA waits for is already resident (or
// done): the spin below cannot deadlock whatever the grid size.
//
// The look-back is done by one warp: lane l reads the descriptor of tile (tile - 1 - l); the nearest tile that already
// published an inclusive prefix ends the walk, the aggregates in between are added up; 32 tiles per step.
#pragma once

#include "common.cuh"

namespace bn {

constexpr unsigned long long kLbAggregate = 1ull << 62, kLbPrefix = 2ull << 62, kLbValue = (1ull << 62) - 1ull;

__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// bytes of descriptors for n_tiles tiles of NCH channels (+ the ticket counter in front)
static inline size_t lookback_bytes(size_t n_tiles, int nch) { return (1 + n_tiles * (size_t)nch) * sizeof(unsigned long long); }

// Call with ALL 32 lanes of one warp.  agg[ch] = this tile's total of channel ch (the same in every lane).
// Returns in excl[ch] (every lane) the sum of the aggregates of tiles 0 .. tile-1, and publishes this tile's
// inclusive prefix.  desc = descriptors of channel 0, channel ch at desc + ch * n_tiles.
template <int NCH>
__device__ __forceinline__ void lookback_exclusive(unsigned long long* __restrict__ desc, unsigned long long n_tiles, unsigned long long tile,
                                                   const unsigned long long (&agg)[NCH], unsigned long long (&excl)[NCH]) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) excl[ch] = 0;
    unsigned long long mine = agg[0];   // channel `lane`'s aggregate, selected without indexing (the array stays in registers)
#pragma unroll
    for (int ch = 1; ch < NCH; ++ch) mine = lane == (unsigned)ch ? agg[ch] : mine;
    if (tile == 0) {
        if (lane < NCH) lb_store(desc + lane * n_tiles, kLbPrefix | mine);
        return;
    }
    if (lane < NCH) lb_store(desc + lane * n_tiles + tile, kLbAggregate | mine);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const unsigned long long* d = desc + ch * n_tiles;
        unsigned long long running = 0;
        long long look = (long long)tile - 1;   // lane 0 looks here, lane l at look - l
        for (;;) {
            const long long idx = look - (long long)lane;
            unsigned long long v = idx >= 0 ? lb_load(d + idx) : kLbPrefix;   // before tile 0: an inclusive prefix of 0
            while (__any_sync(0xffffffffu, (v >> 62) == 0)) {
                if ((v >> 62) == 0) v = lb_load(d + idx);
            }
            const unsigned pmask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
            const unsigned upto = pmask ? (unsigned)__ffs(pmask) - 1u : 31u;   // lanes 0 .. upto take part
            unsigned long long part = lane <= upto ? (v & kLbValue) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            running += part;
            if (pmask) break;
            look -= 32;
        }
        excl[ch] = running;
        if (lane == 0) lb_store(desc + ch * n_tiles + tile, kLbPrefix | ((running + agg[ch]) & kLbValue));
    }
}

}  // namespace bn
