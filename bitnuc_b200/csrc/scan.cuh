// scan.cuh -- exclusive prefix sums of per-item counts (u64) over n items, as two small kernels: per-CTA sums (each CTA also
// adds its sum to the total of its group of 256 CTAs), then CTA-local scan + offset, where a CTA gets the sum of everything
// before it from at most 255 block sums of its own group plus the totals of the groups before (two loads per thread and one
// block reduction).  The one-CTA scan of the block sums that used to sit between the two (37 us for the 39 K sums of 40 M
// reads: five rounds of load - scan - store by a single CTA) is gone.  out[i] = sum of count(j), j < i, for i in [0, n].
//
// Item layout inside a CTA tile of 1024 items is WARP-STRIPED: warp w owns items [128 w, 128 w + 128) of the tile and
// lane l touches items 128 w + 32 i + l, i < 4 -- every load the count functor makes and every offset store is a
// coalesced 256-byte warp transaction (a blocked layout, 4 consecutive items per thread, leaves 3/4 of every sector
// to the L1).  The warp scans its four rows one after the other, carrying the running total from row to row.
//
// One channel (count(i) -> u64) or two at once (count(i) -> ulonglong2: split_packed needs the left and right word
// counts of the same reads), and a hook that runs once per item right where its offsets are known (encode_batch
// notes the owner of every output tile there).  A single-pass chained scan with decoupled look-back was measured
// and is slower at this tile size.
#pragma once

#include "common.cuh"
#include "launch.cuh"

namespace bn {

constexpr int kScanItems = 4;                      // items per thread
constexpr int kScanTile = kThreads * kScanItems;   // items per CTA

struct ScanNoHook {
    __device__ __forceinline__ void operator()(unsigned long long, unsigned long long, unsigned long long) const {}
    __device__ __forceinline__ void operator()(unsigned long long, unsigned long long, unsigned long long, unsigned long long,
                                               unsigned long long) const {}
};

template <int NCH, typename F>
__device__ __forceinline__ void scan_counts(const F& count, unsigned long long i, unsigned long long (&v)[NCH]) {
    if constexpr (NCH == 1) {
        v[0] = count(i);
    } else {
        const ulonglong2 c = count(i);
        v[0] = c.x;
        v[1] = c.y;
    }
}

// item index of (row i, this thread) in the warp-striped tile layout
__device__ __forceinline__ unsigned long long scan_item(unsigned i) {
    return (unsigned long long)blockIdx.x * kScanTile + (threadIdx.x >> 5) * (32 * kScanItems) + 32 * i + (threadIdx.x & 31);
}

constexpr int kScanGroup = 256;   // CTAs per group (= kThreads: a thread fetches one block sum of its group and one group total)
static_assert(kScanGroup == kThreads, "one block sum per thread");
// per channel: block sums [n_blocks] | group totals [ceil(n_blocks / kScanGroup)] (zeroed before the first kernel)
__host__ __device__ __forceinline__ unsigned long long scan_groups(unsigned long long n_blocks) { return (n_blocks + kScanGroup - 1) / kScanGroup; }
__host__ __device__ __forceinline__ unsigned long long scan_channel_words(unsigned long long n_blocks) { return n_blocks + scan_groups(n_blocks); }

// sums[ch * scan_channel_words + b] = sum of channel ch over tile b; the group totals behind them
template <int NCH, typename F>
__global__ void __launch_bounds__(kThreads)
scan_block_sums_kernel(F count, unsigned long long n, unsigned long long n_blocks, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long scratch[32];
    unsigned long long s[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) s[ch] = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const unsigned long long r = scan_item(i);
        if (r < n) {
            unsigned long long c[NCH];
            scan_counts<NCH>(count, r, c);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) s[ch] += c[ch];
        }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const unsigned long long t = block_sum_u64(s[ch], scratch);
        if (threadIdx.x == 0) {
            unsigned long long* ch_sums = sums + ch * scan_channel_words(n_blocks);
            ch_sums[blockIdx.x] = t;
            if (t) atomicAdd(ch_sums + n_blocks + blockIdx.x / kScanGroup, t);
        }
    }
}

// hook(i, out[i], count(i)) -- or hook(i, out_a[i], a(i), out_b[i], b(i)) with two channels -- once per item
template <int NCH, typename F, typename H>
__global__ void __launch_bounds__(kThreads)
scan_offsets_kernel(F count, unsigned long long n, const unsigned long long* __restrict__ sums, unsigned long long n_blocks,
                    uint64_t* __restrict__ out_a, uint64_t* __restrict__ out_b, H hook) {
    __shared__ unsigned long long warp_tot[NCH][32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long warp_before[NCH][kWarpsPerBlock], tile_before[NCH];
    unsigned long long c[kScanItems][NCH], excl[kScanItems][NCH], carry[NCH], before[NCH];
    // everything before this tile: thread t takes the block sum of tile (first tile of the group) + t if that tile precedes
    // this one, and the total of group t if that group precedes this one's -- asked for with the items, reduced below
    const unsigned long long group = blockIdx.x / kScanGroup, in_group = blockIdx.x % kScanGroup;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const unsigned long long* ch_sums = sums + ch * scan_channel_words(n_blocks);
        carry[ch] = 0;
        before[ch] = threadIdx.x < in_group ? ch_sums[group * kScanGroup + threadIdx.x] : 0ull;
        for (unsigned long long g = threadIdx.x; g < group; g += kThreads) before[ch] += ch_sums[n_blocks + g];
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {  // all loads first (independent), the row scans below
        const unsigned long long r = scan_item(i);
        if (r < n) {
            scan_counts<NCH>(count, r, c[i]);
        } else {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) c[i][ch] = 0;
        }
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            unsigned long long inc = c[i][ch];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += t;
            }
            excl[i][ch] = carry[ch] + inc - c[i][ch];                   // exclusive prefix inside the warp's 128 items
            carry[ch] += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const unsigned long long b = warp_sum_u64(before[ch]);   // rides on the two barriers the tile scan needs anyway
        if (lane == 0) warp_tot[ch][warp] = carry[ch], warp_before[ch][warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const unsigned long long wb = warp_sum_u64(lane < kWarpsPerBlock ? warp_before[ch][lane] : 0ull);
            if (lane == 0) tile_before[ch] = wb;
            const unsigned long long w = lane < kWarpsPerBlock ? warp_tot[ch][lane] : 0;
            unsigned long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (unsigned)o) winc += t;
            }
            if (lane < kWarpsPerBlock) warp_tot[ch][lane] = winc - w;   // exclusive prefix of the warp totals
        }
    }
    __syncthreads();
    unsigned long long base[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) base[ch] = tile_before[ch] + warp_tot[ch][warp];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const unsigned long long r = scan_item(i);
        if (r < n) {
            out_a[r] = base[0] + excl[i][0];
            if constexpr (NCH == 2) {
                out_b[r] = base[1] + excl[i][1];
                hook(r, base[0] + excl[i][0], c[i][0], base[1] + excl[i][1], c[i][1]);
            } else {
                hook(r, base[0] + excl[i][0], c[i][0]);
            }
        }
    }
    if (blockIdx.x == n_blocks - 1 && threadIdx.x == kThreads - 1) {   // the tile's last item slot: everything before it + itself = the total
        out_a[n] = base[0] + excl[kScanItems - 1][0] + c[kScanItems - 1][0];
        if constexpr (NCH == 2) out_b[n] = base[1] + excl[kScanItems - 1][1] + c[kScanItems - 1][1];
    }
}

// bytes of scratch (`sums`) the scans need for n items
static inline size_t scan_scratch_bytes(size_t n) { return (size_t)scan_channel_words(ceil_div(n ? n : 1, kScanTile)) * sizeof(unsigned long long); }
static inline size_t scan2_scratch_bytes(size_t n) { return 2 * scan_scratch_bytes(n); }

template <int NCH, typename F, typename H>
static void launch_scan_impl(F count, size_t n, unsigned long long* sums, uint64_t* out_a, uint64_t* out_b, cudaStream_t s, H hook) {
    const unsigned long long n_blocks = ceil_div(n, kScanTile);
    for (int ch = 0; ch < NCH; ++ch)   // the group totals are accumulated with atomics
        cudaMemsetAsync(sums + ch * scan_channel_words(n_blocks) + n_blocks, 0, scan_groups(n_blocks) * sizeof(unsigned long long), s);
    scan_block_sums_kernel<NCH><<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, n_blocks, sums);
    scan_offsets_kernel<NCH><<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, sums, n_blocks, out_a, out_b, hook);
}

// out[i] = sum of count(j), j < i, for i in [0, n]; n >= 1
template <typename F, typename H = ScanNoHook>
static void launch_exclusive_scan(F count, size_t n, unsigned long long* sums, uint64_t* out, cudaStream_t s, H hook = H()) {
    launch_scan_impl<1>(count, n, sums, out, nullptr, s, hook);
}

// The same scan for a count that is expensive to evaluate (a dependent random load): the first pass stores every count
// in `cache` (n entries), the last pass reads it back from there instead of evaluating the functor a second time.
template <typename F>
struct ScanStoreCount {
    F count;
    uint64_t* cache;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long i) const {
        const unsigned long long c = count(i);
        cache[i] = c;
        return c;
    }
};
struct ScanLoadCount {
    const uint64_t* cache;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long i) const { return cache[i]; }
};
template <typename F>
static void launch_exclusive_scan_cached(F count, size_t n, unsigned long long* sums, uint64_t* cache, uint64_t* out, cudaStream_t s) {
    const unsigned long long n_blocks = ceil_div(n, kScanTile);
    cudaMemsetAsync(sums + n_blocks, 0, scan_groups(n_blocks) * sizeof(unsigned long long), s);
    scan_block_sums_kernel<1><<<(unsigned)n_blocks, kThreads, 0, s>>>(ScanStoreCount<F>{count, cache}, n, n_blocks, sums);
    scan_offsets_kernel<1><<<(unsigned)n_blocks, kThreads, 0, s>>>(ScanLoadCount{cache}, n, sums, n_blocks, out, nullptr, ScanNoHook());
}

// two channels: count(i) returns (a, b) as a ulonglong2
template <typename F, typename H = ScanNoHook>
static void launch_exclusive_scan2(F count, size_t n, unsigned long long* sums, uint64_t* out_a, uint64_t* out_b, cudaStream_t s,
                                   H hook = H()) {
    launch_scan_impl<2>(count, n, sums, out_a, out_b, s, hook);
}

}  // namespace bn
