// scan.cuh -- exclusive prefix sum of a per-item count over n items (u64), as three small kernels:
// per-CTA sums, a one-CTA scan of the sums, CTA-local scan + offset.  out[n] receives the total.
#pragma once

#include "common.cuh"
#include "launch.cuh"

namespace bn {

constexpr int kScanItems = 4;                      // items per thread
constexpr int kScanTile = kThreads * kScanItems;   // items per CTA

template <typename F>
__global__ void __launch_bounds__(kThreads)
scan_block_sums_kernel(F count, unsigned long long n, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long scratch[32];
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (r0 + i < n) s += count(r0 + i);
    s = block_sum_u64(s, scratch);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

// exclusive scan of sums[0..n) in place by one CTA; sums[n] = total
static __global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned long long* __restrict__ sums, unsigned long long n) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long base = 0; base < n; base += blockDim.x) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned long long v = i < n ? sums[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (unsigned)o) winc += t;
            }
            warp_tot[lane] = winc - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (i < n) sums[i] = carry + warp_tot[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[warp] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n] = carry_s;
}

struct ScanNoHook {
    __device__ __forceinline__ void operator()(unsigned long long, unsigned long long, unsigned long long) const {}
};

// hook(i, out[i], count(i)) is called once per item (e.g. to note which item owns a given output position)
template <typename F, typename H>
__global__ void __launch_bounds__(kThreads)
scan_offsets_kernel(F count, unsigned long long n, const unsigned long long* __restrict__ sums, unsigned long long n_blocks,
                    uint64_t* __restrict__ out, H hook) {
    __shared__ unsigned long long warp_tot[32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long c[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        c[i] = r0 + i < n ? count(r0 + i) : 0;
        s += c[i];
    }
    unsigned long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < kWarpsPerBlock ? warp_tot[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        if (lane < kWarpsPerBlock) warp_tot[lane] = winc - w;
    }
    __syncthreads();
    unsigned long long run = sums[blockIdx.x] + warp_tot[warp] + inc - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (r0 + i < n) {
            out[r0 + i] = run;
            hook(r0 + i, run, c[i]);
        }
        run += c[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[n_blocks];
}

// bytes of scratch (`sums`) the scan needs for n items
static inline size_t scan_scratch_bytes(size_t n) { return (ceil_div(n ? n : 1, kScanTile) + 1) * sizeof(unsigned long long); }

// out[i] = sum of count(j), j < i, for i in [0, n]; n >= 1
template <typename F, typename H = ScanNoHook>
static void launch_exclusive_scan(F count, size_t n, unsigned long long* sums, uint64_t* out, cudaStream_t s, H hook = H()) {
    const unsigned long long n_blocks = ceil_div(n, kScanTile);
    scan_block_sums_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, sums);
    scan_sums_kernel<<<1, 1024, 0, s>>>(sums, n_blocks);
    scan_offsets_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, sums, n_blocks, out, hook);
}

// ---- two channels at once: count(i) returns (a, b); out_a / out_b receive the two exclusive prefix sums.  One pass
// over the items instead of two when both sums come from the same inputs (split_packed: left and right word counts).
template <typename F>
__global__ void __launch_bounds__(kThreads)
scan2_block_sums_kernel(F count, unsigned long long n, unsigned long long* __restrict__ sums_a, unsigned long long* __restrict__ sums_b) {
    __shared__ unsigned long long scratch[32];
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (r0 + i < n) {
            const ulonglong2 c = count(r0 + i);
            a += c.x;
            b += c.y;
        }
    a = block_sum_u64(a, scratch);
    b = block_sum_u64(b, scratch);
    if (threadIdx.x == 0) {
        sums_a[blockIdx.x] = a;
        sums_b[blockIdx.x] = b;
    }
}

// one CTA per channel
static __global__ void __launch_bounds__(1024) scan2_sums_kernel(unsigned long long* __restrict__ sums_a, unsigned long long* __restrict__ sums_b,
                                                                 unsigned long long n) {
    unsigned long long* sums = blockIdx.x ? sums_b : sums_a;
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long base = 0; base < n; base += blockDim.x) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned long long v = i < n ? sums[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (unsigned)o) winc += t;
            }
            warp_tot[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (i < n) sums[i] = carry + warp_tot[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[warp] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n] = carry_s;
}

// hook(i, out_a[i], a(i), out_b[i], b(i)) is called once per item, right where its offsets are known
template <typename F, typename H>
__global__ void __launch_bounds__(kThreads)
scan2_offsets_kernel(F count, unsigned long long n, const unsigned long long* __restrict__ sums_a, const unsigned long long* __restrict__ sums_b,
                     unsigned long long n_blocks, uint64_t* __restrict__ out_a, uint64_t* __restrict__ out_b, H hook) {
    __shared__ unsigned long long warp_a[32], warp_b[32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    ulonglong2 c[kScanItems];
    unsigned long long sa = 0, sb = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        c[i] = r0 + i < n ? count(r0 + i) : make_ulonglong2(0, 0);
        sa += c[i].x;
        sb += c[i].y;
    }
    unsigned long long ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= (unsigned)o) ia += ta, ib += tb;
    }
    if (lane == 31) warp_a[warp] = ia, warp_b[warp] = ib;
    __syncthreads();
    if (warp == 0) {
        unsigned long long wa = lane < kWarpsPerBlock ? warp_a[lane] : 0, wb = lane < kWarpsPerBlock ? warp_b[lane] : 0, xa = wa, xb = wb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
            if (lane >= (unsigned)o) xa += ta, xb += tb;
        }
        if (lane < kWarpsPerBlock) warp_a[lane] = xa - wa, warp_b[lane] = xb - wb;
    }
    __syncthreads();
    unsigned long long ra = sums_a[blockIdx.x] + warp_a[warp] + ia - sa, rb = sums_b[blockIdx.x] + warp_b[warp] + ib - sb;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (r0 + i < n) {
            out_a[r0 + i] = ra;
            out_b[r0 + i] = rb;
            hook(r0 + i, ra, c[i].x, rb, c[i].y);
        }
        ra += c[i].x;
        rb += c[i].y;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out_a[n] = sums_a[n_blocks];
        out_b[n] = sums_b[n_blocks];
    }
}

struct Scan2NoHook {
    __device__ __forceinline__ void operator()(unsigned long long, unsigned long long, unsigned long long, unsigned long long,
                                               unsigned long long) const {}
};

static inline size_t scan2_scratch_bytes(size_t n) { return 2 * scan_scratch_bytes(n); }

template <typename F, typename H = Scan2NoHook>
static void launch_exclusive_scan2(F count, size_t n, unsigned long long* sums, uint64_t* out_a, uint64_t* out_b, cudaStream_t s, H hook = H()) {
    const unsigned long long n_blocks = ceil_div(n, kScanTile);
    unsigned long long* sums_b = sums + n_blocks + 1;
    scan2_block_sums_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, sums, sums_b);
    scan2_sums_kernel<<<2, 1024, 0, s>>>(sums, sums_b, n_blocks);
    scan2_offsets_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(count, n, sums, sums_b, n_blocks, out_a, out_b, hook);
}

}  // namespace bn
