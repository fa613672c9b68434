// hamming.cu -- Hamming distance in the packed domain (sm_100a).
//
// Replaces /root/reference/src/utils/functions/hamming/scalar.rs:11-48 (hdist_scalar, here over
// n pairs) and /root/reference/src/utils/functions/hamming/multi.rs:12-67,122-160 (hdist, whole
// sequence).  XOR, fold each 2-bit group to one mismatch bit, popcount.
//
// HBM-bound: 20 bytes per pair (8 + 8 in, 4 out) for the pairs kernel, 0.5 bytes per base for
// the whole-sequence reduction.  The two 32-bit halves of a word are folded into one register
// (low half on even bits, high half on odd bits) so each 32-base word costs one POPC.
// Work is cut into chunks of tiles handed out by the hardware CTA scheduler (see codec.cu).
#include "common.cuh"
#include "launch.cuh"

namespace bn {

constexpr int kHamU = 4;
constexpr int kHamThreads = 512;
constexpr int kPairsT = 1;   // tiles per warp, pairs kernel
constexpr int kSumT = 4;     // tiles per warp, reduction kernel (fewer CTAs -> fewer atomics)

// mismatches among the bases selected by (mlo, mhi) = 0x55555555-pattern masks of the two halves
__device__ __forceinline__ uint32_t pair_distance(uint2 u, uint2 v, uint32_t mlo, uint32_t mhi) {
    const uint32_t xl = u.x ^ v.x, xh = u.y ^ v.y;
    const uint32_t a = (xl | (xl >> 1)) & mlo;
    const uint32_t b = (xh | (xh >> 1)) & mhi;
    return __popc(a | (b << 1));
}

// out[i] = hdist_scalar(u[i], v[i], len); two pairs per 128-bit load.
__global__ void __launch_bounds__(kHamThreads)
hdist_pairs_kernel(const uint4* __restrict__ u, const uint4* __restrict__ v, uint2* __restrict__ out,
                   unsigned long long n_vec, uint32_t mlo, uint32_t mhi) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * kHamU;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<kHamThreads, 1, kPairsT> walk(n_tiles);
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long i0 = t * kTile + lane;
        uint4 a[kHamU], b[kHamU];
#pragma unroll
        for (int j = 0; j < kHamU; ++j) {
            a[j] = ld128<LD_PLAIN>(u + i0 + 32 * j);
            b[j] = ld128<LD_PLAIN>(v + i0 + 32 * j);
        }
#pragma unroll
        for (int j = 0; j < kHamU; ++j) {
            uint2 d;
            d.x = pair_distance(make_uint2(a[j].x, a[j].y), make_uint2(b[j].x, b[j].y), mlo, mhi);
            d.y = pair_distance(make_uint2(a[j].z, a[j].w), make_uint2(b[j].z, b[j].w), mlo, mhi);
            st_stream_v2(out + i0 + 32 * j, d);
        }
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += kHamThreads) {
            const uint4 a = ld128<LD_PLAIN>(u + i), b = ld128<LD_PLAIN>(v + i);
            uint2 d;
            d.x = pair_distance(make_uint2(a.x, a.y), make_uint2(b.x, b.y), mlo, mhi);
            d.y = pair_distance(make_uint2(a.z, a.w), make_uint2(b.z, b.w), mlo, mhi);
            out[i] = d;
        }
    }
}

// one pair per thread (odd last pair, or misaligned pointers)
__global__ void __launch_bounds__(kThreads)
hdist_pairs_scalar_kernel(const uint64_t* __restrict__ u, const uint64_t* __restrict__ v, uint32_t* __restrict__ out,
                          unsigned long long first, unsigned long long n_pairs, uint32_t mlo, uint32_t mhi) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += step) {
        const uint64_t a = u[i], b = v[i];
        out[i] = pair_distance(make_uint2((uint32_t)a, (uint32_t)(a >> 32)), make_uint2((uint32_t)b, (uint32_t)(b >> 32)), mlo, mhi);
    }
}

// *total += mismatches over n_bases bases; n_vec = full 128-bit vectors (2 words each); the
// remaining (< 2) full words and the masked tail word are handled by one thread.
__global__ void __launch_bounds__(kHamThreads)
hdist_sum_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, unsigned long long n_vec,
                 unsigned long long n_bases, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long scratch[32];
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * kHamU;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<kHamThreads, 1, kSumT> walk(n_tiles);
    uint32_t s = 0;  // <= kSumT * kHamU * 64 per thread
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long i0 = t * kTile + lane;
        uint4 x[kHamU], y[kHamU];
#pragma unroll
        for (int j = 0; j < kHamU; ++j) {
            x[j] = ld128<LD_PLAIN>(a + i0 + 32 * j);
            y[j] = ld128<LD_PLAIN>(b + i0 + 32 * j);
        }
#pragma unroll
        for (int j = 0; j < kHamU; ++j) {
            s += pair_distance(make_uint2(x[j].x, x[j].y), make_uint2(y[j].x, y[j].y), 0x55555555u, 0x55555555u);
            s += pair_distance(make_uint2(x[j].z, x[j].w), make_uint2(y[j].z, y[j].w), 0x55555555u, 0x55555555u);
        }
    }
    unsigned long long acc = s;
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += kHamThreads) {
            const uint4 x = ld128<LD_PLAIN>(a + i), y = ld128<LD_PLAIN>(b + i);
            acc += pair_distance(make_uint2(x.x, x.y), make_uint2(y.x, y.y), 0x55555555u, 0x55555555u);
            acc += pair_distance(make_uint2(x.z, x.w), make_uint2(y.z, y.w), 0x55555555u, 0x55555555u);
        }
        if (threadIdx.x == 0) {
            const uint64_t* wa = reinterpret_cast<const uint64_t*>(a);
            const uint64_t* wb = reinterpret_cast<const uint64_t*>(b);
            const unsigned long long full = n_bases / 32;
            for (unsigned long long w = n_vec * 2; w < full; ++w) {
                const uint64_t d = wa[w] ^ wb[w];
                acc += __popcll((d | (d >> 1)) & 0x5555555555555555ull);
            }
            const unsigned rem = (unsigned)(n_bases % 32);
            if (rem) {
                const uint64_t d = (wa[full] ^ wb[full]) & ((1ull << (2 * rem)) - 1ull);
                acc += __popcll((d | (d >> 1)) & 0x5555555555555555ull);
            }
        }
    }
    const unsigned long long block_total = block_sum_u64(acc, scratch);
    if (threadIdx.x == 0 && block_total) atomicAdd(total, block_total);
}

// misaligned pointers: one word per thread
__global__ void __launch_bounds__(kThreads)
hdist_sum_scalar_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, unsigned long long n_bases,
                        unsigned long long* __restrict__ total) {
    __shared__ unsigned long long scratch[32];
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long full = n_bases / 32;
    const unsigned rem = (unsigned)(n_bases % 32);
    unsigned long long acc = 0;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < full + (rem ? 1 : 0); w += step) {
        uint64_t d = a[w] ^ b[w];
        if (w == full) d &= (1ull << (2 * rem)) - 1ull;
        acc += __popcll((d | (d >> 1)) & 0x5555555555555555ull);
    }
    const unsigned long long block_total = block_sum_u64(acc, scratch);
    if (threadIdx.x == 0 && block_total) atomicAdd(total, block_total);
}

cudaError_t launch_hdist(const DeviceInfo& di, const uint64_t* d_a, const uint64_t* d_b, size_t n_bases,
                         unsigned long long* d_total, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess || n_bases == 0) return e;
    if ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b)) & 15u) {
        static const int per_sm = blocks_per_sm(hdist_sum_scalar_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        hdist_sum_scalar_kernel<<<grid_for(ceil_div(ceil_div(n_bases, 32), kThreads), resident), kThreads, 0, s>>>(
            d_a, d_b, n_bases, d_total);
        return cudaGetLastError();
    }
    const unsigned long long n_vec = n_bases / 64;
    const unsigned long long ctas = TileWalk<kHamThreads, 1, kSumT>::ctas(n_vec / (32 * kHamU));
    hdist_sum_kernel<<<(unsigned)(ctas ? ctas : 1), kHamThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(d_a), reinterpret_cast<const uint4*>(d_b), n_vec, n_bases, d_total);
    return cudaGetLastError();
}

cudaError_t launch_hdist_pairs(const DeviceInfo& di, const uint64_t* d_u, const uint64_t* d_v, size_t n_pairs,
                               uint32_t len, uint32_t* d_out, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    const uint64_t mask = len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1ull);
    const uint32_t mlo = (uint32_t)mask & 0x55555555u, mhi = (uint32_t)(mask >> 32) & 0x55555555u;
    unsigned long long first_scalar = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_u) | reinterpret_cast<uintptr_t>(d_v)) & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_out) & 7u) == 0;
    if (aligned && n_pairs >= 2) {
        const unsigned long long n_vec = n_pairs / 2;
        const unsigned long long ctas = TileWalk<kHamThreads, 1, kPairsT>::ctas(n_vec / (32 * kHamU));
        hdist_pairs_kernel<<<(unsigned)(ctas ? ctas : 1), kHamThreads, 0, s>>>(
            reinterpret_cast<const uint4*>(d_u), reinterpret_cast<const uint4*>(d_v), reinterpret_cast<uint2*>(d_out), n_vec,
            mlo, mhi);
        first_scalar = n_vec * 2;
    }
    if (first_scalar < n_pairs) {
        static const int per_sm = blocks_per_sm(hdist_pairs_scalar_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        hdist_pairs_scalar_kernel<<<grid_for(ceil_div(n_pairs - first_scalar, kThreads), resident), kThreads, 0, s>>>(
            d_u, d_v, d_out, first_scalar, n_pairs, mlo, mhi);
    }
    return cudaGetLastError();
}

}  // namespace bn
