// counts.cu -- base_counts / gc_content on packed sequences (sm_100a).
//
// Replaces /root/reference/src/utils/analysis.rs:3-39, which decodes every base back to ASCII
// (PackedSequence::to_vec -> per-base get(), src/sequence.rs:116-135,198-212) and then counts
// bytes.  Here the counts come straight from the packed words: with lo = even bits and hi = odd
// bits of a word, C = lo & ~hi, G = hi & ~lo, T = lo & hi.  Only popc(lo), popc(hi), popc(lo & hi)
// are accumulated (C = nL - nT, G = nH - nT, T = nT) and A = n_bases - C - G - T, so the zero
// padding of a tail word can never be counted as 'A'.  Two 32-bit words share one POPC by putting
// the second word's bits on the free odd/even positions.
//
// gc_content keeps the reference's exact operation order from exact integer counts:
// (gc as f64 / len as f64) * 100.0, IEEE round-to-nearest, no FMA contraction.
//
// HBM-bound: 0.25 bytes per base for one long sequence; 8*ceil(len/32) + 32 + 8 bytes per read
// for the per-read batch.
#include "common.cuh"
#include "launch.cuh"

namespace bn {

struct Lht {
    uint32_t l, h, t;
};

// a, b: two packed 32-bit words (16 bases each)
__device__ __forceinline__ void count2(uint32_t a, uint32_t b, Lht& c) {
    const uint32_t L = (a & 0x55555555u) | ((b << 1) & 0xAAAAAAAAu);
    const uint32_t H = ((a >> 1) & 0x55555555u) | (b & 0xAAAAAAAAu);
    c.l += __popc(L);
    c.h += __popc(H);
    c.t += __popc(L & H);
}

__device__ __forceinline__ void count_word64(uint64_t w, unsigned long long& l, unsigned long long& h, unsigned long long& t) {
    Lht c = {0, 0, 0};
    count2((uint32_t)w, (uint32_t)(w >> 32), c);
    l += c.l;
    h += c.h;
    t += c.t;
}

__device__ __forceinline__ double gc_percent(unsigned long long gc, unsigned long long len) {
    if (len == 0) return 0.0;
    return __dmul_rn(__ddiv_rn(__ull2double_rn(gc), __ull2double_rn(len)), 100.0);
}

constexpr int kCntU = 4;
constexpr int kCntThreads = 512;
constexpr int kCntT = 4;  // tiles per warp: fewer CTAs -> fewer atomics on the three accumulators

// acc[1] += popc(lo), acc[2] += popc(hi), acc[3] += popc(lo & hi) over the first n_bases bases.
__global__ void __launch_bounds__(kCntThreads)
base_counts_kernel(const uint4* __restrict__ in, unsigned long long n_vec, unsigned long long n_bases,
                   unsigned long long* __restrict__ acc) {
    __shared__ unsigned long long scratch[32];
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * kCntU;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<kCntThreads, 1, kCntT> walk(n_tiles);
    Lht c = {0, 0, 0};  // <= kCntT * kCntU * 64 per thread
    for (unsigned long long tile = walk.first; tile < walk.end; tile += walk.step) {
        const unsigned long long i0 = tile * kTile + lane;
        uint4 v[kCntU];
#pragma unroll
        for (int j = 0; j < kCntU; ++j) v[j] = ld128<LD_PLAIN>(in + i0 + 32 * j);
#pragma unroll
        for (int j = 0; j < kCntU; ++j) {
            count2(v[j].x, v[j].y, c);
            count2(v[j].z, v[j].w, c);
        }
    }
    unsigned long long l = c.l, h = c.h, t = c.t;
    if (blockIdx.x == gridDim.x - 1) {  // vectors after the last full tile, then single words, tail word masked
        Lht d = {0, 0, 0};
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += kCntThreads) {
            const uint4 v = ld128<LD_PLAIN>(in + i);
            count2(v.x, v.y, d);
            count2(v.z, v.w, d);
        }
        l += d.l;
        h += d.h;
        t += d.t;
        if (threadIdx.x == 0) {
            const uint64_t* w = reinterpret_cast<const uint64_t*>(in);
            const unsigned long long full = n_bases / 32;
            for (unsigned long long i = n_vec * 2; i < full; ++i) count_word64(w[i], l, h, t);
            const unsigned rem = (unsigned)(n_bases % 32);
            if (rem) count_word64(w[full] & ((1ull << (2 * rem)) - 1ull), l, h, t);
        }
    }
    l = block_sum_u64(l, scratch);
    h = block_sum_u64(h, scratch);
    t = block_sum_u64(t, scratch);
    if (threadIdx.x == 0) {
        if (l) atomicAdd(acc + 1, l);
        if (h) atomicAdd(acc + 2, h);
        if (t) atomicAdd(acc + 3, t);
    }
}

// misaligned pointer: one word per thread
__global__ void __launch_bounds__(kThreads)
base_counts_scalar_kernel(const uint64_t* __restrict__ w, unsigned long long n_bases, unsigned long long* __restrict__ acc) {
    __shared__ unsigned long long scratch[32];
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long full = n_bases / 32;
    const unsigned rem = (unsigned)(n_bases % 32);
    unsigned long long l = 0, h = 0, t = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < full + (rem ? 1 : 0); i += step)
        count_word64(i == full ? w[i] & ((1ull << (2 * rem)) - 1ull) : w[i], l, h, t);
    l = block_sum_u64(l, scratch);
    h = block_sum_u64(h, scratch);
    t = block_sum_u64(t, scratch);
    if (threadIdx.x == 0) {
        if (l) atomicAdd(acc + 1, l);
        if (h) atomicAdd(acc + 2, h);
        if (t) atomicAdd(acc + 3, t);
    }
}

// (_, nL, nH, nT) -> [A, C, G, T] in place, plus gc
__global__ void base_counts_finalize_kernel(unsigned long long* counts, unsigned long long n_bases, double* gc) {
    const unsigned long long l = counts[1], h = counts[2], t = counts[3];
    const unsigned long long c = l - t, g = h - t;
    counts[0] = n_bases - c - g - t;
    counts[1] = c;
    counts[2] = g;
    counts[3] = t;
    if (gc) *gc = gc_percent(c + g, n_bases);
}

constexpr int kBatchRounds = 4;  // reads per thread (or per warp): fewer CTAs -> fewer atomics on the totals

// Per-read batch.  LANES = lanes cooperating on one read (1 for short reads, 32 for long ones).
template <int LANES>
__global__ void __launch_bounds__(kThreads)
base_counts_batch_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets,
                         const uint64_t* __restrict__ lens, unsigned long long n_reads, unsigned long long fixed_len,
                         unsigned long long* __restrict__ counts4, double* __restrict__ gc,
                         unsigned long long* __restrict__ totals) {
    // lens == nullptr: fixed-length reads, read r = words[r * ceil(fixed_len/32) ..) (no index arrays to fetch)
    __shared__ unsigned long long scratch[32];
    // a CTA owns kBatchRounds * (kThreads / LANES) consecutive reads (hardware CTA scheduling, see codec.cu)
    const unsigned sub = threadIdx.x % LANES;
    constexpr unsigned kGroups = kThreads / LANES;
    const unsigned long long first = (unsigned long long)blockIdx.x * (kGroups * kBatchRounds) + threadIdx.x / LANES;
    unsigned long long ta = 0, tc = 0, tg = 0, tt = 0;
    // (fetching the index entries of all four rounds up front -- one trip to memory instead of four dependent ones -- was
    // measured: the unrolled body takes 60 registers instead of 40 / 48, and 0.159 -> 0.167 ms offset-indexed, 0.130 -> 0.148 ms
    // fixed-length for 10 M x 150 bp: residency wins here)
    for (int it = 0; it < kBatchRounds; ++it) {  // uniform trip count: shuffles stay converged
        const unsigned long long r = first + (unsigned long long)it * kGroups;
        unsigned long long l = 0, h = 0, t = 0, len = 0;
        if (r < n_reads) {
            len = lens ? lens[r] : fixed_len;
            const uint64_t* w = words + (lens ? word_offsets[r] : r * ((fixed_len + 31) / 32));
            const unsigned long long full = len / 32;
            const unsigned rem = (unsigned)(len % 32);
            for (unsigned long long i = sub; i < full; i += LANES) count_word64(__ldg(w + i), l, h, t);
            if (rem && sub == full % LANES) count_word64(__ldg(w + full) & ((1ull << (2 * rem)) - 1ull), l, h, t);
        }
        if (LANES > 1) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) {
                l += __shfl_xor_sync(0xffffffffu, l, o);
                h += __shfl_xor_sync(0xffffffffu, h, o);
                t += __shfl_xor_sync(0xffffffffu, t, o);
            }
        }
        if (r < n_reads && sub == 0) {
            const unsigned long long c = l - t, g = h - t, a = len - c - g - t;
            if (counts4) {
                ulonglong2* o = reinterpret_cast<ulonglong2*>(counts4 + 4 * r);
                o[0] = make_ulonglong2(a, c);
                o[1] = make_ulonglong2(g, t);
            }
            if (gc) gc[r] = gc_percent(c + g, len);
            ta += a;
            tc += c;
            tg += g;
            tt += t;
        }
    }
    if (totals) {
        ta = block_sum_u64(ta, scratch);
        tc = block_sum_u64(tc, scratch);
        tg = block_sum_u64(tg, scratch);
        tt = block_sum_u64(tt, scratch);
        if (threadIdx.x == 0) {
            if (ta) atomicAdd(totals + 0, ta);
            if (tc) atomicAdd(totals + 1, tc);
            if (tg) atomicAdd(totals + 2, tg);
            if (tt) atomicAdd(totals + 3, tt);
        }
    }
}

cudaError_t launch_base_counts(const DeviceInfo& di, const uint64_t* d_words, size_t n_bases,
                               unsigned long long* d_counts, double* d_gc, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_counts, 0, 4 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_bases) {
        if (reinterpret_cast<uintptr_t>(d_words) & 15u) {
            static const int per_sm = blocks_per_sm(base_counts_scalar_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
            base_counts_scalar_kernel<<<grid_for(ceil_div(ceil_div(n_bases, 32), kThreads), resident), kThreads, 0, s>>>(
                d_words, n_bases, d_counts);
        } else {
            const unsigned long long n_vec = n_bases / 64;
            const unsigned long long ctas = TileWalk<kCntThreads, 1, kCntT>::ctas(n_vec / (32 * kCntU));
            base_counts_kernel<<<(unsigned)(ctas ? ctas : 1), kCntThreads, 0, s>>>(reinterpret_cast<const uint4*>(d_words), n_vec,
                                                                                   n_bases, d_counts);
        }
    }
    base_counts_finalize_kernel<<<1, 1, 0, s>>>(d_counts, n_bases, d_gc);
    return cudaGetLastError();
}

cudaError_t launch_base_counts_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets,
                                     const uint64_t* d_lens, size_t n_reads, size_t fixed_len, size_t n_words_hint,
                                     unsigned long long* d_counts4, double* d_gc, unsigned long long* d_totals,
                                     cudaStream_t s) {
    if (d_totals) {
        cudaError_t e = cudaMemsetAsync(d_totals, 0, 4 * sizeof(unsigned long long), s);
        if (e != cudaSuccess) return e;
    }
    if (n_reads == 0) return cudaSuccess;
    if (n_words_hint / n_reads >= 64) {  // long reads: a warp per read
        base_counts_batch_kernel<32><<<(unsigned)ceil_div(n_reads, kWarpsPerBlock * kBatchRounds), kThreads, 0, s>>>(
            d_words, d_word_offsets, d_lens, n_reads, fixed_len, d_counts4, d_gc, d_totals);
    } else {
        base_counts_batch_kernel<1><<<(unsigned)ceil_div(n_reads, kThreads * kBatchRounds), kThreads, 0, s>>>(
            d_words, d_word_offsets, d_lens, n_reads, fixed_len, d_counts4, d_gc, d_totals);
    }
    return cudaGetLastError();
}

}  // namespace bn
