// batch.cu -- variable-length read batches: offset-indexed ASCII reads -> per-read packed words (sm_100a).
//
// Replaces the caller-side loop `for read in reads { PackedSequence::new(read)? }`
// (/root/reference/src/sequence.rs:40-52 -> src/utils/packing/avx.rs:130-151): every read starts
// on a fresh 64-bit word, an empty read takes no words (sequence.rs:42-46).
//
// Two steps on the device:
//   1. word offsets = exclusive prefix sum of ceil(len/32) over the reads (block sums, one-CTA scan
//      of the sums, block-local scan + offset);
//   2. encode, one thread per OUTPUT word: the word's read is found by a search of the word-offset
//      array that is narrowed per CTA tile first (tiles are pulled from a global atomic counter), its <= 32 source bytes are fetched as nine aligned
//      32-bit loads and funnel-shifted into place, bytes past the end of the read are replaced by 'A'.
// Work per thread is uniform whatever the length mix (50 bp .. 10 kbp).  HBM-bound at
// 1 B/base in + 8 B per word out + 16 B per read of offsets.
#include "common.cuh"
#include "launch.cuh"

namespace bn {

constexpr int kScanItems = 4;                          // reads per thread in the scan kernels
constexpr int kScanTile = kThreads * kScanItems;       // reads per CTA

__device__ __forceinline__ unsigned long long words_of_read(const uint64_t* __restrict__ offsets, unsigned long long r) {
    return (offsets[r + 1] - offsets[r] + 31) / 32;
}

__global__ void __launch_bounds__(kThreads)
batch_block_sums_kernel(const uint64_t* __restrict__ offsets, unsigned long long n_reads, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long scratch[32];
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (r0 + i < n_reads) s += words_of_read(offsets, r0 + i);
    s = block_sum_u64(s, scratch);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

// exclusive scan of sums[0..n) in place by one CTA; sums[n] = total
__global__ void __launch_bounds__(1024) batch_scan_sums_kernel(unsigned long long* __restrict__ sums, unsigned long long n) {
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

__global__ void __launch_bounds__(kThreads)
batch_word_offsets_kernel(const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                          const unsigned long long* __restrict__ sums, unsigned long long n_blocks,
                          uint64_t* __restrict__ word_offsets) {
    __shared__ unsigned long long warp_tot[32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long c[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        c[i] = r0 + i < n_reads ? words_of_read(offsets, r0 + i) : 0;
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
        if (r0 + i < n_reads) word_offsets[r0 + i] = run;
        run += c[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) word_offsets[n_reads] = sums[n_blocks];
}

// index of the read owning output word w: the last r in [lo, hi] with word_offsets[r] <= w
__device__ __forceinline__ unsigned long long owner_read(const uint64_t* __restrict__ wo, unsigned long long lo,
                                                         unsigned long long hi, unsigned long long w) {
    while (lo < hi) {
        const unsigned long long mid = lo + (hi - lo + 1) / 2;
        if (__ldg(wo + mid) <= w) lo = mid; else hi = mid - 1;
    }
    return lo;
}

constexpr int kBatchItems = 8;                          // output words per thread
constexpr int kBatchTile = kThreads * kBatchItems;      // output words per CTA tile

__global__ void __launch_bounds__(kThreads)
encode_batch_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                    const uint64_t* __restrict__ word_offsets, uint64_t* __restrict__ out,
                    uint32_t* __restrict__ read_status, unsigned long long* __restrict__ status,
                    unsigned long long* __restrict__ tile_counter) {
    __shared__ unsigned long long range[2];
    __shared__ unsigned long long tile_s;
    const unsigned long long total_words = word_offsets[n_reads];
    const unsigned long long buf_lo = offsets[0], buf_hi = offsets[n_reads];  // valid byte range of `bytes`
    const unsigned long long n_tiles = ceil_div(total_words, kBatchTile);
    for (;;) {  // persistent CTAs pull tiles from a global counter: dynamic balance, word count unknown to the host
        __syncthreads();
        if (threadIdx.x == 0) tile_s = atomicAdd(tile_counter, 1ull);
        __syncthreads();
        const unsigned long long tile = tile_s;
        if (tile >= n_tiles) break;
        const unsigned long long w0 = tile * kBatchTile;
        const unsigned long long w1 = w0 + kBatchTile < total_words ? w0 + kBatchTile : total_words;
        if (threadIdx.x < 2) range[threadIdx.x] = owner_read(word_offsets, 0, n_reads - 1, threadIdx.x ? w1 - 1 : w0);
        __syncthreads();
        const unsigned long long r_lo = range[0], r_hi = range[1];
#pragma unroll 2
        for (int it = 0; it < kBatchItems; ++it) {
            const unsigned long long w = w0 + (unsigned long long)it * kThreads + threadIdx.x;
            if (w >= w1) break;
            const unsigned long long r = owner_read(word_offsets, r_lo, r_hi, w);
            const unsigned long long rb = __ldg(offsets + r), re = __ldg(offsets + r + 1);
            const unsigned long long src = rb + (w - __ldg(word_offsets + r)) * 32ull;  // byte offset in `bytes`
            const int nb = (int)(re - src < 32 ? re - src : 32);                        // bases in this word
            const uintptr_t addr = reinterpret_cast<uintptr_t>(bytes) + src;
            const unsigned sh = 8 * (unsigned)(addr & 3u);
            uint32_t x[9];
            const unsigned long long a0 = src - (addr & 3u);  // may wrap below buf_lo: checked next
            if ((addr & 3u) <= src - buf_lo && a0 + 36 <= buf_hi) {
                const uint32_t* p = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
#pragma unroll
                for (int i = 0; i < 9; ++i) x[i] = __ldg(p + i);
            } else {  // the aligned window pokes outside the buffer: byte loads with bounds
#pragma unroll
                for (int i = 0; i < 9; ++i) x[i] = 0;
                for (int j = 0; j < nb; ++j) {
                    const unsigned q = (unsigned)(addr & 3u) + j;
                    x[q >> 2] |= (uint32_t)bytes[src + j] << (8 * (q & 3));
                }
            }
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t f = __funnelshift_r(x[i], x[i + 1], sh);
                const int keep = nb - 4 * i;
                const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u);
                v[i] = (f & m) | (0x41414141u & ~m);
            }
            uint32_t bad = 0;
            const uint32_t lo = pack16(make_uint4(v[0], v[1], v[2], v[3]), bad);
            const uint32_t hi = pack16(make_uint4(v[4], v[5], v[6], v[7]), bad);
            out[w] = ((uint64_t)hi << 32) | lo;
            if (bad & kValidMask) {
                for (int j = 0; j < nb; ++j) {
                    const uint32_t b = bytes[src + j];
                    if (!byte_is_valid(b)) {
                        report_invalid(status, src + j, b);
                        if (read_status) atomicMin(read_status + r, (uint32_t)(src + j - rb));
                        break;
                    }
                }
            }
        }
    }
}

size_t encode_batch_scratch_bytes(size_t n_reads) {
    return (ceil_div(n_reads ? n_reads : 1, kScanTile) + 2) * sizeof(unsigned long long);  // block sums, total, tile counter
}

cudaError_t launch_encode_batch(const DeviceInfo& di, const uint8_t* d_bytes, const uint64_t* d_offsets,
                                size_t n_reads, uint64_t* d_out_words, uint64_t* d_out_word_offsets,
                                uint32_t* d_read_status, unsigned long long* d_status, void* d_scratch,
                                cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_reads == 0) return cudaMemsetAsync(d_out_word_offsets, 0, sizeof(uint64_t), s);
    if (d_read_status) {
        e = cudaMemsetAsync(d_read_status, 0xFF, n_reads * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
    }
    unsigned long long* sums = static_cast<unsigned long long*>(d_scratch);
    const unsigned long long n_blocks = ceil_div(n_reads, kScanTile);
    unsigned long long* tile_counter = sums + n_blocks + 1;
    e = cudaMemsetAsync(tile_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    batch_block_sums_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(d_offsets, n_reads, sums);
    batch_scan_sums_kernel<<<1, 1024, 0, s>>>(sums, n_blocks);
    batch_word_offsets_kernel<<<(unsigned)n_blocks, kThreads, 0, s>>>(d_offsets, n_reads, sums, n_blocks, d_out_word_offsets);
    static const int resident = resident_blocks(encode_batch_kernel, kThreads, di);
    // the number of output words is only known on the device: launch a full persistent grid
    encode_batch_kernel<<<resident, kThreads, 0, s>>>(d_bytes, d_offsets, n_reads, d_out_word_offsets, d_out_words,
                                                      d_read_status, d_status, tile_counter);
    return cudaGetLastError();
}

}  // namespace bn
