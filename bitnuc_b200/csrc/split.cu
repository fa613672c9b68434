// split.cu -- split_packed over a batch of packed reads (sm_100a).  SURVEY.md 8(f) rank 1.
//
// Replaces the caller-side loop over /root/reference/src/utils/functions/split.rs:14-102: read r
// (lens[r] bases, `ebuf` = words[word_offsets[r] .. word_offsets[r+1]), normally ceil(lens[r]/32) words) is
// split at base idx[r] into a left buffer of idx/32 + 1 words (an all-zero extra word when idx % 32 == 0,
// split.rs:51,72-77) and a right buffer of ebuf.len() - idx/32 words (the loop at split.rs:83-94 pushes one
// word per remaining input word); idx == 0 / idx == len copy ebuf through (split.rs:33-42); an empty ebuf
// gives two empty outputs (split.rs:45-47).  The reference's sequential carry loop (split.rs:83-94) makes right word j
//     (ebuf[c+j] >> shift) | (ebuf[c+j-1] << (64 - shift))      c = idx/32, shift = 2*(idx%32), j >= 1
// i.e. it carries the PREVIOUS word's low bits upward; that is reproduced bit for bit (for reads longer
// than 32 bases split off a word boundary it is not the suffix of the read).  The final
// `carry != 0 && rbuf.len() < right_chunks` push (split.rs:97-99) can never fire when ebuf holds at least
// ceil(len/32) words, because ceil(len/32) - idx/32 >= ceil((len-idx)/32); a shorter ebuf is rejected here
// (the reference either panics at split.rs:77 or returns a truncated right half).
//
// Every output word depends on at most two input words: one thread per read here (reads on this path
// are short: barcode | insert splits, benches/functions_benchmark.rs:59 uses 30-280 bases).
// Status word = min over failing reads of (read index << 1 | kind): kind 0 = idx > len ->
// IndexOutOfBounds{idx, len} (split.rs:22-27), kind 1 = 0 < idx < len and a non-empty ebuf shorter than ceil(len/32) words.
#include "common.cuh"
#include "launch.cuh"
#include "lookback.cuh"

namespace bn {

// split_one: the words of one read, written to lo / ro (global memory or the CTA's staged image of its output spans)
__device__ __forceinline__ void split_one(const uint64_t* __restrict__ w, unsigned long long nw, unsigned long long slen,
                                          unsigned long long i, uint64_t* __restrict__ lo, uint64_t* __restrict__ ro) {
    if (i == 0) {
        for (unsigned long long j = 0; j < nw; ++j) ro[j] = __ldg(w + j);
    } else if (i == slen) {
        for (unsigned long long j = 0; j < nw; ++j) lo[j] = __ldg(w + j);
    } else if (nw) {
        const unsigned long long c = i / 32;
        const unsigned sh = 2 * (unsigned)(i % 32);
        for (unsigned long long j = 0; j < c; ++j) lo[j] = __ldg(w + j);
        uint64_t prev = __ldg(w + c);
        lo[c] = sh ? prev & ((1ull << sh) - 1ull) : 0ull;
        ro[0] = prev >> sh;
        for (unsigned long long j = 1; c + j < nw; ++j) {
            const uint64_t cur = __ldg(w + c + j);
            ro[j] = sh ? (cur >> sh) | (prev << (64 - sh)) : cur;
            prev = cur;
        }
    }
}

constexpr int kSpCap = 2048;   // words of each output span a CTA can stage (256 reads x 8 words: reads up to 256 bases on average)

// ONE pass: a CTA takes a tile of 256 consecutive reads (tiles numbered by a ticket, so look-back never waits for a CTA
// that is not running), reads their (word_offsets, lens, idx) once, scans the two word counts inside the CTA, gets the
// totals of all earlier tiles by decoupled look-back (lookback.cuh), writes the offsets, and splits one read per thread
// into a shared-memory image of the tile's two output spans -- which are contiguous in `left` / `right` because the
// reads are consecutive -- stored with coalesced 8-byte-per-lane stores.  (The three-launch scan + thread-per-read
// kernel it replaces read the shapes three times and wrote every output word as a lone 8-byte store: 0.50 of the HBM
// roofline on 40 M short reads.)  A tile whose spans do not fit the image writes straight to global memory.
__global__ void __launch_bounds__(kThreads)
split_packed_fused_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets,
                          const uint64_t* __restrict__ lens, const uint64_t* __restrict__ idx, unsigned long long n_reads,
                          uint64_t* __restrict__ left, uint64_t* __restrict__ left_offsets, uint64_t* __restrict__ right,
                          uint64_t* __restrict__ right_offsets, unsigned long long* __restrict__ status,
                          unsigned long long* __restrict__ lb, unsigned long long n_tiles) {
    __shared__ uint64_t s_left[kSpCap], s_right[kSpCap];
    __shared__ unsigned long long s_tile, s_warp[2][kWarpsPerBlock], s_base[2], s_tot[2];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(lb, 1ull);
    __syncthreads();
    const unsigned long long tile = s_tile;
    const unsigned long long r = tile * kThreads + tid;
    unsigned long long wo = 0, nw = 0, slen = 0, i = 0, nl = 0, nr = 0;
    bool live = false;
    if (r < n_reads) {
        wo = word_offsets[r];
        nw = word_offsets[r + 1] - wo;
        slen = lens[r];
        i = idx[r];
        const bool oob = i > slen;
        if (oob || (i && i < slen && nw && nw < (slen + 31) / 32)) {   // an error: takes no room
            const unsigned long long key = r << 1 | (oob ? 0ull : 1ull);
            if (key < ld_volatile_u64(status)) atomicMin(status, key);
        } else {
            live = true;
            if (i == 0) nr = nw;
            else if (i == slen) nl = nw;
            else if (nw) nl = i / 32 + 1, nr = nw - i / 32;
        }
    }
    // exclusive scan of (nl, nr) over the CTA
    unsigned long long il = nl, ir = nr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, ir, o);
        if (lane >= (unsigned)o) il += a, ir += b;
    }
    if (lane == 31) s_warp[0][warp] = il, s_warp[1][warp] = ir;
    __syncthreads();
    if (warp == 0) {
        unsigned long long a = lane < kWarpsPerBlock ? s_warp[0][lane] : 0ull, b = lane < kWarpsPerBlock ? s_warp[1][lane] : 0ull;
        unsigned long long ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < kWarpsPerBlock; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= (unsigned)o) ia += x, ib += y;
        }
        if (lane < kWarpsPerBlock) s_warp[0][lane] = ia - a, s_warp[1][lane] = ib - b;
        const unsigned long long agg[2] = {__shfl_sync(0xffffffffu, ia, kWarpsPerBlock - 1), __shfl_sync(0xffffffffu, ib, kWarpsPerBlock - 1)};
        unsigned long long excl[2];
        lookback_exclusive<2>(lb + 1, n_tiles, tile, agg, excl);
        if (lane == 0) s_base[0] = excl[0], s_base[1] = excl[1], s_tot[0] = agg[0], s_tot[1] = agg[1];
    }
    __syncthreads();
    const unsigned long long ll = s_warp[0][warp] + il - nl, lr = s_warp[1][warp] + ir - nr;   // offsets inside the tile's spans
    const unsigned long long base_l = s_base[0], base_r = s_base[1], tot_l = s_tot[0], tot_r = s_tot[1];
    if (r < n_reads) {
        left_offsets[r] = base_l + ll;
        right_offsets[r] = base_r + lr;
        if (r + 1 == n_reads) left_offsets[n_reads] = base_l + ll + nl, right_offsets[n_reads] = base_r + lr + nr;
    }
    const bool staged = tot_l <= kSpCap && tot_r <= kSpCap;
    if (live) split_one(words + wo, nw, slen, i, staged ? s_left + ll : left + base_l + ll, staged ? s_right + lr : right + base_r + lr);
    if (staged) {
        __syncthreads();
        for (unsigned long long k = tid; k < tot_l; k += kThreads) left[base_l + k] = s_left[k];
        for (unsigned long long k = tid; k < tot_r; k += kThreads) right[base_r + k] = s_right[k];
    }
}

// ticket + two channels of tile descriptors
size_t split_packed_scratch_bytes(size_t n_reads) { return lookback_bytes(ceil_div(n_reads ? n_reads : 1, kThreads), 2); }

cudaError_t launch_split_packed_batch(const DeviceInfo&, const uint64_t* d_words, const uint64_t* d_word_offsets,
                                      const uint64_t* d_lens, const uint64_t* d_idx, size_t n_reads, uint64_t* d_left,
                                      uint64_t* d_left_offsets, uint64_t* d_right, uint64_t* d_right_offsets,
                                      unsigned long long* d_status, void* d_scratch, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_reads == 0) {
        e = cudaMemsetAsync(d_left_offsets, 0, sizeof(uint64_t), s);
        return e != cudaSuccess ? e : cudaMemsetAsync(d_right_offsets, 0, sizeof(uint64_t), s);
    }
    const unsigned long long n_tiles = ceil_div(n_reads, kThreads);
    e = cudaMemsetAsync(d_scratch, 0, lookback_bytes(n_tiles, 2), s);
    if (e != cudaSuccess) return e;
    split_packed_fused_kernel<<<(unsigned)n_tiles, kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, d_idx, n_reads, d_left, d_left_offsets,
                                                                    d_right, d_right_offsets, d_status,
                                                                    static_cast<unsigned long long*>(d_scratch), n_tiles);
    return cudaGetLastError();
}

}  // namespace bn
