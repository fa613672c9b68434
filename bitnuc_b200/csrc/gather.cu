// gather.cu -- PackedSequence::get / PackedSequence::slice over a batch of queries (sm_100a).  SURVEY.md 8(f) rank 2.
//
// Replaces the caller-side loops over /root/reference/src/sequence.rs:116-135 (get: one 2-bit poke per call) and
// sequence.rs:198-212 (slice: a Vec filled by per-base get()): query q names a read of a packed batch (words at
// word_offsets[read], lens[read] bases) and a base index / a base range; the bases come out as upper-case ASCII
// without decoding the whole read.
//   get:   index >= len            -> IndexOutOfBounds{index, len}      (sequence.rs:117-122)
//   slice: start > end || end > len -> InvalidRange{start, end, len}     (sequence.rs:199-205)
// The first failing query in index order is reported (status word = min failing query index); failing queries
// produce no output (slice) or 0 (get).
//
// slice: an exclusive prefix sum of (end - start) places every query's bytes -- computed inside slice_short_kernel itself
// (2048-query tiles, decoupled look-back: lookback.cuh), which also validates the queries and cuts the ranges of up to 64
// bases, one query per lane, staged per warp for coalesced stores; longer ranges are queued and cut a warp per range
// (slice_long_kernel).
#include "common.cuh"
#include "launch.cuh"
#include "lookback.cuh"

namespace bn {

// ASCII of the 4 bases starting at base b of the read whose words (viewed as 32-bit halves) start at w32:
// a funnel shift of two halves, the low byte spread to one code per nibble, PRMT as a 4-entry LUT.
__device__ __forceinline__ uint32_t ascii4_at(const uint32_t* __restrict__ w32, unsigned long long b) {
    const unsigned long long i = b >> 4;
    const unsigned sh = 2u * (unsigned)(b & 15u);
    const uint32_t lo = __ldg(w32 + i);
    const uint32_t hi = sh > 24 ? __ldg(w32 + i + 1) : 0u;  // only reached when all 4 bases exist, so that half does too
    uint32_t t = __funnelshift_r(lo, hi, sh) & 0xFFu;
    t = (t | (t << 4)) & 0x0F0Fu;
    t = (t | (t << 2)) & 0x3333u;
    return __byte_perm(0x54474341u, 0u, t);
}
__device__ __forceinline__ uint8_t ascii1_at(const uint32_t* __restrict__ w32, unsigned long long b) {
    return (uint8_t)(0x54474341u >> (8u * ((__ldg(w32 + (b >> 4)) >> (2u * (unsigned)(b & 15u))) & 3u)));
}

constexpr unsigned kSliceShort = 64;                      // ranges up to this many bases are cut by one thread each
constexpr unsigned kSliceStage = 32 * kSliceShort + 16;   // bytes a warp stages for its 32 queries (+ alignment slack)

// ONE pass over the queries.  A CTA takes a tile of 2048 consecutive queries (tiles numbered by a ticket; warp w owns 8 rows of 32
// of them, one look-back per tile -- per 256 queries the chain of spinning tiles was slower than the scan), validates each
// (the only place lens[read] -- a random sector per query -- is fetched; a failing query is reported as
// status = min failing query index and takes no room), scans the range lengths inside the CTA, gets the bytes of all
// earlier tiles by decoupled look-back (lookback.cuh) and writes out_offsets -- the three-launch cached scan this
// replaces wrote every length to HBM and read it, q_read, q_start and the offsets back.
// Short ranges (<= 64 bases) are then cut one per thread, so a warp keeps 32 independent gathers in flight (this path is
// a latency-bound chain query -> read offsets -> words).  The 2n bits of a range are pulled into a 128-bit register
// window (<= 3 words) and leave 4 bases at a time (spread + PRMT as a 4-entry LUT).  The 32 ranges of a warp are
// adjacent in the output (prefix sums of consecutive queries), so the threads write into a shared-memory image of
// that span -- laid out at the same 16-byte phase as the global span -- and the warp then stores it with coalesced
// 128-bit stores.  Longer ranges are queued for slice_long_kernel.
#ifndef BN_SL_ROWS
#define BN_SL_ROWS 8
#endif
constexpr int kSlRows = BN_SL_ROWS;                        // rows of 32 queries per warp: a CTA tile holds 8 warps x 8 rows x 32 = 2048 queries
constexpr int kSlTile = kThreads * kSlRows;
#ifndef BN_SL_BATCH
#define BN_SL_BATCH 2   // 2: 0.307 ms, 48 registers; 4: 0.336 ms, 62 registers; 8: 0.50 ms, 118 registers (10 M queries)
#endif
constexpr int kSlBatch = BN_SL_BATCH;             // rows of pass A whose loads are in flight together

// One short range (n <= kSliceShort bases from base s of the words at w) -> n ASCII bytes at o; `phase` = the low two bits
// of the output address (o may be the staged image: same phase).
__device__ __forceinline__ void slice_cut(const uint64_t* __restrict__ w, unsigned long long s, unsigned n, unsigned phase, uint8_t* __restrict__ o) {
    // the range's <= 64 bases as one 128-bit window, from ONE set of loads (the head bytes used to fetch their words
    // first, then the body fetched them again: two dependent trips, even if the second hit L1)
    const unsigned long long wi = s >> 5;
    const unsigned sh = 2u * (unsigned)(s & 31u);
    const unsigned need_bits = sh + 2u * n;                      // bits needed counted from the start of word wi
    const uint64_t w0 = __ldg(w + wi);
    const uint64_t w1 = need_bits > 64 ? __ldg(w + wi + 1) : 0ull;
    const uint64_t w2 = need_bits > 128 ? __ldg(w + wi + 2) : 0ull;
    uint64_t lo = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    uint64_t hi = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
    // head: bases up to the first 4-byte aligned output address
    const unsigned head = min((4u - phase) & 3u, n);
    if (head) {
        for (unsigned i = 0; i < head; ++i) o[i] = (uint8_t)(0x54474341u >> (8u * ((unsigned)(lo >> (2 * i)) & 3u)));
        lo = (lo >> (2 * head)) | (hi << (64 - 2 * head));
        hi >>= 2 * head;
        n -= head;
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o + head);
    const unsigned body = n / 4;
#pragma unroll
    for (unsigned g = 0; g < kSliceShort / 4; ++g) {
        if (g < body) {
            uint32_t t = (uint32_t)((g < 8 ? lo >> (8 * g) : hi >> (8 * (g - 8))) & 0xFFu);
            t = (t | (t << 4)) & 0x0F0Fu;
            t = (t | (t << 2)) & 0x3333u;
            o32[g] = __byte_perm(0x54474341u, 0u, t);
        }
    }
    const unsigned done = 4 * body;
    if (done < n) {
        const uint32_t t = (uint32_t)((body < 8 ? lo >> (8 * body) : hi >> (8 * (body - 8))) & 0xFFu);
        for (unsigned i = 0; done + i < n; ++i) o[head + done + i] = (uint8_t)(0x54474341u >> (8u * ((t >> (2 * i)) & 3u)));
    }
}

__global__ void __launch_bounds__(kThreads)
slice_short_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets, const uint64_t* __restrict__ lens,
                   unsigned long long n_reads, const uint64_t* __restrict__ q_read, const uint64_t* __restrict__ q_start,
                   const uint64_t* __restrict__ q_end, unsigned long long nq, uint8_t* __restrict__ out,
                   uint64_t* __restrict__ out_offsets, unsigned long long* __restrict__ status, unsigned long long* __restrict__ long_count,
                   unsigned long long* __restrict__ long_list, unsigned long long* __restrict__ lb, unsigned long long n_tiles) {
    __shared__ __align__(16) uint8_t stage[kWarpsPerBlock][kSliceStage];
    __shared__ uint32_t s_cnt[kSlRows][kThreads];               // range lengths from pass A (written and read by the same thread);
                                                                // 0xFFFFFFFF: longer than that, derived again from the query in pass B
    __shared__ unsigned long long s_src[kSlRows][kThreads];     // (index of the word holding the range's first base) << 5 | its place in the word
    __shared__ unsigned long long s_tile, s_warp[kWarpsPerBlock], s_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(lb, 1ull);
    __syncthreads();
    const unsigned long long tile = s_tile;
    const unsigned long long row0 = tile * kSlTile + warp * (32 * kSlRows) + lane;   // this lane's query in row 0
    // ---- pass A: validate every query of the warp's rows, add up the range lengths.  The rows are taken kSlBatch at a time
    // and every level of the dependent chain (query -> lens / word_offsets of its read) is issued for the whole batch before
    // anything is tested: written row by row, each row's validity branches sat between its loads and the next row's, and the
    // eight rows' three round trips ran one after the other (24 per tile; cuobjdump showed the LDGs of row i + 1 behind the
    // branches of row i).
    unsigned long long wsum = 0;
#pragma unroll
    for (int i0 = 0; i0 < kSlRows; i0 += kSlBatch) {
        unsigned long long rd[kSlBatch], qs[kSlBatch], qe[kSlBatch], len[kSlBatch], wo[kSlBatch];
        bool in[kSlBatch];
#pragma unroll
        for (int j = 0; j < kSlBatch; ++j) {
            const unsigned long long q = row0 + 32 * (i0 + j);
            in[j] = q < nq;
            rd[j] = in[j] ? __ldg(q_read + q) : 0ull;
            qs[j] = in[j] ? __ldg(q_start + q) : 0ull;
            qe[j] = in[j] ? __ldg(q_end + q) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kSlBatch; ++j) {
            const bool known = in[j] && rd[j] < n_reads;
            len[j] = known ? __ldg(lens + rd[j]) : 0ull;
            wo[j] = known ? __ldg(word_offsets + rd[j]) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kSlBatch; ++j) {
            const unsigned long long q = row0 + 32 * (i0 + j);
            unsigned long long cnt = 0, src = 0;
            if (in[j]) {
                if (rd[j] >= n_reads || qs[j] > qe[j] || qe[j] > len[j]) {
                    if (q < ld_volatile_u64(status)) atomicMin(status, q);
                } else {
                    cnt = qe[j] - qs[j];
                    // the word that holds base `start` and the base's place inside it: pass B goes straight to the words
                    src = ((wo[j] + (qs[j] >> 5)) << 5) | (qs[j] & 31u);
                }
            }
            s_cnt[i0 + j][threadIdx.x] = cnt < 0xFFFFFFFFull ? (uint32_t)cnt : 0xFFFFFFFFu;
            s_src[i0 + j][threadIdx.x] = src;
            wsum += cnt;
        }
    }
    wsum = warp_sum_u64(wsum);
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long a = lane < kWarpsPerBlock ? s_warp[lane] : 0ull;
        unsigned long long ia = a;
#pragma unroll
        for (int o = 1; o < kWarpsPerBlock; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, ia, o);
            if (lane >= (unsigned)o) ia += x;
        }
        const unsigned long long agg[1] = {__shfl_sync(0xffffffffu, ia, kWarpsPerBlock - 1)};
        unsigned long long excl[1];
        lookback_exclusive<1>(lb + 1, n_tiles, tile, agg, excl);
        if (lane < kWarpsPerBlock) s_warp[lane] = ia - a;   // exclusive prefix of the warp totals
        if (lane == 0) s_base = excl[0];
    }
    __syncthreads();
    // ---- pass B: row by row, the warp on its own
    unsigned long long row_base = s_base + s_warp[warp];
#pragma unroll 1
    for (int row = 0; row < kSlRows; ++row) {
    const unsigned long long q = row0 + 32 * row;
    if (row + 1 < kSlRows) {   // ask for the next row's words now: its cut then starts on L1 hits
        const unsigned long long nsrc = s_src[row + 1][threadIdx.x];
        const unsigned ncnt = s_cnt[row + 1][threadIdx.x];
        if (ncnt && ncnt <= kSliceShort) {
            prefetch_l1(words + (nsrc >> 5));
            prefetch_l1(words + ((nsrc + ncnt - 1) >> 5));
        }
    }
    unsigned long long cnt = s_cnt[row][threadIdx.x];
    if (cnt == 0xFFFFFFFFull) cnt = q_end[q] - q_start[q];
    unsigned long long inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    const unsigned long long oo = row_base + inc - cnt;   // where this query's bytes start
    row_base += __shfl_sync(0xffffffffu, inc, 31);
    if (q < nq) {
        out_offsets[q] = oo;
        if (q + 1 == nq) out_offsets[nq] = oo + cnt;
    }
    unsigned n = 0;
    unsigned long long s = 0;
    const uint64_t* w = words;
    if (cnt > kSliceShort) {
        long_list[atomicAdd(long_count, 1ull)] = q;
    } else if (cnt) {
        n = (unsigned)cnt;
        const unsigned long long src = s_src[row][threadIdx.x];
        s = src & 31u;              // the range starts at base s of the word at words[src >> 5]
        w = words + (src >> 5);
    }
    // the warp's output span [span_lo, span_hi) and its staged image; a span with a long range inside is too big to stage
    const unsigned long long span_lo = __shfl_sync(0xffffffffu, oo, 0);
    const unsigned long long span_hi = __shfl_sync(0xffffffffu, oo + cnt, 31);
    const uintptr_t g_lo = reinterpret_cast<uintptr_t>(out) + span_lo, g_base = g_lo & ~(uintptr_t)15;
    const bool staged = g_lo - g_base + (span_hi - span_lo) <= kSliceStage;
    if (n) {   // two call sites: behind one `staged ? shared : global` pointer every store is a generic ST behind a branch
        const unsigned phase = (unsigned)((reinterpret_cast<uintptr_t>(out) + oo) & 3u);   // global and staged addresses share it
        if (staged) slice_cut(w, s, n, phase, stage[warp] + (reinterpret_cast<uintptr_t>(out) + oo - g_base));
        else slice_cut(w, s, n, phase, out + oo);
    }
    // (A span can hold one long range and still fit the stage when its other queries are empty: the copy-out below
    // then also writes that range's bytes, with whatever the image holds.  slice_long_kernel runs after this kernel on
    // the same stream and writes the range itself, so the final bytes are right.)
    if (staged) {  // warp-uniform: store the staged span, 16 bytes per lane step; ragged ends byte-wise
        __syncwarp();
        const unsigned a = (unsigned)(g_lo - g_base), b = a + (unsigned)(span_hi - span_lo);   // valid bytes [a, b) of the image
        for (unsigned c = 16 * lane; c < b; c += 16 * 32) {
            if (c >= a && c + 16 <= b) {
                *reinterpret_cast<uint4*>(g_base + c) = *reinterpret_cast<const uint4*>(stage[warp] + c);
            } else {
                for (unsigned i = c < a ? a : c; i < c + 16 && i < b; ++i) *reinterpret_cast<uint8_t*>(g_base + i) = stage[warp][i];
            }
        }
        __syncwarp();   // the image is reused by the next row
    }
    }
}

// Long ranges: a warp per queued query, four bases (one 32-bit store) per lane step = 128 contiguous bytes per warp step.
__global__ void __launch_bounds__(kThreads)
slice_long_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets, const uint64_t* __restrict__ q_read,
                  const uint64_t* __restrict__ q_start, const uint64_t* __restrict__ q_end, uint8_t* __restrict__ out,
                  const uint64_t* __restrict__ out_offsets, const unsigned long long* __restrict__ long_count,
                  const unsigned long long* __restrict__ long_list) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long n_long = *long_count;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    for (unsigned long long i = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n_long; i += n_warps) {
        const unsigned long long q = long_list[i];
        const unsigned long long s = q_start[q], n = q_end[q] - s;
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(words + word_offsets[q_read[q]]);
        uint8_t* o = out + out_offsets[q];
        const unsigned head = (4u - (unsigned)(reinterpret_cast<uintptr_t>(o) & 3u)) & 3u;   // n > 64 > head
        if (lane < head) o[lane] = ascii1_at(w32, s + lane);
        const unsigned long long body = (n - head) / 4;
        uint32_t* o32 = reinterpret_cast<uint32_t*>(o + head);
        const unsigned long long b0 = s + head;
        for (unsigned long long j = lane; j < body; j += 32) o32[j] = ascii4_at(w32, b0 + 4 * j);
        const unsigned long long done = head + 4 * body;
        if (done + lane < n) o[done + lane] = ascii1_at(w32, s + done + lane);
    }
}

__global__ void __launch_bounds__(kThreads)
get_batch_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets, const uint64_t* __restrict__ lens,
                 unsigned long long n_reads, const uint64_t* __restrict__ q_read, const uint64_t* __restrict__ q_index,
                 unsigned long long nq, uint8_t* __restrict__ out, unsigned long long* __restrict__ status) {
    const unsigned long long q = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (q >= nq) return;
    const unsigned long long r = q_read[q], i = q_index[q];
    // lens[r] and word_offsets[r] leave together (asked for after the validity test, the word offset would start its
    // round trip only when the length has arrived: ptxas sinks a load below an exit that does not need it)
    const bool in_table = r < n_reads;
    const unsigned long long len = in_table ? __ldg(lens + r) : 0ull, wo = in_table ? __ldg(word_offsets + r) : 0ull;
    if (i >= len || wo == ~0ull) {   // (wo == ~0 never holds: it makes the exit depend on wo)
        out[q] = 0;
        if (q < ld_volatile_u64(status)) atomicMin(status, q);
        return;
    }
    const uint64_t x = __ldg(words + wo + (i >> 5));
    out[q] = (uint8_t)(0x54474341u >> (8 * ((x >> (2 * (i & 31))) & 3)));
}

// ticket + tile descriptors, then the queue of long ranges (count + one entry per query)
static inline size_t slice_lb_words(size_t nq) { return lookback_bytes(ceil_div(nq ? nq : 1, kSlTile), 1) / sizeof(unsigned long long); }
size_t slice_batch_scratch_bytes(size_t nq) { return (slice_lb_words(nq) + nq + 1) * sizeof(unsigned long long); }

cudaError_t launch_slice_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                               size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_start, const uint64_t* d_q_end, size_t nq,
                               uint8_t* d_out, uint64_t* d_out_offsets, unsigned long long* d_status, void* d_scratch, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (nq == 0) return cudaMemsetAsync(d_out_offsets, 0, sizeof(uint64_t), s);
    unsigned long long* lb = static_cast<unsigned long long*>(d_scratch);
    unsigned long long* long_count = lb + slice_lb_words(nq);
    const unsigned long long n_tiles = ceil_div(nq, kSlTile);
    e = cudaMemsetAsync(lb, 0, (slice_lb_words(nq) + 1) * sizeof(unsigned long long), s);   // ticket, descriptors, long_count
    if (e != cudaSuccess) return e;
    slice_short_kernel<<<(unsigned)n_tiles, kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_start, d_q_end, nq, d_out,
                                                             d_out_offsets, d_status, long_count, long_count + 1, lb, n_tiles);
    static const int per_sm = blocks_per_sm(slice_long_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
    slice_long_kernel<<<grid_for(ceil_div(nq, kWarpsPerBlock), resident), kThreads, 0, s>>>(d_words, d_word_offsets, d_q_read, d_q_start, d_q_end,
                                                                                          d_out, d_out_offsets, long_count, long_count + 1);
    return cudaGetLastError();
}

cudaError_t launch_get_batch(const DeviceInfo&, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                             size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_index, size_t nq, uint8_t* d_out,
                             unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || nq == 0) return e;
    get_batch_kernel<<<(unsigned)ceil_div(nq, kThreads), kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_index, nq,
                                                                          d_out, d_status);
    return cudaGetLastError();
}

}  // namespace bn
