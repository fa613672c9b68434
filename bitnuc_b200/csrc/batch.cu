// batch.cu -- variable-length read batches: offset-indexed ASCII reads -> per-read packed words (sm_100a).
//
// Replaces the caller-side loop `for read in reads { PackedSequence::new(read)? }`
// (/root/reference/src/sequence.rs:40-52 -> src/utils/packing/avx.rs:130-151): every read starts
// on a fresh 64-bit word, an empty read takes no words (sequence.rs:42-46).
//
// Two steps on the device:
//   1. word offsets = exclusive prefix sum of ceil(len/32) over the reads (block sums, one-CTA scan
//      of the sums, block-local scan + offset);
//   2. encode, one thread per 16-base HALF of an output word (see encode_batch_kernel): warps walk
//      consecutive output words, each WARP pulls its tiles of 512 words from a global atomic counter (no
//      CTA barrier anywhere), the owning read is found once per tile by a 32-ary warp search and then
//      advanced linearly.
// Work per warp step is uniform whatever the length mix (50 bp .. 10 kbp).  HBM-bound at
// 1 B/base in + 8 B per word out + 16 B per read of offsets.
#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

// words taken by read r: ceil(len/32); an empty read takes none
struct WordsOfRead {
    const uint64_t* offsets;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long r) const {
        return (offsets[r + 1] - offsets[r] + 31) / 32;
    }
};

// Index of the read owning output word w (the last r in [0, n_reads-1] with word_offsets[r] <= w), found by a
// whole warp: 32 probes per step instead of one, so the dependent chain is ~log32(n) loads long instead of
// log2(n).  All lanes return the answer.
__device__ __forceinline__ unsigned long long owner_read_warp(const uint64_t* __restrict__ wo, unsigned long long n_reads,
                                                              unsigned long long w) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long lo = 0, hi = n_reads - 1;  // invariant: wo[lo] <= w, answer in [lo, hi]
    while (hi > lo) {
        const unsigned long long step = (hi - lo) / 32 + 1;   // probes lo + (l+1)*step, l = 0..31, cover (lo, hi]
        const unsigned long long p = lo + (lane + 1ull) * step;
        const bool le = p <= hi && __ldg(wo + p) <= w;
        const unsigned k = __popc(__ballot_sync(0xffffffffu, le));  // probes are monotone: the first k are true
        const unsigned long long nlo = lo + k * step, nhi = nlo + step - 1;
        lo = nlo;
        hi = nhi < hi ? nhi : hi;
    }
    return lo;
}

constexpr int kWarpWords = 512;                          // consecutive output words per warp per tile
constexpr int kGroupWords = 16;                          // words per warp step: one 16-base half-word per lane

// Rare paths, out of line: byte-wise fetch of a lane's (<= 16) bytes, and the first invalid byte of them.
static __device__ __noinline__ uint4 batch_load_bytes(const uint8_t* p, int nb) {
    uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
    for (int j = 0; j < nb; ++j) w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)p[j] << (8 * (j & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
static __device__ __noinline__ void batch_report_invalid(const uint8_t* bytes, unsigned long long src, int nb, unsigned long long rb,
                                                         unsigned long long r, uint32_t* read_status, unsigned long long* status) {
    for (int j = 0; j < nb; ++j) {
        const uint32_t b = bytes[src + j];
        if (!byte_is_valid(b)) {
            report_invalid(status, src + j, b);
            if (read_status) atomicMin(read_status + r, (uint32_t)(src + j - rb));
            return;
        }
    }
}

// 16 bytes starting 4*WS + sh8/8 bytes into the aligned vector pair (v, n)
template <int WS>
__device__ __forceinline__ uint4 align16(uint4 v, uint4 n, unsigned sh8) {
    const uint32_t w[8] = {v.x, v.y, v.z, v.w, n.x, n.y, n.z, n.w};
    return make_uint4(__funnelshift_r(w[WS], w[WS + 1], sh8), __funnelshift_r(w[WS + 1], w[WS + 2], sh8),
                      __funnelshift_r(w[WS + 2], w[WS + 3], sh8), __funnelshift_r(w[WS + 3], w[WS + 4], sh8));
}
// 32 bytes starting s (0..15, per lane) bytes into the aligned vector triple (x, y, z): two levels of word
// selects, then one funnel shift per word.
__device__ __forceinline__ void align32_lane(uint4 x, uint4 y, uint4 z, unsigned s, uint4& lo, uint4& hi) {
    const uint32_t w[12] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, z.x, z.y, z.z, z.w};
    const bool q2 = s & 8u, q1 = s & 4u;
    uint32_t a[10], b[9];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = q2 ? w[i + 2] : w[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) b[i] = q1 ? a[i + 1] : a[i];
    const unsigned sh8 = 8u * (s & 3u);
    lo = make_uint4(__funnelshift_r(b[0], b[1], sh8), __funnelshift_r(b[1], b[2], sh8), __funnelshift_r(b[2], b[3], sh8),
                    __funnelshift_r(b[3], b[4], sh8));
    hi = make_uint4(__funnelshift_r(b[4], b[5], sh8), __funnelshift_r(b[5], b[6], sh8), __funnelshift_r(b[6], b[7], sh8),
                    __funnelshift_r(b[7], b[8], sh8));
}

// Fast path body: `n_iter` steps of U groups (16 complete words each) of one read.  p = this lane's first
// aligned vector, q = this lane's first output half-word; WS < 0 means the read starts 16-byte aligned.
// Returns the number of the first step (0-based) in which this lane saw an invalid byte, or ~0u.
template <int WS, int U>
__device__ __forceinline__ unsigned fast_groups(const uint4* __restrict__ p, uint32_t* __restrict__ q, unsigned n_iter, unsigned sh8) {
    unsigned first_bad = ~0u;
    for (unsigned it = 0; it < n_iter; ++it) {
        uint4 x[U], y[U];
#pragma unroll
        for (int j = 0; j < U; ++j) x[j] = ld128<LD_PLAIN>(p + 32 * j);
        if (WS >= 0) {
#pragma unroll
            for (int j = 0; j < U; ++j) y[j] = ld128<LD_PLAIN>(p + 32 * j + 1);
        }
        uint32_t bad = 0;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint4 v = WS >= 0 ? align16<(WS >= 0 ? WS : 0)>(x[j], y[j], sh8) : x[j];
            st_stream_u32(q + 32 * j, pack16(v, bad));
        }
        if ((bad & kValidMask) && first_bad == ~0u) first_bad = it;
        p += 32 * U;
        q += 32 * U;
    }
    return first_bad;
}

template <int U>
__device__ __forceinline__ unsigned fast_groups_any(const uint4* p, uint32_t* q, unsigned n_iter, unsigned s) {
    if (s == 0) return fast_groups<-1, U>(p, q, n_iter, 0);
    const unsigned sh8 = 8 * (s & 3u);
    switch (s >> 2) {
    case 0: return fast_groups<0, U>(p, q, n_iter, sh8);
    case 1: return fast_groups<1, U>(p, q, n_iter, sh8);
    case 2: return fast_groups<2, U>(p, q, n_iter, sh8);
    default: return fast_groups<3, U>(p, q, n_iter, sh8);
    }
}

// One thread per 16-base HALF of an output word.  A warp walks kWarpWords consecutive output words in groups
// of 16 words (32 half-words).  When the next group(s) lie inside one read, the 32 half-words are one
// contiguous, uniformly misaligned 512-byte run, fetched with two coalesced aligned 128-bit loads per lane and
// put in place by a funnel shift whose word part is a template constant (long reads live here).  Otherwise
// the next 32 words take the mixed pass, one word per lane: every lane finds the read owning its word (shuffle
// binary search over the next 32 read boundaries) and fetches its bytes with its own alignment, so short reads
// (several per group) keep all lanes busy too.
__global__ void __launch_bounds__(kThreads, 4)
encode_batch_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                    const uint64_t* __restrict__ word_offsets, uint64_t* __restrict__ out,
                    uint32_t* __restrict__ read_status, unsigned long long* __restrict__ status,
                    unsigned long long* __restrict__ tile_counter) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long total_words = word_offsets[n_reads];
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(bytes) + offsets[0];        // valid address range of the bytes
    const uintptr_t buf_hi = reinterpret_cast<uintptr_t>(bytes) + offsets[n_reads];
    const unsigned long long n_tiles = ceil_div(total_words, kWarpWords);
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    for (;;) {  // persistent WARPS pull tiles from a global counter: dynamic balance, no CTA barrier anywhere
        unsigned long long tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const unsigned long long ww0 = tile * kWarpWords;                             // this warp's words [ww0, ww1)
        const unsigned long long ww1 = ww0 + kWarpWords < total_words ? ww0 + kWarpWords : total_words;
        unsigned long long r = owner_read_warp(word_offsets, n_reads, ww0);          // warp-uniform
        unsigned long long wo_next = __ldg(word_offsets + r + 1);
        {   // The warp's source bytes are one contiguous run (reads are adjacent in the byte buffer) of at most
            // 32 bytes per word: pull it into L2 now, so the dependent per-group loads below see L2 latency,
            // not DRAM latency.
            const uintptr_t p0 = (reinterpret_cast<uintptr_t>(bytes) + __ldg(offsets + r) + (ww0 - __ldg(word_offsets + r)) * 32ull) & ~(uintptr_t)127;
            uintptr_t p1 = p0 + (ww1 - ww0) * 32ull + 128;
            if (p1 > buf_hi) p1 = buf_hi;
            for (uintptr_t p = p0 + 128ull * lane; p < p1; p += 128ull * 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
        unsigned long long wb = ww0;
        while (wb < ww1) {
            while (wo_next <= wb) wo_next = __ldg(word_offsets + (++r) + 1);  // read owning word wb (skips empty reads)
            // ---- fast path: groups of 16 complete words of read r.  The 32 lanes' bytes are one contiguous,
            // uniformly misaligned 512-byte run: two coalesced aligned 128-bit loads + a funnel shift each.
            if (wo_next - wb >= kGroupWords && ww1 - wb >= kGroupWords) {
                const unsigned long long rb = __ldg(offsets + r), re = __ldg(offsets + r + 1), wo = __ldg(word_offsets + r);
                const unsigned long long full_end = wo + (re - rb) / 32;                   // words [wo, full_end) are complete
                const unsigned s = (unsigned)((reinterpret_cast<uintptr_t>(bytes) + rb) & 15u);  // misalignment of every half
                unsigned long long lim = wo_next < ww1 ? wo_next : ww1;
                if (full_end < lim) lim = full_end;
                unsigned long long groups = lim > wb ? (lim - wb) / kGroupWords : 0;       // complete groups ahead in this read
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(bytes) + rb + (wb - wo) * 32ull - s;  // aligned start of the run
                if (a0 < buf_lo || a0 + 32 > buf_hi) groups = 0;                           // first vector of the batch: mixed pass
                else if (groups > (buf_hi - a0 - 32) / 512) groups = (buf_hi - a0 - 32) / 512;   // keep every aligned load inside the buffer
                if (groups) {
                    constexpr int U = 2;
                    const uint4* p = reinterpret_cast<const uint4*>(a0) + lane;
                    uint32_t* q = out32 + 2 * wb + lane;
                    const unsigned n2 = (unsigned)(groups / U), n1 = (unsigned)(groups % U);
                    unsigned fb = fast_groups_any<U>(p, q, n2, s);
                    if (fb != ~0u) fb *= U;
                    if (n1) {
                        const unsigned f1 = fast_groups_any<1>(p + 32 * U * n2, q + 32 * U * n2, n1, s);
                        if (fb == ~0u && f1 != ~0u) fb = U * n2 + f1;
                    }
                    if (fb != ~0u) {  // rare: this lane saw an invalid byte from group `fb` on (U groups checked together)
                        for (unsigned long long g = fb; g < groups; ++g) {
                            const unsigned long long src = rb + (wb + g * kGroupWords - wo) * 32ull + 16u * lane;
                            bool found = false;
                            for (int j = 0; j < 16 && !found; ++j) found = !byte_is_valid(bytes[src + j]);
                            if (found) {
                                batch_report_invalid(bytes, src, 16, rb, r, read_status, status);
                                break;
                            }
                        }
                    }
                    wb += groups * kGroupWords;
                    continue;
                }
            }
            // ---- mixed pass: the next (<= 32) output words, one per lane, whatever reads they belong to.  Lane j
            // fetches the word offset of read r+1+j (one coalesced load); every lane then finds the read owning
            // its word by a binary search over those 32 boundaries with shuffles (32-bit, relative to wb), and
            // fetches its 32 bases with its own alignment: <= 3 aligned vectors, two levels of word selects, one
            // funnel shift per word.  Bases past the end of a read are cut from the packed word (zero padding).
            const unsigned gcount = ww1 - wb < 32 ? (unsigned)(ww1 - wb) : 32u;
            const unsigned long long bidx = r + 1 + lane;
            const unsigned long long bj = bidx <= n_reads ? __ldg(word_offsets + bidx) : ~0ull;
            const unsigned long long first_w = wb - __ldg(word_offsets + r);            // index of word wb inside read r
            const unsigned long long d64 = bj - wb;                                      // > 0: read r owns word wb
            const unsigned bd = d64 > 0xFFFFFFFEull ? 0xFFFFFFFFu : (unsigned)d64;
            const bool active = lane < gcount;
            const bool overflow = __shfl_sync(0xffffffffu, bd, 31) < gcount;             // > 32 reads start here (empty reads): rare
            unsigned long long ri, wi;   // owning read, index of this lane's word inside it
            if (!overflow) {
                unsigned c = 0;  // number of boundaries <= lane
#pragma unroll
                for (unsigned st = 16; st; st >>= 1) {
                    const unsigned t = __shfl_sync(0xffffffffu, bd, c + st - 1);
                    if (t <= lane) c += st;
                }
                const unsigned prev = __shfl_sync(0xffffffffu, bd, (c + 31u) & 31u);     // first word of read r + c, relative to wb
                if (!active) c = 0;
                ri = r + c;
                wi = c ? (unsigned long long)(lane - prev) : first_w + lane;
            } else {
                unsigned long long lo = r, hi = n_reads - 1;
                if (active) {
                    while (lo < hi) {
                        const unsigned long long mid = lo + (hi - lo + 1) / 2;
                        if (__ldg(word_offsets + mid) <= wb + lane) lo = mid; else hi = mid - 1;
                    }
                }
                ri = lo;
                wi = wb + lane - __ldg(word_offsets + ri);
            }
            const unsigned long long rb_i = __ldg(offsets + ri), re_i = __ldg(offsets + ri + 1);
            const unsigned long long src = rb_i + wi * 32ull;                            // byte offset of this lane's 32 bases
            const unsigned nb = !active || src >= re_i ? 0u : (re_i - src < 32 ? (unsigned)(re_i - src) : 32u);
            if (nb) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(bytes) + src;
                const unsigned s = (unsigned)(a & 15u);
                const uintptr_t a0 = a - s;
                const unsigned nv = (s + nb + 15u) >> 4;                                 // aligned vectors holding the bases: 1..3
                uint4 lo4, hi4;
                if (a0 >= buf_lo && a0 + 16u * nv <= buf_hi) {
                    const uint4* pv = reinterpret_cast<const uint4*>(a0);
                    const uint4 x = ld128<LD_PLAIN>(pv);
                    const uint4 y = nv > 1 ? ld128<LD_PLAIN>(pv + 1) : x;
                    const uint4 z = nv > 2 ? ld128<LD_PLAIN>(pv + 2) : y;
                    align32_lane(x, y, z, s, lo4, hi4);
                } else {  // the aligned window pokes outside the byte buffer (first / last vector of the batch)
                    lo4 = batch_load_bytes(bytes + src, nb < 16 ? (int)nb : 16);
                    hi4 = batch_load_bytes(bytes + src + 16, nb > 16 ? (int)nb - 16 : 0);
                }
                // Bytes past the end of the read belong to the next read: they are cut from the packed word below
                // and, should one of them trip the validity test, the exact per-byte check sorts it out.
                uint32_t bad = 0;
                const uint32_t c_lo = pack16(lo4, bad), c_hi = pack16(hi4, bad);
                uint64_t w64 = ((uint64_t)c_hi << 32) | c_lo;
                if (nb < 32) w64 &= (1ull << (2 * nb)) - 1ull;
                out[wb + lane] = w64;
                if (bad & kValidMask) batch_report_invalid(bytes, src, (int)nb, rb_i, ri, read_status, status);
            }
            // advance to the read owning word wb + gcount
            if (wb + gcount < ww1) {
                if (!overflow) {
                    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, bd <= gcount));
                    const unsigned nx = __shfl_sync(0xffffffffu, bd, cnt & 31u);
                    r += cnt;
                    wo_next = cnt < 32 && nx != 0xFFFFFFFFu ? wb + nx : __ldg(word_offsets + r + 1);
                } else {
                    r = owner_read_warp(word_offsets, n_reads, wb + gcount);
                    wo_next = __ldg(word_offsets + r + 1);
                }
            }
            wb += gcount;
        }
    }
}

size_t encode_batch_scratch_bytes(size_t n_reads) {
    return scan_scratch_bytes(n_reads) + sizeof(unsigned long long);  // block sums + total, then the tile counter
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
    unsigned long long* tile_counter = sums + scan_scratch_bytes(n_reads) / sizeof(unsigned long long);
    e = cudaMemsetAsync(tile_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    launch_exclusive_scan(WordsOfRead{d_offsets}, n_reads, sums, d_out_word_offsets, s);
    static const int resident = resident_blocks(encode_batch_kernel, kThreads, di);
    // the number of output words is only known on the device: launch a full persistent grid
    encode_batch_kernel<<<resident, kThreads, 0, s>>>(d_bytes, d_offsets, n_reads, d_out_word_offsets, d_out_words,
                                                      d_read_status, d_status, tile_counter);
    return cudaGetLastError();
}

}  // namespace bn
