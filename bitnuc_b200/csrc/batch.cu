// batch.cu -- variable-length read batches: offset-indexed ASCII reads -> per-read packed words (sm_100a).
//
// Replaces the caller-side loop `for read in reads { PackedSequence::new(read)? }`
// (/root/reference/src/sequence.rs:40-52 -> src/utils/packing/avx.rs:130-151): every read starts
// on a fresh 64-bit word, an empty read takes no words (sequence.rs:42-46).
//
// Two steps on the device:
//   1. word offsets = exclusive prefix sum of ceil(len/32) over the reads (scan.cuh); the same pass notes, for
//      every tile of 2048 output words, the read that owns the tile's first word;
//   2. encode_batch_kernel, PACK THEN ALIGN.  The reads are adjacent in the byte buffer, so the bytes behind a
//      tile of output words are one contiguous span (<= 64 KiB).  Phase 1 streams that span exactly like the
//      contiguous encode -- aligned, coalesced 128-bit loads, 16 bases -> one 32-bit code, validated as a
//      whole -- and parks the codes in shared memory.  Phase 2 cuts the output words out of the code strip:
//      word j of a read is the 64-bit window starting 2 * (byte offset of its first base) bits into the strip
//      (three LDS + two funnel shifts), so the misalignment of a read costs nothing per byte.  Short reads are
//      assembled one read per thread, long ones 32 words per warp; no search anywhere.
// Work is uniform whatever the length mix (100 bp reads .. 10 kbp reads .. one chromosome).  HBM-bound at
// 1 B/base in + 8 B per word out + 16 B per read of offsets.
#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

constexpr int kLongWords = 64;                     // a read with more words than this inside the tile is cut into chunks

// words taken by read r: ceil(len/32); an empty read takes none
struct WordsOfRead {
    const uint64_t* offsets;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long r) const {
        return (offsets[r + 1] - offsets[r] + 31) / 32;
    }
};

// scan hook: read r owns words [start, start + count); note it as the owner of every tile whose first word it holds
template <int kTileWords>   // compile-time (a power of two): the division below is per read, it must stay a shift
struct NoteTileOwners {
    unsigned long long* tile_owner;
    unsigned long long max_tiles;
    __device__ __forceinline__ void operator()(unsigned long long r, unsigned long long start, unsigned long long count) const {
        if (count == 0) return;
        for (unsigned long long t = (start + kTileWords - 1) / kTileWords; t < max_tiles && t * kTileWords < start + count; ++t)
            tile_owner[t] = r;
    }
};

// Index of the read owning output word w (the last r in [0, n_reads-1] with word_offsets[r] <= w), found by a
// whole warp: 32 probes per step, so the dependent chain is ~log32(n) loads long.  All lanes return the answer.
// Only used for tiles beyond the owner table (a caller that under-declared n_bytes).
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

// Rare path, out of line: a vector at the edge of the byte buffer, fetched byte-wise ('A' outside [lo, hi)).
static __device__ __noinline__ uint4 batch_load_edge(uintptr_t a, uintptr_t lo, uintptr_t hi) {
    uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
    for (int j = 0; j < 16; ++j)
        if (a + j >= lo && a + j < hi)
            w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)*reinterpret_cast<const uint8_t*>(a + j) << (8 * (j & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Rare path, out of line: report every invalid byte of the vector at address a that lies inside the batch.
// status = min(offset << 8 | byte) over the batch; read_status[r] = min position inside read r.
static __device__ __noinline__ void batch_report_vector(const uint8_t* bytes, uintptr_t a, uintptr_t lo, uintptr_t hi,
                                                        const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                                                        uint32_t* read_status, unsigned long long* status) {
    for (int j = 0; j < 16; ++j) {
        if (a + j < lo || a + j >= hi) continue;
        const uint32_t b = *reinterpret_cast<const uint8_t*>(a + j);
        if (byte_is_valid(b)) continue;
        const unsigned long long off = (unsigned long long)(a + j - reinterpret_cast<uintptr_t>(bytes));
        report_invalid(status, off, b);
        if (read_status) {  // the read holding byte `off`: the last r with offsets[r] <= off
            unsigned long long l = 0, h = n_reads - 1;
            while (l < h) {
                const unsigned long long mid = l + (h - l + 1) / 2;
                if (offsets[mid] <= off) l = mid; else h = mid - 1;
            }
            atomicMin(read_status + l, (uint32_t)(off - offsets[l]));
        }
    }
}

// The output word whose first base sits `rel` bytes into the span: a 64-bit window of the code strip.
__device__ __forceinline__ uint64_t cut_word(const uint32_t* __restrict__ codes, unsigned rel) {
    const unsigned vi = rel >> 4, sh = 2u * (rel & 15u);
    const uint32_t c0 = codes[vi], c1 = codes[vi + 1], c2 = codes[vi + 2];
    return ((uint64_t)__funnelshift_r(c1, c2, sh) << 32) | __funnelshift_r(c0, c1, sh);
}

struct LongSeg {
    unsigned rel;    // byte offset of the segment's first base inside the span
    unsigned first;  // its first word, relative to the tile
    unsigned count;  // words
    unsigned tail;   // bases in its last word (32 unless that is the read's ragged last word)
};

// kTileWords = output words per CTA tile (32 bytes of bases each), kBThreads = CTA size, kMinCtas = CTAs per SM,
// kPackU = independent 128-bit loads per thread in phase 1
template <int kTileWords, int kBThreads, int kMinCtas, int kPackU>
__global__ void __launch_bounds__(kBThreads, kMinCtas)
encode_batch_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                    const uint64_t* __restrict__ word_offsets, uint64_t* __restrict__ out,
                    uint32_t* __restrict__ read_status, unsigned long long* __restrict__ status,
                    unsigned long long* __restrict__ tile_counter, const unsigned long long* __restrict__ tile_owner,
                    unsigned long long max_tiles) {
    constexpr int kStripCodes = 2 * kTileWords + 8;    // one 32-bit code per aligned 16-byte vector of the span, + slack
    constexpr int kLongCap = kTileWords / kLongWords + 1;
    constexpr int kThreads = kBThreads, kWarpsPerBlock = kBThreads / 32;
    __shared__ uint32_t codes[kStripCodes];
    __shared__ LongSeg segs[kLongCap];
    __shared__ uint16_t chunks[kTileWords / 32 + kLongCap];   // (segment << 8 | chunk of 32 words), <= 64 chunks per segment
    __shared__ unsigned n_segs, n_chunks;
    __shared__ unsigned long long tile_s, r_s[2];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long total_words = word_offsets[n_reads];
    const unsigned long long n_tiles = ceil_div(total_words, kTileWords);
    const uintptr_t base = reinterpret_cast<uintptr_t>(bytes);
    const uintptr_t buf_lo = base + offsets[0], buf_hi = base + offsets[n_reads];   // valid address range of the bytes
    // (Drawing the NEXT tile's ticket while the current tile is packed -- the atomic's round trip and the barrier that
    // publishes it off the critical path -- was measured: 6.70 against 6.77 ms on the cfg 5 mix, 1.43 against 1.42 ms on
    // short reads, and the ticket carried across the loop spills at this kernel's 48-register cap.  Not kept.)
    for (;;) {  // persistent CTAs pull tiles from a global counter (the word count is only known on the device)
        __syncthreads();  // the previous tile is done with codes / segs
        if (tid == 0) {
            tile_s = atomicAdd(tile_counter, 1ull);
            n_segs = 0;
            n_chunks = 0;
        }
        __syncthreads();
        const unsigned long long tile = tile_s;
        if (tile >= n_tiles) break;
        const unsigned long long w0 = tile * kTileWords;
        const bool last = tile + 1 == n_tiles;
        const unsigned long long w1 = last ? total_words : w0 + kTileWords;
        // reads of the tile: r0 owns word w0; r1 owns word w1 (it may start right there), or is the last read
        unsigned long long r0, r1;
        if (tile + 1 < max_tiles) {
            r0 = tile_owner[tile];
            r1 = last ? n_reads - 1 : tile_owner[tile + 1];
        } else {  // beyond the owner table: search
            if (warp < 2 && (warp == 0 || !last)) {
                const unsigned long long rr = owner_read_warp(word_offsets, n_reads, warp ? w1 : w0);
                if (lane == 0) r_s[warp] = rr;
            }
            __syncthreads();
            r0 = r_s[0];
            r1 = last ? n_reads - 1 : r_s[1];
        }
        // the tile's bytes: [p0, p1) -- one contiguous span, reads are adjacent in the buffer
        const uintptr_t p0 = base + offsets[r0] + (w0 - word_offsets[r0]) * 32ull;
        const uintptr_t p1 = last ? buf_hi : base + offsets[r1] + (w1 - word_offsets[r1]) * 32ull;
        const uintptr_t a0 = p0 & ~(uintptr_t)15;
        const unsigned nvec = (unsigned)((p1 - a0 + 15) >> 4);                       // <= 2 * kTileWords + 1
        // phase 2a's first read of this thread: fetch its offsets now, so that their latency hides behind phase 1
        unsigned long long m_wo = 0, m_rb = 0;
        unsigned m_nw = 0, m_len = 0;
        if (r0 + tid <= r1) {
            m_wo = __ldg(word_offsets + r0 + tid), m_rb = __ldg(offsets + r0 + tid);
            const unsigned long long nw = __ldg(word_offsets + r0 + tid + 1) - m_wo, len = __ldg(offsets + r0 + tid + 1) - m_rb;
            m_nw = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)nw;   // 0xFFFFFFFF: too long to keep here, re-fetched below
            m_len = (unsigned)len;
        }
        // ---- phase 1: pack the span, 16 bases per thread step, into the code strip
        {
            const uint4* src = reinterpret_cast<const uint4*>(a0);
            const bool interior = a0 >= buf_lo && a0 + 16ull * nvec <= buf_hi;
            unsigned v = tid;
            if (interior) {
                for (; v + (kPackU - 1) * kThreads < nvec; v += kPackU * kThreads) {
                    uint4 x[kPackU];
#pragma unroll
                    for (int j = 0; j < kPackU; ++j) x[j] = ld128<LD_NC_NOALLOC>(src + v + j * kThreads);
                    uint32_t bad = 0;
#pragma unroll
                    for (int j = 0; j < kPackU; ++j) codes[v + j * kThreads] = pack16(x[j], bad);
                    if (bad & kValidMask) {
#pragma unroll 1
                        for (int j = 0; j < kPackU; ++j)
                            batch_report_vector(bytes, a0 + 16ull * (v + j * kThreads), buf_lo, buf_hi, offsets, n_reads, read_status, status);
                    }
                }
            }
            for (; v < nvec; v += kThreads) {  // leftover vectors, and every vector of a tile at the edge of the buffer
                const uintptr_t a = a0 + 16ull * v;
                const uint4 x = a >= buf_lo && a + 16 <= buf_hi ? ld128<LD_NC_NOALLOC>(src + v) : batch_load_edge(a, buf_lo, buf_hi);
                uint32_t bad = 0;
                codes[v] = pack16(x, bad);
                if (bad & kValidMask) batch_report_vector(bytes, a, buf_lo, buf_hi, offsets, n_reads, read_status, status);
            }
        }
        __syncthreads();
        // ---- phase 2a: one read per thread; a read's words are consecutive 64-bit windows of the strip, 32 bytes apart
        // (software-pipelined: the offsets of the thread's next read are in flight while it cuts the current one)
        unsigned long long c_wo, c_wn, c_rb, c_re;
        if (m_nw != 0xFFFFFFFFu) {
            c_wo = m_wo, c_wn = m_wo + m_nw, c_rb = m_rb, c_re = m_rb + m_len;   // fetched before phase 1
        } else {
            c_wo = __ldg(word_offsets + r0 + tid), c_wn = __ldg(word_offsets + r0 + tid + 1);
            c_rb = __ldg(offsets + r0 + tid), c_re = __ldg(offsets + r0 + tid + 1);
        }
        for (unsigned long long r = r0 + tid; r <= r1; r += kThreads) {
            const unsigned long long wo_r = c_wo, wo_n = c_wn, rb = c_rb, re = c_re;
            if (r + kThreads <= r1) {
                c_wo = __ldg(word_offsets + r + kThreads), c_wn = __ldg(word_offsets + r + kThreads + 1);
                c_rb = __ldg(offsets + r + kThreads), c_re = __ldg(offsets + r + kThreads + 1);
            }
            const unsigned long long wf = wo_r > w0 ? wo_r : w0, wl = wo_n < w1 ? wo_n : w1;   // its words inside the tile
            if (wf >= wl) continue;
            unsigned rel = (unsigned)(base + rb + (wf - wo_r) * 32ull - a0);
            const unsigned cnt = (unsigned)(wl - wf);
            const unsigned tail = wl == wo_n ? (unsigned)(re - rb - (wo_n - 1 - wo_r) * 32ull) : 32u;
            if (cnt > kLongWords) {  // long: leave it to the warps, 32 words at a time
                const unsigned slot = atomicAdd(&n_segs, 1u);
                const unsigned nch = (cnt + 31u) / 32u;
                const unsigned cb = atomicAdd(&n_chunks, nch);
                segs[slot] = LongSeg{rel, (unsigned)(wf - w0), cnt, tail};
                for (unsigned c = 0; c < nch; ++c) chunks[cb + c] = (uint16_t)(slot << 8 | c);
                continue;
            }
            uint64_t* o = out + wf;
            for (unsigned j = 0; j + 1 < cnt; ++j, rel += 32) o[j] = cut_word(codes, rel);
            uint64_t w = cut_word(codes, rel);
            if (tail < 32) w &= (1ull << (2 * tail)) - 1ull;   // zero padding of the read's last word
            o[cnt - 1] = w;
        }
        __syncthreads();
        // ---- phase 2b: long segments, one chunk of 32 consecutive words per warp step (coalesced 256-byte stores)
        const unsigned nc = n_chunks;
        for (unsigned k = warp; k < nc; k += kWarpsPerBlock) {
            const unsigned ch = chunks[k];
            const LongSeg sg = segs[ch >> 8];
            const unsigned j = (ch & 0xFFu) * 32u + lane;
            if (j < sg.count) {
                uint64_t w = cut_word(codes, sg.rel + 32u * j);
                if (j + 1 == sg.count && sg.tail < 32) w &= (1ull << (2 * sg.tail)) - 1ull;
                out[w0 + sg.first + j] = w;
            }
        }
    }
}

constexpr int kMinTileWords = 2048;  // the owner table is sized for the smallest tile any variant uses

static inline unsigned long long batch_max_tiles(size_t n_reads, size_t n_bytes, unsigned tile_words) {
    return ceil_div((unsigned long long)n_bytes / 32 + n_reads, tile_words) + 2;
}

size_t encode_batch_scratch_bytes(size_t n_reads, size_t n_bytes) {
    // block sums + total, the tile counter, then the tile-owner table
    return scan_scratch_bytes(n_reads) + sizeof(unsigned long long) * (1 + batch_max_tiles(n_reads, n_bytes, kMinTileWords));
}

template <int kTileWords, int kBThreads, int kMinCtas, int kPackU = 4>
static cudaError_t launch_encode_batch_variant(const DeviceInfo& di, const uint8_t* d_bytes, const uint64_t* d_offsets, size_t n_reads,
                                               size_t n_bytes, uint64_t* d_out_words, uint64_t* d_out_word_offsets, uint32_t* d_read_status,
                                               unsigned long long* d_status, unsigned long long* sums, unsigned long long* tile_counter,
                                               unsigned long long* tile_owner, cudaStream_t s) {
    const unsigned long long max_tiles = batch_max_tiles(n_reads, n_bytes, kTileWords);
    launch_exclusive_scan(WordsOfRead{d_offsets}, n_reads, sums, d_out_word_offsets, s, NoteTileOwners<kTileWords>{tile_owner, max_tiles});
    static const int per_sm = blocks_per_sm(encode_batch_kernel<kTileWords, kBThreads, kMinCtas, kPackU>, kBThreads);
    const int resident = per_sm * di.sm_count;
    // the number of output words is only known on the device: launch a full persistent grid.  (One tile per CTA over a grid
    // sized for the upper bound of the word count -- the form that took fastq_lines_kernel from 1.33 to 1.12 ms -- was measured
    // here too: 7.76 against 6.77 ms on the cfg 5 mix, 1.53 against 1.42 ms on short reads; this kernel sits at its register
    // cap either way and the empty CTAs past the real tile count cost more than the ticket.)
    encode_batch_kernel<kTileWords, kBThreads, kMinCtas, kPackU><<<resident, kBThreads, 0, s>>>(
        d_bytes, d_offsets, n_reads, d_out_word_offsets, d_out_words, d_read_status, d_status, tile_counter, tile_owner, max_tiles);
    return cudaGetLastError();
}

cudaError_t launch_encode_batch(const DeviceInfo& di, const uint8_t* d_bytes, const uint64_t* d_offsets,
                                size_t n_reads, size_t n_bytes, uint64_t* d_out_words, uint64_t* d_out_word_offsets,
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
    unsigned long long* tile_owner = tile_counter + 1;
    e = cudaMemsetAsync(tile_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    // tile / CTA shape from an on-device sweep (profiles/r01_sweep_encode_batch.txt): 2048-word tiles, 128 threads
    // (32 vectors per thread per tile), 10 CTAs per SM
    return launch_encode_batch_variant<2048, 128, 10>(di, d_bytes, d_offsets, n_reads, n_bytes, d_out_words, d_out_word_offsets,
                                                      d_read_status, d_status, sums, tile_counter, tile_owner, s);
}

}  // namespace bn
