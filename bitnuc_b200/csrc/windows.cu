// windows.cu -- every k-mer of a sequence: `for w in seq.windows(k) { as_2bit(w)? }` as one kernel (sm_100a).
// SURVEY.md 8(f) rank 3; the usage pattern is /root/reference/README.md:160-180, the per-window function
// /root/reference/src/utils/packing/mod.rs:81-110.
//
// out[i] = as_2bit(seq[i .. i+k]) for i in [0, n-k].  Pack once, cut many: a CTA owns 2048 consecutive windows; it
// packs the 2048 + k - 1 bases behind them exactly like the streaming encode (aligned 128-bit loads, 16 bases -> one
// 32-bit code, validated as whole vectors) into a shared-memory code strip, then every thread cuts windows out of
// the strip -- window i is the 2k-bit field starting 2i bits into it (three LDS + two funnel shifts + a mask) -- and
// the warp stores 32 consecutive words.  HBM-bound on the output: 1 byte in, 8 bytes out per base.
//
// Errors (the caller's loop with `?`): n < k -> no windows, nothing is looked at; k > 32 -> SequenceTooLong(k)
// (checked on the host); the first window holding an invalid byte reports InvalidBase(byte) -- that byte is the
// first invalid byte of the sequence, so the status word is the usual min(offset << 8 | byte).
#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

constexpr int kWinThreads = 128;
constexpr int kWinTile = 2048;                                  // windows per CTA
constexpr int kWinStrip = (kWinTile + 32 + 15 + 15) / 16 + 4;  // codes: vectors of the span + slack

static __device__ __noinline__ uint4 win_load_edge(const uint8_t* seq, long long off, unsigned long long n) {
    uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};  // 'A' outside the sequence
    for (int j = 0; j < 16; ++j)
        if (off + j >= 0 && (unsigned long long)(off + j) < n)
            w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)seq[off + j] << (8 * (j & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
static __device__ __noinline__ void win_report(uint4 v, long long off, unsigned long long n, unsigned long long* status) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 16; ++j) {
        const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (off + j >= 0 && (unsigned long long)(off + j) < n && !byte_is_valid(b)) {
            report_invalid(status, (unsigned long long)(off + j), b);
            return;
        }
    }
}

__global__ void __launch_bounds__(kWinThreads)
kmer_windows_kernel(const uint8_t* __restrict__ seq, unsigned long long n, unsigned k, uint64_t* __restrict__ out,
                    unsigned long long* __restrict__ status) {
    __shared__ uint32_t codes[kWinStrip];
    const unsigned tid = threadIdx.x;
    const unsigned long long n_win = n - k + 1;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kWinTile;        // first window of the CTA
    const unsigned cnt = (unsigned)(n_win - i0 < kWinTile ? n_win - i0 : kWinTile);
    const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(seq) + i0) & 15u);
    const long long a0 = (long long)i0 - mis;                                       // sequence offset of the strip's first byte
    const unsigned nvec = (mis + cnt + k - 1 + 15u) / 16u;                          // <= 132
    for (unsigned v = tid; v < nvec; v += kWinThreads) {
        const long long off = a0 + 16ll * v;
        const uint4 x = off >= 0 && (unsigned long long)off + 16 <= n ? ld128<LD_NC_NOALLOC>(reinterpret_cast<const uint4*>(seq + off))
                                                                       : win_load_edge(seq, off, n);
        uint32_t bad = 0;
        codes[v] = pack16(x, bad);
        if (bad & kValidMask) win_report(x, off, n, status);
    }
    __syncthreads();
    const uint32_t mlo = k >= 16 ? 0xFFFFFFFFu : (1u << (2 * k)) - 1u;
    const uint32_t mhi = k >= 32 ? 0xFFFFFFFFu : k <= 16 ? 0u : (1u << (2 * k - 32)) - 1u;
    uint2* o = reinterpret_cast<uint2*>(out + i0);
#pragma unroll 4
    for (unsigned i = tid; i < cnt; i += kWinThreads) {
        const unsigned rel = mis + i;
        const unsigned vi = rel >> 4, sh = 2u * (rel & 15u);
        const uint32_t c0 = codes[vi], c1 = codes[vi + 1], c2 = codes[vi + 2];
        st_stream_v2(o + i, make_uint2(__funnelshift_r(c0, c1, sh) & mlo, __funnelshift_r(c1, c2, sh) & mhi));
    }
}

// ============================================================================ per-read windows ==
// The same for a batch of reads (offset-indexed ASCII, as in encode_batch): out_offsets = exclusive scan of
// max(0, len - k + 1), windows never cross a read.  Tiles are cut over the INPUT bytes: a CTA owns 4096 consecutive
// bytes of the batch, packs them (+ the k-1 bytes after them) into a code strip, and every warp walks 1024 of those
// byte positions 32 at a time (lane = position, so the stores of a warp are consecutive words except across a read
// end).  A lane finds the read of its first position by a binary search bounded by the tile-owner table (the scan
// hook notes which read holds the first byte of every tile) and then only advances.
// Reads shorter than k have no window, so their bytes are never validated (`windows(k)` yields nothing for them).

constexpr int kWinBatchTile = 4096;   // input bytes per CTA
constexpr int kWinBatchStrip = (kWinBatchTile + 32 + 15 + 15) / 16 + 4;
constexpr int kWinBatchReads = 512;   // reads of a tile whose offsets are kept in shared memory

struct WindowsOfRead {
    const uint64_t* offsets;
    unsigned long long k;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long r) const {
        const unsigned long long len = offsets[r + 1] - offsets[r];
        return len >= k ? len - k + 1 : 0;
    }
};

// scan hook: read r holds bytes [offsets[r], offsets[r+1]); note it as the owner of every byte tile whose first byte it holds
struct NoteByteTileOwners {
    const uint64_t* offsets;
    unsigned long long* tile_owner;
    unsigned long long max_tiles;
    __device__ __forceinline__ void operator()(unsigned long long r, unsigned long long, unsigned long long) const {
        const unsigned long long b0 = offsets[r] - offsets[0], b1 = offsets[r + 1] - offsets[0];
        for (unsigned long long t = ceil_div(b0, kWinBatchTile); t < max_tiles && t * kWinBatchTile < b1; ++t) tile_owner[t] = r;
    }
};

// rare path: report the invalid bytes of the vector at sequence offset `off` that belong to reads with >= k bases
static __device__ __noinline__ void win_batch_report(uint4 v, long long off, const uint64_t* __restrict__ offsets,
                                                     unsigned long long n_reads, unsigned k, unsigned long long* status) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 16; ++j) {
        const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        const long long o = off + j;
        if (byte_is_valid(b) || o < (long long)offsets[0] || o >= (long long)offsets[n_reads]) continue;
        unsigned long long l = 0, h = n_reads - 1;   // the read holding byte o: the last r with offsets[r] <= o
        while (l < h) {
            const unsigned long long mid = l + (h - l + 1) / 2;
            if (offsets[mid] <= (unsigned long long)o) l = mid; else h = mid - 1;
        }
        if (offsets[l + 1] - offsets[l] >= k) report_invalid(status, (unsigned long long)o, b);
    }
}

__global__ void __launch_bounds__(kWinThreads)
kmer_windows_batch_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offsets, unsigned long long n_reads,
                          unsigned k, const uint64_t* __restrict__ out_offsets, uint64_t* __restrict__ out,
                          const unsigned long long* __restrict__ tile_owner, unsigned long long n_tiles,
                          unsigned long long* __restrict__ status) {
    __shared__ uint32_t codes[kWinBatchStrip];
    __shared__ int s_end[kWinBatchReads];
    __shared__ uint64_t* s_obase[kWinBatchReads];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the four values everything else depends on, in one trip to memory (the owner entries of a tile past the end are
    // allocated but unset: such a CTA returns before using them)
    const unsigned long long b_lo = __ldg(offsets), b_hi = __ldg(offsets + n_reads);  // the batch's bytes
    const unsigned long long own0 = __ldg(tile_owner + blockIdx.x), own1 = __ldg(tile_owner + blockIdx.x + 1);
    const unsigned long long t_lo = b_lo + (unsigned long long)blockIdx.x * kWinBatchTile;
    // the grid is sized from the caller's bound on the byte count.  (The second condition is never true for a tile inside the
    // batch -- own0 is a read index there.  It makes the exit DEPEND on the owner entries: without it ptxas sinks their two
    // loads below the exit, i.e. behind the round trip of the batch bounds, whatever the order they are written in.)
    if (t_lo >= b_hi || (own0 & own1) == ~0ull) return;
    n_tiles = ceil_div(b_hi - b_lo, kWinBatchTile);
    const unsigned long long t_hi = t_lo + kWinBatchTile < b_hi ? t_lo + kWinBatchTile : b_hi;
    const unsigned long long s_hi = t_hi + k - 1 < b_hi ? t_hi + k - 1 : b_hi;        // the strip also holds the k-1 bytes after the tile
    const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(bytes) + t_lo) & 15u);
    const long long a0 = (long long)t_lo - mis;                                      // batch offset of the strip's first byte
    const unsigned nvec = (unsigned)((s_hi - a0 + 15) / 16);
    // the thread's (up to three) vectors of the strip first: their addresses need the batch bounds only, so they leave
    // before the owner entries have arrived; then its first table entry, which needs them; then the packing
    constexpr int kVecPerThread = (kWinBatchStrip + kWinThreads - 1) / kWinThreads;
    uint4 x[kVecPerThread];
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
        const unsigned v = tid + j * kWinThreads;
        const long long off = a0 + 16ll * v;
        x[j] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        if (v < nvec) {
            if (off >= (long long)b_lo && (unsigned long long)off + 16 <= b_hi) {
                x[j] = ld128<LD_NC_NOALLOC>(reinterpret_cast<const uint4*>(bytes + off));
            } else {  // a vector at the edge of the batch: byte-wise, 'A' outside
                uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
                for (int b = 0; b < 16; ++b)
                    if (off + b >= (long long)b_lo && (unsigned long long)(off + b) < b_hi)
                        w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | ((uint32_t)bytes[off + b] << (8 * (b & 3)));
                x[j] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    // reads of the tile: r_first holds byte t_lo, r_last holds byte t_hi (or is the last read)
    const unsigned long long r_first = own0, r_last = blockIdx.x + 1 < n_tiles ? own1 : n_reads - 1;
    const bool table = r_last - r_first < (unsigned long long)kWinBatchReads;       // the tile's reads fit the shared table
    const unsigned nr = table ? (unsigned)(r_last - r_first) + 1 : 0u;
    unsigned long long e_st = 0, e_en = 0, e_out = 0;   // a 4 KiB tile of 125-bp reads has ~33 entries
    if (tid < nr) e_st = __ldg(offsets + r_first + tid), e_en = __ldg(offsets + r_first + tid + 1), e_out = __ldg(out_offsets + r_first + tid);
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
        const unsigned v = tid + j * kWinThreads;
        if (v < nvec) {
            uint32_t bad = 0;
            codes[v] = pack16(x[j], bad);
            if (bad & kValidMask) win_batch_report(x[j], a0 + 16ll * v, offsets, n_reads, k, status);
        }
    }
    // per read of the table: its end relative to the tile (clamped: only compared with positions < 4096) and the address
    // of the window that would start at tile position 0
    for (unsigned i = tid; i < nr; i += kWinThreads) {
        if (i != tid) e_st = __ldg(offsets + r_first + i), e_en = __ldg(offsets + r_first + i + 1), e_out = __ldg(out_offsets + r_first + i);
        const long long st = (long long)(e_st - t_lo), en = (long long)(e_en - t_lo);
        s_end[i] = en > (1 << 30) ? (1 << 30) : (int)en;
        s_obase[i] = out + e_out - st;
    }
    __syncthreads();
    const uint32_t mlo = k >= 16 ? 0xFFFFFFFFu : (1u << (2 * k)) - 1u;
    const uint32_t mhi = k >= 32 ? 0xFFFFFFFFu : k <= 16 ? 0u : (1u << (2 * k - 32)) - 1u;
    const unsigned long long w_end = t_lo + 1024ull * (warp + 1) < t_hi ? t_lo + 1024ull * (warp + 1) : t_hi;
    unsigned long long p = t_lo + 1024ull * warp + lane;                             // this lane's first byte position
    const uint2 keep = make_uint2(mlo, mhi);
    if (table) {
        // lanes search and advance in the shared table, and a window costs one add for its address
        // The warp owns the windows that start in its 1024 bytes of the tile.  It walks the reads overlapping that
        // range one after the other (warp-uniform), and its lanes stride over the window starts of the current read:
        // no per-lane bookkeeping, consecutive lanes -> consecutive output words.
        const int lo = 1024 * (int)warp, hi = (int)(w_end - t_lo), kk = (int)k;
        if (lo >= hi) return;
        unsigned i = 0, top = nr - 1;                                                // the first read that ends after lo
        while (i < top) {
            const unsigned mid = (i + top) / 2;
            if (s_end[mid] > lo) top = mid; else i = mid + 1;
        }
        int r_lo = lo;                                                               // that read starts at or before lo
        for (; i < nr; ++i) {
            const int r_hi = s_end[i];
            uint64_t* obase = s_obase[i];
            const int a = r_lo > lo ? r_lo : lo, b = r_hi - kk + 1 < hi ? r_hi - kk + 1 : hi;   // window starts [a, b)
            // a lane's windows are 32 positions apart: the same shift every time, two code words further on, 256 bytes further
            // out -- written with exactly those three induction variables (ncu: the kernel is issue-bound, and the loop as the
            // compiler derived it from `q` spent 18 instructions per trip, 5 of them re-deriving these)
            const int q0 = a + (int)lane;
            if (q0 < b) {
                const unsigned rel = (unsigned)q0 + mis, sh = 2u * (rel & 15u);
                const uint32_t* cp = codes + (rel >> 4);
                uint2* op = reinterpret_cast<uint2*>(obase + q0);
                int trips = (b - q0 + 31) >> 5;
#pragma unroll 1   // three or four trips per 125-bp read: an unrolled body with its remainder ladder costs more than it saves
                do {
                    const uint32_t c0 = cp[0], c1 = cp[1], c2 = cp[2];
                    st_stream_v2(op, make_uint2(__funnelshift_r(c0, c1, sh) & keep.x, __funnelshift_r(c1, c2, sh) & keep.y));
                    cp += 2;
                    op += 32;
                } while (--trips);
            }
            if (r_hi >= hi) break;
            r_lo = r_hi;
        }
        return;
    }
    // more reads in the tile than the table holds (runs of empty or tiny reads): the same walk on global memory
    if (p >= w_end) return;
    unsigned long long r = r_first, hi = r_last;                                     // the last r in [r_first, r_last] with offsets[r] <= p
    while (r < hi) {
        const unsigned long long mid = r + (hi - r + 1) / 2;
        if (__ldg(offsets + mid) <= p) r = mid; else hi = mid - 1;
    }
    unsigned long long r_lo = __ldg(offsets + r), r_hi = __ldg(offsets + r + 1), r_out = __ldg(out_offsets + r);
    for (; p < w_end; p += 32) {
        while (p >= r_hi) {
            ++r;
            r_lo = r_hi;
            r_hi = __ldg(offsets + r + 1);
            r_out = __ldg(out_offsets + r);
        }
        if (p + k <= r_hi) {
            const unsigned rel = (unsigned)(p - a0);
            const unsigned vi = rel >> 4, sh = 2u * (rel & 15u);
            const uint32_t c0 = codes[vi], c1 = codes[vi + 1], c2 = codes[vi + 2];
            st_stream_v2(reinterpret_cast<uint2*>(out + r_out + (p - r_lo)),
                         make_uint2(__funnelshift_r(c0, c1, sh) & keep.x, __funnelshift_r(c1, c2, sh) & keep.y));
        }
    }
}

size_t kmer_windows_batch_scratch_bytes(size_t n_reads, size_t n_bytes) {
    return scan_scratch_bytes(n_reads) + (ceil_div(n_bytes ? n_bytes : 1, kWinBatchTile) + 2) * sizeof(unsigned long long);
}

cudaError_t launch_kmer_windows_batch(const DeviceInfo&, const uint8_t* d_bytes, const uint64_t* d_offsets, size_t n_reads, size_t n_bytes,
                                      uint32_t k, uint64_t* d_out, uint64_t* d_out_offsets, unsigned long long* d_status, void* d_scratch,
                                      cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_reads == 0) return cudaMemsetAsync(d_out_offsets, 0, sizeof(uint64_t), s);
    unsigned long long* sums = static_cast<unsigned long long*>(d_scratch);
    unsigned long long* tile_owner = sums + scan_scratch_bytes(n_reads) / sizeof(unsigned long long);
    const unsigned long long n_tiles = ceil_div(n_bytes, kWinBatchTile);   // the caller's bound on offsets[n] - offsets[0]
    launch_exclusive_scan(WindowsOfRead{d_offsets, k}, n_reads, sums, d_out_offsets, s, NoteByteTileOwners{d_offsets, tile_owner, n_tiles + 1});
    if (n_tiles) kmer_windows_batch_kernel<<<(unsigned)n_tiles, kWinThreads, 0, s>>>(d_bytes, d_offsets, n_reads, k, d_out_offsets, d_out,
                                                                                      tile_owner, n_tiles, d_status);
    return cudaGetLastError();
}

cudaError_t launch_kmer_windows(const DeviceInfo&, const uint8_t* d_seq, size_t n, uint32_t k, uint64_t* d_out,
                                unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || k == 0 || n < k) return e;
    const unsigned long long n_win = n - k + 1;
    kmer_windows_kernel<<<(unsigned)ceil_div(n_win, kWinTile), kWinThreads, 0, s>>>(d_seq, n, k, d_out, d_status);
    return cudaGetLastError();
}

}  // namespace bn
