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
// Every output word depends on at most two input words: one lane per read (reads on this path are short: barcode |
// insert splits, benches/functions_benchmark.rs:59 uses 30-280 bases), in ONE launch that also places the outputs
// (split_packed_fused_kernel below: 2048-read tiles, decoupled look-back, rows software-pipelined).
// Status word = min over failing reads of (read index << 1 | kind): kind 0 = idx > len ->
// IndexOutOfBounds{idx, len} (split.rs:22-27), kind 1 = 0 < idx < len and a non-empty ebuf shorter than ceil(len/32) words.
#include "common.cuh"
#include "launch.cuh"
#include "lookback.cuh"

namespace bn {

// split_one: the words of one read, written to lo / ro (global memory or the warp's staged image of its output spans).
// A read of up to kSpPre words is fetched with kSpPre independent predicated loads BEFORE anything is stored: the plain
// loop (load a word, store a word) serialises one L2 / HBM round trip per word -- the kernel was latency-bound on it.
constexpr int kSpPre = 8;
__device__ __forceinline__ void split_one(const uint64_t* __restrict__ w, unsigned long long nw, unsigned long long slen,
                                          unsigned long long i, uint64_t* __restrict__ lo, uint64_t* __restrict__ ro) {
    if (nw == 0) return;
    // c = index of the word holding base i; left takes words [0, c] (word c masked), right words [c, nw) shifted down;
    // idx == 0 / idx == len copy the read through to one side (split.rs:33-42)
    const bool all_right = i == 0, all_left = i == slen && i != 0;
    const unsigned long long c = all_right ? 0 : i / 32;
    const unsigned sh = all_right || all_left ? 0u : 2 * (unsigned)(i % 32);
    if (nw <= kSpPre) {
        const unsigned n = (unsigned)nw, cc = (unsigned)c;   // 32-bit from here on: every `j < nw` on 64 bits is two instructions
        uint64_t x[kSpPre];
#pragma unroll
        for (int j = 0; j < kSpPre; ++j) x[j] = (unsigned)j < n ? __ldg(w + j) : 0ull;
        if (all_left) {
#pragma unroll
            for (int j = 0; j < kSpPre; ++j)
                if ((unsigned)j < n) lo[j] = x[j];
            return;
        }
        if (!all_right) {
#pragma unroll
            for (int j = 0; j < kSpPre; ++j) {
                if ((unsigned)j < cc) lo[j] = x[j];
                else if ((unsigned)j == cc) lo[j] = sh ? x[j] & ((1ull << sh) - 1ull) : 0ull;
            }
        }
#pragma unroll
        for (int j = 0; j < kSpPre; ++j) {
            if ((unsigned)j >= cc && (unsigned)j < n) {
                const uint64_t prev = j > 0 ? x[j > 0 ? j - 1 : 0] : 0ull;
                ro[(unsigned)j - cc] = sh ? (x[j] >> sh) | ((unsigned)j > cc ? prev << (64 - sh) : 0ull) : x[j];
            }
        }
        return;
    }
    if (all_left) {
        for (unsigned long long j = 0; j < nw; ++j) lo[j] = __ldg(w + j);
        return;
    }
    if (!all_right) {
        for (unsigned long long j = 0; j < c; ++j) lo[j] = __ldg(w + j);
        lo[c] = sh ? __ldg(w + c) & ((1ull << sh) - 1ull) : 0ull;
    }
    uint64_t prev = 0;
    for (unsigned long long j = c; j < nw; ++j) {
        const uint64_t cur = __ldg(w + j);
        ro[j - c] = sh ? (cur >> sh) | (j > c ? prev << (64 - sh) : 0ull) : cur;
        prev = cur;
    }
}

constexpr int kSpRows = 8;                            // rows of 32 reads per warp: a CTA tile holds 8 warps x 8 rows x 32 = 2048 reads
constexpr int kSpTile = kThreads * kSpRows;
#ifndef BN_SP_BATCH
#define BN_SP_BATCH 4   // 2: 1.152 ms, 4: 1.115 ms, 8: 1.221 ms (80 registers) -- 1.221 ms row by row (40 M reads)
#endif
constexpr int kSpBatch = BN_SP_BATCH;                 // rows of pass A whose loads are in flight together
constexpr int kSpCap = 256;                           // words of each output span a warp can stage per row (32 reads x 8 words)

// (left, right) word counts of a read; an error takes no room and is reported when `status` is given
__device__ __forceinline__ bool split_shape(unsigned long long r, unsigned long long nw, unsigned long long slen, unsigned long long i,
                                            unsigned& nl, unsigned& nr, unsigned long long* status) {
    nl = nr = 0;
    const bool oob = i > slen;
    if (oob || (i && i < slen && nw && nw < (slen + 31) / 32)) {
        if (status) {
            const unsigned long long key = r << 1 | (oob ? 0ull : 1ull);
            if (key < ld_volatile_u64(status)) atomicMin(status, key);
        }
        return false;
    }
    if (i == 0) nr = (unsigned)nw;
    else if (i == slen) nl = (unsigned)nw;
    else if (nw) nl = (unsigned)(i / 32 + 1), nr = (unsigned)(nw - i / 32);
    return true;
}

// ONE pass.  A CTA takes a tile of 2048 consecutive reads (tiles numbered by a ticket, so look-back never waits for a CTA
// that is not running); warp w owns reads [256 w, 256 w + 256) of the tile as 8 rows of 32 (lane l <-> read 32 i + l: every
// load of the shape arrays and every offset store is a coalesced 256-byte warp transaction).
//   pass A  the warp reads (word_offsets, lens, idx) of its rows and adds up the two word counts;
//           the 8 warp totals are scanned, and warp 0 gets the totals of all earlier tiles by decoupled look-back
//           (lookback.cuh) -- one look-back per 2048 reads (per 256 reads the chain of spinning tiles was slower than
//           the three-launch scan it replaces: 2.39 ms against 1.46 on 40 M reads);
//   pass B  row by row the warp reads the shapes again (L1 / L2 hits: the tile's 48 KiB were fetched microseconds ago),
//           scans the row, writes the offsets, splits one read per lane into a shared-memory image of the row's two
//           output spans -- contiguous in `left` / `right` because the reads are consecutive -- and stores the image with
//           coalesced 8-byte-per-lane stores.  A row whose spans do not fit the image writes straight to global memory.
__global__ void __launch_bounds__(kThreads)
split_packed_fused_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets,
                          const uint64_t* __restrict__ lens, const uint64_t* __restrict__ idx, unsigned long long n_reads,
                          uint64_t* __restrict__ left, uint64_t* __restrict__ left_offsets, uint64_t* __restrict__ right,
                          uint64_t* __restrict__ right_offsets, unsigned long long* __restrict__ status,
                          unsigned long long* __restrict__ lb, unsigned long long n_tiles) {
    __shared__ uint64_t s_left[kWarpsPerBlock][kSpCap], s_right[kWarpsPerBlock][kSpCap];
    __shared__ unsigned long long s_tile, s_warp[2][kWarpsPerBlock], s_base[2];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(lb, 1ull);
    __syncthreads();
    const unsigned long long tile = s_tile;
    const unsigned long long row0 = tile * kSpTile + warp * (32 * kSpRows) + lane;   // this lane's read in row 0
    // ---- pass A: the warp's totals.  The shape loads of kSpBatch rows are issued together, then tested: row by row, the
    // branches of a row's test sat between its loads and the next row's, and the eight rows' round trips ran one after the other
    unsigned long long wl = 0, wr = 0;
#pragma unroll
    for (int i0 = 0; i0 < kSpRows; i0 += kSpBatch) {
        unsigned long long wo[kSpBatch], wn[kSpBatch], sl[kSpBatch], ix[kSpBatch];
#pragma unroll
        for (int j = 0; j < kSpBatch; ++j) {
            const unsigned long long r = row0 + 32 * (i0 + j);
            const bool in = r < n_reads;
            wo[j] = in ? __ldg(word_offsets + r) : 0ull;
            wn[j] = in ? __ldg(word_offsets + r + 1) : 0ull;
            sl[j] = in ? __ldg(lens + r) : 0ull;
            ix[j] = in ? __ldg(idx + r) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kSpBatch; ++j) {
            const unsigned long long r = row0 + 32 * (i0 + j);
            unsigned nl = 0, nr = 0;
            if (r < n_reads) split_shape(r, wn[j] - wo[j], sl[j], ix[j], nl, nr, status);
            wl += nl;
            wr += nr;
        }
    }
    wl = warp_sum_u64(wl);
    wr = warp_sum_u64(wr);
    if (lane == 0) s_warp[0][warp] = wl, s_warp[1][warp] = wr;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long a = lane < kWarpsPerBlock ? s_warp[0][lane] : 0ull, b = lane < kWarpsPerBlock ? s_warp[1][lane] : 0ull;
        unsigned long long ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < kWarpsPerBlock; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= (unsigned)o) ia += x, ib += y;
        }
        if (lane < kWarpsPerBlock) s_warp[0][lane] = ia - a, s_warp[1][lane] = ib - b;   // exclusive prefix of the warp totals
        const unsigned long long agg[2] = {__shfl_sync(0xffffffffu, ia, kWarpsPerBlock - 1), __shfl_sync(0xffffffffu, ib, kWarpsPerBlock - 1)};
        unsigned long long excl[2];
        lookback_exclusive<2>(lb + 1, n_tiles, tile, agg, excl);
        if (lane == 0) s_base[0] = excl[0], s_base[1] = excl[1];
    }
    __syncthreads();
    // ---- pass B: row by row, the warp on its own
    unsigned long long base_l = s_base[0] + s_warp[0][warp], base_r = s_base[1] + s_warp[1][warp];
    // Software pipeline over the rows (a row is two dependent round trips otherwise: shapes, then words): while row i is
    // split, the words of row i + 1 are on their way into L1 (its shapes arrived an iteration ago) and the shapes of row
    // i + 2 are being fetched.
    struct Shape {
        unsigned long long wo, nw, slen, ix;
    };
    auto load_shape = [&](unsigned long long r) {
        Shape sh{0, 0, 0, 0};
        if (r < n_reads) {
            sh.wo = word_offsets[r];
            sh.nw = word_offsets[r + 1] - sh.wo;
            sh.slen = lens[r];
            sh.ix = idx[r];
        }
        return sh;
    };
    Shape cur = load_shape(row0), nxt = load_shape(row0 + 32);
#pragma unroll 1
    for (int i = 0; i < kSpRows; ++i) {
        const unsigned long long r = row0 + 32 * i;
        if (i + 1 < kSpRows && nxt.nw && nxt.nw <= 64) {   // (a hint for sane shapes only: a bogus word count must not form an address)
            prefetch_l1(words + nxt.wo);
            prefetch_l1(words + nxt.wo + nxt.nw - 1);
        }
        const Shape nn = i + 2 < kSpRows ? load_shape(r + 64) : Shape{0, 0, 0, 0};
        const unsigned long long wo = cur.wo, nw = cur.nw, slen = cur.slen, ix = cur.ix;
        cur = nxt;
        nxt = nn;
        unsigned nl = 0, nr = 0;
        bool live = false;
        if (r < n_reads) live = split_shape(r, nw, slen, ix, nl, nr, nullptr);
        unsigned il = nl, ir = nr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, ir, o);
            if (lane >= (unsigned)o) il += a, ir += b;
        }
        const unsigned tot_l = __shfl_sync(0xffffffffu, il, 31), tot_r = __shfl_sync(0xffffffffu, ir, 31);
        const unsigned ll = il - nl, lr = ir - nr;               // offsets inside the row's spans
        if (r < n_reads) {
            left_offsets[r] = base_l + ll;
            right_offsets[r] = base_r + lr;
            if (r + 1 == n_reads) left_offsets[n_reads] = base_l + il, right_offsets[n_reads] = base_r + ir;
        }
        const bool staged = tot_l <= (unsigned)kSpCap && tot_r <= (unsigned)kSpCap;   // warp-uniform
        if (live) {   // two call sites: behind one `staged ? shared : global` pointer every store was a generic ST behind a branch
            if (staged) split_one(words + wo, nw, slen, ix, s_left[warp] + ll, s_right[warp] + lr);
            else split_one(words + wo, nw, slen, ix, left + base_l + ll, right + base_r + lr);
        }
        if (staged) {
            __syncwarp();
            for (unsigned k = lane; k < tot_l; k += 32) left[base_l + k] = s_left[warp][k];
            for (unsigned k = lane; k < tot_r; k += 32) right[base_r + k] = s_right[warp][k];
            __syncwarp();
        }
        base_l += tot_l;
        base_r += tot_r;
    }
}

// ticket + two channels of tile descriptors
size_t split_packed_scratch_bytes(size_t n_reads) { return lookback_bytes(ceil_div(n_reads ? n_reads : 1, kSpTile), 2); }

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
    const unsigned long long n_tiles = ceil_div(n_reads, kSpTile);
    e = cudaMemsetAsync(d_scratch, 0, lookback_bytes(n_tiles, 2), s);
    if (e != cudaSuccess) return e;
    split_packed_fused_kernel<<<(unsigned)n_tiles, kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, d_idx, n_reads, d_left, d_left_offsets,
                                                                    d_right, d_right_offsets, d_status,
                                                                    static_cast<unsigned long long*>(d_scratch), n_tiles);
    return cudaGetLastError();
}

}  // namespace bn
