// fastq.cu -- FASTQ text in HBM -> record offsets -> per-read packed words, without the host ever parsing (sm_100a).
//
// the reference has no parser; its README shows the caller's loop (/root/reference/README.md:160-180,
// `for record in reader { PackedSequence::new(record.seq())? }`).  This file replaces that loop AND the host-side
// construction of the offsets bn_encode_batch wants: the FASTQ bytes are uploaded as they are, the device finds the
// records and encodes every sequence line on a fresh 64-bit word (src/sequence.rs:40-52).
//
// FASTQ as parsed here (strict four-line records): line 4r starts with '@', line 4r+1 is the sequence, line 4r+2
// starts with '+', line 4r+3 (quality) has the length of the sequence; lines end in "\n" or "\r\n"; the last line
// may lack its newline.  A format error wins over an invalid base; among errors of a kind the first in file order.
//
// Device steps (tiles of 16 KiB of text, handed out by the hardware CTA scheduler):
//   1. fastq_lines_kernel    one pass over the text.  A two-op-per-word filter that cannot miss a newline leaves one
//                            vector in five; a ballot makes a bitmap of the survivors and the CTA finishes them densely,
//                            one per thread: exact newline mask, rank by a CTA scan, and every newline leaves a 32-bit
//                            entry in its tile's slot row: position inside the tile, "a '\r' precedes it", "the next
//                            line opens with '@'", "... with '+'"
//      + exclusive scan      -> line index of every tile's first newline, number of lines
//   2. fastq_records_slots_kernel   a warp per tile: record r takes its four entries (walking into the following
//                            tiles where a line crosses a tile boundary) -> seq_offsets[r], seq_lens[r]; header,
//                            separator and quality length checked.  The text is not read again.
//      + exclusive scan      of ceil(len/32) -> word_offsets
//      A tile with more than 2048 lines (average line under 8 bytes) does not fit its slot row: the count pass raises a
//      flag and step 2 runs the dense form instead (fastq_index_kernel re-reads the text and writes nl[line], then
//      fastq_records_kernel) -- both forms are launched, the device-side flag decides which one works.
//   3. the encode, picked by the average record size (launch_fastq_encode):
//        <= 1 KiB   fastq_encode_reads_kernel   a thread per read: only the 32-byte blocks of its sequence line, 256-bit loads
//        >  1 KiB   fastq_encode_long_kernel<G> 8 / 16 / 32 lanes per read walk the sequence line in chunks of 2 G words
//                   (+ fastq_encode_giant_kernel for reads above 2^20 bases: the whole grid on one read)
//      and, reachable through BN_FQ_VARIANT only, the first form:
//      fastq_encode_kernel   PACK THEN CUT over tiles of 48 KiB of text: the whole tile (headers and qualities too) is
//                            packed like the contiguous encode -- aligned coalesced 128-bit loads, 16 bytes -> one
//                            32-bit code, one "contains a non-ACGT byte" flag per vector via a ballot -- into a
//                            shared-memory strip; then one thread per read that starts in the tile cuts the read's
//                            words out of the strip (batch.cu's cut_word).  Phase 1 also leaves a 16-bit map of the
//                            non-ACGT bytes of every vector: a read is valid when the flags of its interior vectors
//                            are clear and the maps of its two partial end vectors are clear under a byte-range mask
//                            (the first version re-read the end vectors from global memory: ncu showed 128-byte
//                            line fills for them, +70 % DRAM traffic).
//                            The one read that runs past the tile is finished straight from global memory.
// HBM traffic: the text is read by the count pass, and its sequence lines (in whole DRAM granules) by the encode.
// Algorithmic bytes: text once + 8 B per word out + 24 B per read of offsets.
#include <cstdlib>

#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

constexpr int kFqTile = 16384;               // bytes of text per line-index tile
constexpr int kFqThreads = 256;              // x 4 vectors of 16 bytes
constexpr int kFqEncTile = 65536;            // bytes of text per encode tile
constexpr int kFqLongWords = 64;             // a read with more words than this inside the strip is cut by whole warps
constexpr unsigned long long kCrBit = 1ull << 63;
constexpr unsigned long long kFqGiantBases = 1ull << 20;   // a read longer than this is cut by the whole grid, not by one warp
constexpr int kFqSlots = 2048;               // line entries per tile (more lines than this: the dense fallback)
constexpr uint32_t kSlotPos = 0x3FFFu, kSlotCr = 1u << 14, kSlotAt = 1u << 15, kSlotPlus = 1u << 16;

// Record format: FASTQ = four lines per record opening with '@'; FASTA (one sequence line per record, the form read
// processors emit) = two lines per record opening with '>'.  shift = log2(lines per record).
struct TextFormat {
    unsigned shift;
    uint32_t header;
};
static inline TextFormat text_format(int fasta) { return fasta ? TextFormat{1u, (uint32_t)'>'} : TextFormat{2u, (uint32_t)'@'}; }

enum { FQ_BAD_HEADER = 1, FQ_BAD_SEPARATOR = 2, FQ_BAD_QUALITY_LENGTH = 3 };   // kinds of format error (4 = truncated: host)

static inline unsigned long long fastq_tiles(size_t n_bytes) { return (unsigned long long)n_bytes / kFqTile + 1; }   // covers position n_bytes

// ---------------------------------------------------------------- text access -------------------------------------
// The text is read as if a '\n' followed a last line that lacks one, and NUL bytes followed that.

static __device__ __noinline__ uint4 fq_load_edge(const uint8_t* __restrict__ bytes, unsigned long long n, bool virt, unsigned long long pos) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (pos > n) return make_uint4(0u, 0u, 0u, 0u);
    for (int j = 0; j < 16; ++j) {
        const unsigned long long i = pos + j;
        uint32_t b = 0;
        if (i < n) b = bytes[i];
        else if (i == n && virt) b = '\n';
        w[j >> 2] |= b << (8 * (j & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 fq_load(const uint8_t* __restrict__ bytes, unsigned long long n, bool virt, unsigned long long pos) {
    return pos + 16 <= n ? ld128<LD_PLAIN>(reinterpret_cast<const uint4*>(bytes + pos)) : fq_load_edge(bytes, n, virt, pos);
}

// 0x80 in every byte of w that equals the pattern byte (exact: no carries cross a byte)
__device__ __forceinline__ uint32_t eq_flags(uint32_t w, uint32_t pat4) {
    const uint32_t x = w ^ pat4;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}
// the four flags (bits 7, 15, 23, 31) gathered into bits 0..3, byte order kept
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return (((f >> 7) * 0x00204081u) >> 21) & 0xFu; }

constexpr uint32_t kNl4 = 0x0A0A0A0Au;

// bit i set iff byte i of the vector is '\n'
__device__ __forceinline__ uint32_t newline_mask16(uint4 v) {
    return flags_to_nibble(eq_flags(v.x, kNl4)) | (flags_to_nibble(eq_flags(v.y, kNl4)) << 4) |
           (flags_to_nibble(eq_flags(v.z, kNl4)) << 8) | (flags_to_nibble(eq_flags(v.w, kNl4)) << 12);
}

__device__ __forceinline__ void report_min(unsigned long long* word, unsigned long long key) {
    if (key < ld_volatile_u64(word)) atomicMin(word, key);
}

// ---------------------------------------------------------------- 1. newlines per tile -----------------------------

// position of the n-th (0-based) set bit of m, n < popc(m): a binary search on popcounts (__fns does the same for either
// direction and any base in ~55 instructions; this is ~25)
__device__ __forceinline__ unsigned nth_set_bit(uint32_t m, unsigned n) {
    unsigned pos = 0, c = __popc(m & 0xFFFFu);
    if (n >= c) n -= c, pos = 16, m >>= 16;
    c = __popc(m & 0xFFu);
    if (n >= c) n -= c, pos += 8, m >>= 8;
    c = __popc(m & 0xFu);
    if (n >= c) n -= c, pos += 4, m >>= 4;
    c = __popc(m & 0x3u);
    if (n >= c) n -= c, pos += 2, m >>= 2;
    return pos + (n >= (m & 1u) ? 1u : 0u);
}

// the four vectors a thread loads of a tile (lane-consecutive: coalesced)
template <int kEdge>   // 0: the tile lies wholly inside the text, 1: it does not
__device__ __forceinline__ void lines_load_tile(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long tile0, unsigned tid,
                                                uint4 (&x)[4]) {
    if (kEdge == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(bytes + tile0);
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ld128<LD_NC_NOALLOC>(src + tid + j * kFqThreads);
    } else {  // the last tile(s): the virtual newline and the NUL padding come from the edge loader
        const bool virt = n && bytes[n - 1] != '\n';
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = fq_load(bytes, n, virt, tile0 + 16ull * (tid + j * kFqThreads));
    }
}

// One tile per CTA, handed out by the hardware scheduler.  (Persistent CTAs that prefetch the next tile's vectors into
// registers while working on the current one were measured: 54 registers, 4 CTAs per SM, 16 % slower.)
//
// FILTER, THEN FINISH DENSELY.  An exact per-byte compare costs ~8 ALU ops per 32-bit word, and ncu showed this kernel
// instruction-bound.  So every vector first takes a 2-op-per-word test that cannot miss a newline: adding 0x60 to every
// byte sets bit 7 of every byte >= 0x20 (no carry leaves a byte of ASCII text; a carry from a byte >= 0xA0 only adds 1 to
// its neighbour, and 0x0A + 0x60 + 1 still has bit 7 clear), so a vector whose four sums AND to 0x80 in every byte holds
// no '\n'.  In FASTQ text one vector in five survives.  A ballot turns the survivors into a bitmap in file order, and the
// CTA then works on the survivors only, one per thread: exact newline mask, rank by a CTA scan, the bytes next to each
// newline from the shared-memory copy of the tile, one slot entry per newline.
// row = the tile's slot row; returns the tile's number of newlines
template <int kEdge>
__device__ __forceinline__ unsigned lines_filter_tile(const uint8_t* __restrict__ bytes, unsigned long long n, uint32_t* __restrict__ row,
                                                      unsigned long long tile, uint32_t header) {
    __shared__ uint4 raw[kFqTile / 16];                   // the tile's text
    __shared__ uint32_t hitbits[kFqTile / 512];           // one bit per vector: may hold a newline
    __shared__ unsigned warp_tot[kFqThreads / 32];
    static_assert(kFqTile / 512 == 32, "one bitmap word per lane");
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long tile0 = tile * kFqTile;
    uint4 x[4];
    lines_load_tile<kEdge>(bytes, n, tile0, tid, x);
    constexpr uint32_t kAdd = 0x60606060u, kTop = 0x80808080u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned v = tid + j * kFqThreads;
        const uint32_t t = (x[j].x + kAdd) & (x[j].y + kAdd) & (x[j].z + kAdd) & (x[j].w + kAdd);
        raw[v] = x[j];   // (staging only the survivors and fetching the bytes at their ends from global memory: 25 % slower)
        const unsigned b = __ballot_sync(0xffffffffu, (t & kTop) != kTop);
        if (lane == 0) hitbits[v >> 5] = b;
    }
    __syncthreads();
    // every warp: inclusive prefix of the 32 bitmap words' popcounts
    const uint32_t hw = hitbits[lane];
    unsigned inc = __popc(hw);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    const unsigned n_hit = __shfl_sync(0xffffffffu, inc, 31);
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(raw);
    unsigned total = 0;                                    // newlines of the hit vectors handled so far (CTA-uniform)
    for (unsigned i0 = 0; i0 < n_hit; i0 += kFqThreads) {
        const unsigned i = i0 + tid;
        const bool active = i < n_hit;
        unsigned v = 0, cnt = 0, c_inc = 0;
        uint32_t m = 0;
        if (i0 + 32u * warp < n_hit) {   // warp-uniform: a warp with no survivor of this round only joins the barriers
            const unsigned ii = active ? i : n_hit - 1;
            // hit vector #ii: the bitmap word k that holds it (= number of words whose inclusive prefix is <= ii), then its bit
            unsigned k = 0;
#pragma unroll
            for (unsigned step = 16; step; step >>= 1) {
                const unsigned t = __shfl_sync(0xffffffffu, inc, k + step - 1);
                if (t <= ii) k += step;
            }
            const unsigned prev_inc = __shfl_sync(0xffffffffu, inc, k ? k - 1 : 0);
            const uint32_t word = __shfl_sync(0xffffffffu, hw, k);
            v = 32u * k + nth_set_bit(word, ii - (k ? prev_inc : 0u));
            m = active ? newline_mask16(raw[v]) : 0u;   // exact; bit b <-> byte b of the vector
            cnt = __popc(m);
            c_inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, c_inc, o);
                if (lane >= (unsigned)o) c_inc += t;
            }
        }
        if (lane == 31) warp_tot[warp] = c_inc;
        __syncthreads();
        // this warp's base and the round's total from the eight warp totals: two warp reductions (REDUX), no loop
        const unsigned wt = lane < kFqThreads / 32 ? warp_tot[lane] : 0u;
        const unsigned round_total = __reduce_add_sync(0xffffffffu, wt);
        unsigned rank = total + c_inc - cnt + __reduce_add_sync(0xffffffffu, lane < warp ? wt : 0u);
        while (m) {
            const unsigned q = 16u * v + (__ffs((int)m) - 1);
            m &= m - 1;
            uint32_t prev, next;
            if (q > 0) prev = rb[q - 1];
            else prev = tile0 ? bytes[tile0 - 1] : 0u;
            if (q + 1 < (unsigned)kFqTile) next = rb[q + 1];        // past the text the staged bytes are the virtual newline / NUL
            else next = tile0 + q + 1 < n ? bytes[tile0 + q + 1] : 0u;
            if (rank < (unsigned)kFqSlots)
                row[rank] = q | (prev == '\r' ? kSlotCr : 0u) | (next == header ? kSlotAt : 0u) | (next == '+' ? kSlotPlus : 0u);
            ++rank;
        }
        total += round_total;
        if (i0 + kFqThreads < n_hit) __syncthreads();   // warp_tot is reused by the next round (CTA-uniform: usually there is none)
    }
    return total;
}

// One tile per CTA, NO loop over tiles in the kernel, and the edge loader in an instantiation of its own: the grid-stride
// loop that used to wrap this body and the rarely taken call of the edge loader cost 6 registers (37 against 31: 6
// against 8 CTAs per SM), and this latency-bound kernel pays for residency -- 1.33 -> 1.16 ms on 20 M x 150 bp FASTQ,
// 1.17 -> 1.00 ms on 10 kbp reads, 0.70 -> 0.63 ms on FASTA (profiles/README.md).
// (A BLOCKED form -- a thread owns 64 or 128 consecutive bytes as 256-bit loads, exact mask in registers, one CTA scan, a walk
// over its own set bits -- was measured beside it: 0.92 ms with 128 bytes per thread and 1.02 ms with 64 on 10 kbp reads,
// 1.48 / 1.20 ms on 150 bp reads; the choice would have to be made on the device from a probe of the text -- two kernels
// enqueued over the same tiles cost 0.2 ms for the loser's 400 000 empty CTAs, two bodies in one kernel 38 registers.  Dropped.)
// kEdge = false: the tiles wholly inside the text; true: the last tile(s) -- virtual newline, NUL padding (the edge loader)
template <bool kEdge>
__global__ void __launch_bounds__(kFqThreads, kEdge ? 1 : 8)
fastq_lines_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long* __restrict__ counts,
                   uint32_t* __restrict__ slots, unsigned* __restrict__ overflow, unsigned long long first_tile, uint32_t header) {
    const unsigned long long tile = first_tile + blockIdx.x;
    const unsigned total = lines_filter_tile<kEdge ? 1 : 0>(bytes, n, slots + tile * kFqSlots, tile, header);
    if (threadIdx.x == 0) {
        counts[tile] = total;
        if (total > (unsigned)kFqSlots) *overflow = 1u;
    }
}

struct CountOfTile {
    const unsigned long long* counts;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long t) const { return counts[t]; }
};

// ---------------------------------------------------------------- 2. line index ------------------------------------

__global__ void __launch_bounds__(kFqThreads)
fastq_index_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, const uint64_t* __restrict__ line_base,
                   unsigned long long n_reads, uint64_t* __restrict__ nl, unsigned long long* __restrict__ status,
                   const unsigned* __restrict__ overflow, unsigned long long n_tiles, TextFormat fmt) {
    __shared__ __align__(8) uint16_t nlb[kFqTile / 16];
    __shared__ unsigned warp_tot[kFqThreads / 32];
    if (*overflow == 0u) return;   // the slot rows hold every line: fastq_records_slots_kernel does the work
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool virt = n && bytes[n - 1] != '\n';
    // a modest grid strides over the tiles: launched after every count pass, it must cost nothing when it has no work
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long tile0 = tile * kFqTile;
    uint4 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = fq_load(bytes, n, virt, tile0 + 16ull * (tid + j * kFqThreads));
#pragma unroll
    for (int j = 0; j < 4; ++j) nlb[tid + j * kFqThreads] = (uint16_t)newline_mask16(x[j]);
    __syncthreads();
    // thread t owns bytes [64 t, 64 t + 64) of the tile: four consecutive vectors
    const uint2 mm = *reinterpret_cast<const uint2*>(nlb + 4 * tid);
    unsigned long long m = ((unsigned long long)mm.y << 32) | mm.x;
    const unsigned cnt = __popcll(m);
    unsigned inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    unsigned rank = inc - cnt;
    for (unsigned w = 0; w < warp; ++w) rank += warp_tot[w];
    const unsigned long long n_whole = n_reads << fmt.shift;     // nl[] holds the lines of whole records
    const unsigned last_kind = (1u << fmt.shift) - 1u;
    const unsigned long long n_lines = line_base[n_tiles];     // all lines, a trailing partial record included
    if (tile == 0 && tid == 0 && n_lines && bytes[0] != fmt.header) report_min(status + 1, FQ_BAD_HEADER);
    unsigned long long L = line_base[tile] + rank;
    while (m) {
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        const unsigned long long p = tile0 + 64ull * tid + b;        // <= n (n itself only for the virtual newline)
        if (L < n_whole) {
            const bool cr = p > 0 && bytes[p - 1] == '\r';
            nl[L] = p | (cr ? kCrBit : 0ull);
        }
        const unsigned kind = (unsigned)L & last_kind;
        const bool ends_record = kind == last_kind, ends_sequence = fmt.shift == 2 && kind == 1;
        if (L + 1 < n_lines && (ends_record || ends_sequence)) {
            // a record's last line ends here: the next record must open with the header character; a FASTQ sequence line:
            // the separator opens with '+'
            const uint32_t c = p + 1 < n ? bytes[p + 1] : '\n';
            if (c != (ends_record ? fmt.header : (uint32_t)'+'))
                report_min(status + 1, ends_record ? (((L + 1) >> fmt.shift) << 8) | FQ_BAD_HEADER : ((L >> fmt.shift) << 8) | FQ_BAD_SEPARATOR);
        }
        ++L;
    }
    __syncthreads();   // nlb / warp_tot are reused by the next tile
    }
}

__global__ void __launch_bounds__(kThreads)
fastq_records_kernel(const uint64_t* __restrict__ nl, unsigned long long n_reads, uint64_t* __restrict__ seq_off,
                     uint64_t* __restrict__ seq_len, unsigned long long* __restrict__ status, const unsigned* __restrict__ overflow,
                     unsigned shift) {
    if (*overflow == 0u) return;
    for (unsigned long long r = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; r < n_reads; r += (unsigned long long)gridDim.x * kThreads) {
    const ulonglong2 ab = reinterpret_cast<const ulonglong2*>(nl)[r << (shift - 1)];
    const unsigned long long s = (ab.x & ~kCrBit) + 1, e = (ab.y & ~kCrBit) - (ab.y >> 63);
    if (shift == 2) {   // FASTQ: the quality line is as long as the sequence
        const ulonglong2 cd = reinterpret_cast<const ulonglong2*>(nl)[2 * r + 1];
        const unsigned long long qs = (cd.x & ~kCrBit) + 1, qe = (cd.y & ~kCrBit) - (cd.y >> 63);
        if (qe - qs != e - s) report_min(status + 1, (r << 8) | FQ_BAD_QUALITY_LENGTH);
    }
    seq_off[r] = s;
    seq_len[r] = e - s;
    }
}

// records from the slot rows, a warp per tile: the records whose header line ends in the tile
constexpr int kFqRecWarps = 8;
template <unsigned kShift>   // log2(lines per record): compile-time, so the per-record arrays stay in registers
__global__ void __launch_bounds__(32 * kFqRecWarps, 8)   // 32 registers: 8 CTAs per SM (index step 0.358 -> 0.335 ms)
fastq_records_slots_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ line_base, const uint32_t* __restrict__ slots,
                           unsigned long long n_tiles, unsigned long long n_reads, uint64_t* __restrict__ seq_off,
                           uint64_t* __restrict__ seq_len, unsigned long long* __restrict__ status, const unsigned* __restrict__ overflow,
                           TextFormat fmt) {
    const unsigned long long t = (unsigned long long)blockIdx.x * kFqRecWarps + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned shift = kShift, lpr = 1u << kShift;              // lines per record
    // the head of the tile's slot row (1 KiB = 256 entries; FASTQ text of 150-bp reads has ~200 lines per tile) is asked for
    // now, beside the line bases its addresses would otherwise wait for
    if (lane < 8) prefetch_l1(slots + t * kFqSlots + 32u * lane);
    const unsigned long long n_lines = line_base[n_tiles];
    const unsigned long long lb = line_base[t], le = line_base[t + 1];
    // (lb <= le always: the second condition only makes the exit depend on the line bases, so that their loads leave
    // together with the flag's instead of after it has arrived)
    if (*overflow != 0u || lb > le) return;
    const unsigned long long n_records = (n_lines + lpr - 1) >> shift;  // a trailing partial record included (its faults count)
    unsigned long long ra = (lb + lpr - 1) >> shift, rb = (le + lpr - 1) >> shift;
    rb = rb < n_records ? rb : n_records;
    if (t == 0 && lane == 0 && n_lines && bytes[0] != fmt.header) report_min(status + 1, FQ_BAD_HEADER);
    for (unsigned long long r = ra + lane; r < rb; r += 32) {
        unsigned long long pos[4] = {0, 0, 0, 0};
        uint32_t ent[4] = {0, 0, 0, 0};
        const unsigned long long l0 = r << shift;                       // the record's header line
        if (l0 + lpr - 1 < le) {   // the usual case: all its lines end in this tile -- neighbouring entries, 32-bit arithmetic
            const uint32_t* e4 = slots + t * kFqSlots + (unsigned)(l0 - lb);
            const unsigned long long base = t * kFqTile;
#pragma unroll
            for (int k = 0; k < (int)lpr; ++k) {
                ent[k] = e4[k];
                pos[k] = base + (ent[k] & kSlotPos);
            }
        } else {
            unsigned long long tt = t;
#pragma unroll
            for (int k = 0; k < (int)lpr; ++k) {
                const unsigned long long L = l0 + k;
                if (L >= n_lines) break;
                while (L >= line_base[tt + 1]) ++tt;                      // a line that ends in a later tile
                ent[k] = slots[tt * kFqSlots + (L - line_base[tt])];
                pos[k] = tt * kFqTile + (ent[k] & kSlotPos);
            }
        }
        if (shift == 2 && l0 + 2 < n_lines && !(ent[1] & kSlotPlus)) report_min(status + 1, (r << 8) | FQ_BAD_SEPARATOR);
        if (l0 + lpr < n_lines && !(ent[lpr - 1] & kSlotAt)) report_min(status + 1, ((r + 1) << 8) | FQ_BAD_HEADER);
        if (r < n_reads) {
            const unsigned long long s = pos[0] + 1, e = pos[1] - ((ent[1] & kSlotCr) ? 1 : 0);
            if (shift == 2) {   // FASTQ: the quality line is as long as the sequence
                const unsigned long long qs = pos[2] + 1, qe = pos[3] - ((ent[3] & kSlotCr) ? 1 : 0);
                if (qe - qs != e - s) report_min(status + 1, (r << 8) | FQ_BAD_QUALITY_LENGTH);
            }
            seq_off[r] = s;
            seq_len[r] = e - s;
        }
    }
}

struct WordsOfLen {
    const uint64_t* lens;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long r) const { return (lens[r] + 31) / 32; }
};

// ---------------------------------------------------------------- 3. encode ----------------------------------------

// a vector at the end of the text, fetched byte-wise ('A' beyond it)
static __device__ __noinline__ uint4 enc_load_edge(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long pos) {
    uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
    for (int j = 0; j < 16; ++j)
        if (pos + j < n) w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)bytes[pos + j] << (8 * (j & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 enc_load_cached(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long pos) {
    return pos + 16 <= n ? ld128<LD_PLAIN>(reinterpret_cast<const uint4*>(bytes + pos)) : enc_load_edge(bytes, n, pos);
}

// bytes l..h-1 of a 32-bit word (clamped to the word)
__device__ __forceinline__ uint32_t byte_range_mask(int l, int h) {
    l = l < 0 ? 0 : l;
    h = h > 4 ? 4 : h;
    if (l >= h) return 0u;
    return (uint32_t)(((1ull << (8 * h)) - 1ull) & ~((1ull << (8 * l)) - 1ull));
}
// does the vector hold a byte outside ACGTacgt at a byte index in [l, h)?
__device__ __forceinline__ bool bad_in_range(uint4 v, int l, int h) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t s1 = w[i] >> 1, s2 = w[i] >> 2;
        const uint32_t tcol = s2 & ~s1 & 0x01010101u;
        const uint32_t expect = tcol * 0x11u + 0x41414141u;
        acc |= (w[i] ^ expect) & kValidMask & byte_range_mask(l - 4 * i, h - 4 * i);
    }
    return acc != 0;
}

// rare path, out of line: first invalid byte of text[lo, hi) -> status word
static __device__ __noinline__ void fq_report_range(const uint8_t* __restrict__ bytes, unsigned long long lo, unsigned long long hi,
                                                    unsigned long long* status) {
    if ((ld_volatile_u64(status) >> 8) <= lo) return;   // an earlier invalid base is already known
    for (unsigned long long i = lo; i < hi; ++i) {
        const uint32_t b = bytes[i];
        if (!byte_is_valid(b)) {
            report_invalid(status, i, b);
            return;
        }
    }
}

// 16 ASCII bytes -> 32 bits of packed codes (as pack16), and a 16-bit map of the bytes outside ACGTacgt:
// bit 4k + i <-> byte k of word i (byte 4i + k of the vector) -- the order the word-parallel test produces.
__device__ __forceinline__ uint32_t pack4_top_map(uint32_t w, uint32_t& nz) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t code = (s1 ^ s2) & 0x03030303u;
    const uint32_t tcol = s2 & ~s1 & 0x01010101u;
    const uint32_t d = w ^ (tcol * 0x11u + 0x41414141u);    // under kValidMask: zero exactly in the valid bytes
    const uint32_t t = (d & 0x59595959u) + 0x7F7F7F7Fu;     // bit 7 <- one of the checked low bits differs (no carry leaves a byte)
    nz = (t | d) & 0x80808080u;                             // ... or bit 7 itself does
    return code * 0x01041040u;
}
__device__ __forceinline__ uint32_t pack16_map(uint4 v, uint32_t& map16) {
    uint32_t n0, n1, n2, n3;
    const uint32_t p0 = pack4_top_map(v.x, n0), p1 = pack4_top_map(v.y, n1);
    const uint32_t p2 = pack4_top_map(v.z, n2), p3 = pack4_top_map(v.w, n3);
    const uint32_t vb = (n0 >> 7) | (n1 >> 6) | (n2 >> 5) | (n3 >> 4);   // bit 8k + i
    map16 = __byte_perm(vb | (vb >> 4), 0u, 0x4420);                      // bit 4k + i
    const uint32_t lo = __byte_perm(p0, p1, 0x7373), hi = __byte_perm(p2, p3, 0x7373);
    return __byte_perm(lo, hi, 0x5410);
}
// the map bits of the vector's bytes 0 .. x-1 (x <= 16): whole words i < x/4, and bytes k < x%4 of word x/4
__device__ __forceinline__ uint32_t map_below(unsigned x) {
    const unsigned i = x >> 2, k = x & 3u;
    return (0x1111u * ((1u << i) - 1u)) | ((0x1111u & ((1u << (4u * k)) - 1u)) << i);
}

// Is there an invalid byte in tile bytes [lo, hi), hi > lo?  Whole vectors by their flags, the partial ones by their maps.
__device__ __forceinline__ bool strip_range_invalid(const uint32_t* __restrict__ flags, const uint16_t* __restrict__ maps, unsigned lo,
                                                    unsigned hi) {
    const unsigned v_lo = (lo + 15u) >> 4, v_hi = hi >> 4;
    uint32_t bad = 0;
    if (v_lo <= v_hi) {
        if (lo & 15u) bad |= maps[v_lo - 1] & ~map_below(lo & 15u);
        if (hi & 15u) bad |= maps[v_hi] & map_below(hi & 15u);
        if (v_lo < v_hi) {
            const unsigned w_lo = v_lo >> 5, w_hi = (v_hi - 1u) >> 5;
            for (unsigned w = w_lo; w <= w_hi; ++w) {
                uint32_t f = flags[w];
                if (w == w_lo) f &= 0xFFFFFFFFu << (v_lo & 31u);
                if (w == w_hi) f &= 0xFFFFFFFFu >> (31u - ((v_hi - 1u) & 31u));
                bad |= f;
            }
        }
    } else {  // both ends inside one vector
        bad = maps[lo >> 4] & map_below(hi & 15u) & ~map_below(lo & 15u);
    }
    return bad != 0u;
}

// The same test without maps: the two partial end vectors are fetched again, as single 32-byte sectors
// (ld.global.nc.L1::no_allocate -- a plain cached load was served as a 128-byte line fill).
__device__ __forceinline__ uint4 enc_load_stream(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long pos) {
    return pos + 16 <= n ? ld128<LD_NC_NOALLOC>(reinterpret_cast<const uint4*>(bytes + pos)) : enc_load_edge(bytes, n, pos);
}
__device__ __forceinline__ bool strip_range_invalid_reload(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long tile0,
                                                           const uint32_t* __restrict__ flags, unsigned lo, unsigned hi) {
    const unsigned v_lo = (lo + 15u) >> 4, v_hi = hi >> 4;
    bool bad = false;
    if (v_lo <= v_hi) {
        if (lo & 15u) bad |= bad_in_range(enc_load_stream(bytes, n, tile0 + 16ull * (v_lo - 1)), (int)(lo & 15u), 16);
        if (hi & 15u) bad |= bad_in_range(enc_load_stream(bytes, n, tile0 + 16ull * v_hi), 0, (int)(hi & 15u));
        if (v_lo < v_hi) {
            const unsigned w_lo = v_lo >> 5, w_hi = (v_hi - 1u) >> 5;
            for (unsigned w = w_lo; w <= w_hi; ++w) {
                uint32_t f = flags[w];
                if (w == w_lo) f &= 0xFFFFFFFFu << (v_lo & 31u);
                if (w == w_hi) f &= 0xFFFFFFFFu >> (31u - ((v_hi - 1u) & 31u));
                bad |= f != 0u;
            }
        }
    } else {
        bad = bad_in_range(enc_load_stream(bytes, n, tile0 + 16ull * (lo >> 4)), (int)(lo & 15u), (int)(hi & 15u));
    }
    return bad;
}

// the output word whose first base sits `rel` bytes into the strip (same window as batch.cu)
__device__ __forceinline__ uint64_t fq_cut_word(const uint32_t* __restrict__ codes, unsigned rel) {
    const unsigned vi = rel >> 4, sh = 2u * (rel & 15u);
    const uint32_t c0 = codes[vi], c1 = codes[vi + 1], c2 = codes[vi + 2];
    return ((uint64_t)__funnelshift_r(c1, c2, sh) << 32) | __funnelshift_r(c0, c1, sh);
}

// one output word straight from global memory: nb (1..32) bases starting at text offset a
__device__ __forceinline__ uint64_t fq_direct_word(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long a, unsigned nb,
                                                   bool& bad) {
    const unsigned long long a0 = a & ~15ull;
    const unsigned off = (unsigned)(a & 15ull);
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = 0;
        if (16u * k < off + nb) {
            const uint4 x = enc_load_cached(bytes, n, a0 + 16ull * k);
            uint32_t ignore = 0;
            c[k] = pack16(x, ignore);
            bad |= bad_in_range(x, (int)off - 16 * k, (int)(off + nb) - 16 * k);
        }
    }
    const unsigned sh = 2u * off;
    uint64_t w = ((uint64_t)__funnelshift_r(c[1], c[2], sh) << 32) | __funnelshift_r(c[0], c[1], sh);
    if (nb < 32) w &= (1ull << (2 * nb)) - 1ull;
    return w;
}

struct FqLongSeg {
    unsigned long long first;  // its first output word
    unsigned rel;              // byte offset of its first base inside the strip
    unsigned count;            // words
    unsigned tail;             // bases in its last word
    unsigned pad;
};

template <int kTile, int kBThreads, int kMinCtas, bool kMaps = true>
__global__ void __launch_bounds__(kBThreads, kMinCtas)
fastq_encode_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, const uint64_t* __restrict__ line_base,
                    unsigned long long n_tiles1, unsigned long long n_reads, const uint64_t* __restrict__ seq_off,
                    const uint64_t* __restrict__ seq_len, const uint64_t* __restrict__ word_off, uint64_t* __restrict__ out,
                    unsigned long long* __restrict__ status, unsigned shift) {
    constexpr int kMainVecs = kTile / 16;
    constexpr int kOver = 3;                                  // a word that starts in the tile ends at most 47 bytes past it
    constexpr int kRatio = kTile / kFqTile;
    constexpr int kLongCap = kTile / (32 * kFqLongWords) + 2;
    constexpr int kWarps = kBThreads / 32;
    static_assert(kMainVecs % (4 * kBThreads) == 0, "tile must be a whole number of load rounds");
    __shared__ uint32_t codes[kMainVecs + 8];
    __shared__ uint16_t maps[kMaps ? kMainVecs + 8 : 1];      // per vector: which bytes are outside ACGTacgt
    __shared__ uint32_t flags[kMainVecs / 32 + 1];            // per vector: any
    __shared__ FqLongSeg segs[kLongCap];
    __shared__ uint16_t chunks[kMainVecs / 32 + kLongCap + 2];
    __shared__ unsigned n_segs, n_chunks, spill_j0;
    __shared__ unsigned long long spill_r;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long tile0 = (unsigned long long)blockIdx.x * kTile;
    // reads whose header line ends in this tile (so their sequence starts in it, or on the first byte after it)
    const unsigned long long t1a = (unsigned long long)blockIdx.x * kRatio;
    const unsigned long long t1b = t1a + kRatio < n_tiles1 ? t1a + kRatio : n_tiles1;
    const unsigned long long round_up = (1ull << shift) - 1ull;   // records whose header line ends in the tile
    unsigned long long ra = (line_base[t1a] + round_up) >> shift, rb = (line_base[t1b] + round_up) >> shift;
    ra = ra < n_reads ? ra : n_reads;
    rb = rb < n_reads ? rb : n_reads;
    if (ra >= rb) return;   // nothing starts here (e.g. the inside of a long read): the tile is not even loaded
    if (tid == 0) {
        n_segs = 0;
        n_chunks = 0;
        spill_j0 = 0xFFFFFFFFu;
    }
    // ---- phase 1: pack the tile's text, 16 bytes per thread step, into the code strip; one validity flag per vector
    for (unsigned vb = 0; vb < (unsigned)kMainVecs; vb += 4 * kBThreads) {
        uint4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned long long pos = tile0 + 16ull * (vb + j * kBThreads + tid);
            x[j] = pos + 16 <= n ? ld128<LD_NC_NOALLOC>(reinterpret_cast<const uint4*>(bytes + pos)) : enc_load_edge(bytes, n, pos);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned v = vb + j * kBThreads + tid;
            uint32_t map16;
            if constexpr (kMaps) {
                codes[v] = pack16_map(x[j], map16);
                maps[v] = (uint16_t)map16;
            } else {
                map16 = 0;
                codes[v] = pack16(x[j], map16);
                map16 &= kValidMask;
            }
            const unsigned fb = __ballot_sync(0xffffffffu, map16 != 0u);
            if (lane == 0) flags[v >> 5] = fb;
        }
    }
    if (warp == 0) {   // the overhang (and the slack words cut_word may touch)
        const unsigned v = kMainVecs + lane;
        uint32_t map16 = 0, c = 0;
        if (lane < kOver) c = pack16_map(enc_load_cached(bytes, n, tile0 + 16ull * v), map16);
        if (lane < 8) {
            codes[v] = c;
            if constexpr (kMaps) maps[v] = (uint16_t)map16;
        }
        const unsigned fb = __ballot_sync(0xffffffffu, map16 != 0u);
        if (lane == 0) flags[kMainVecs >> 5] = fb;
    }
    __syncthreads();
    // ---- phase 2a: one read per thread
    for (unsigned long long r = ra + tid; r < rb; r += kBThreads) {
        const unsigned long long s = __ldg(seq_off + r), len = __ldg(seq_len + r), wo = __ldg(word_off + r);
        if (len == 0) continue;
        const unsigned long long nw = (len + 31) >> 5;
        const unsigned rel = (unsigned)(s - tile0);                              // 1 .. kTile
        const unsigned cap = (kTile + 16u - rel + 31u) >> 5;                     // words whose first base is < kTile + 16
        const unsigned n_in = nw < cap ? (unsigned)nw : cap;
        const unsigned bases_in = len < 32ull * n_in ? (unsigned)len : 32u * n_in;
        bool invalid;
        if constexpr (kMaps) invalid = strip_range_invalid(flags, maps, rel, rel + bases_in);
        else invalid = strip_range_invalid_reload(bytes, n, tile0, flags, rel, rel + bases_in);
        if (invalid) fq_report_range(bytes, s, s + len, status);
        if (n_in < nw) {   // at most one read runs past the strip
            spill_r = r;
            spill_j0 = n_in;
        }
        const unsigned tail = n_in == nw ? (unsigned)(len - 32ull * (nw - 1)) : 32u;
        if (n_in > (unsigned)kFqLongWords) {
            const unsigned slot = atomicAdd(&n_segs, 1u);
            const unsigned nch = (n_in + 31u) / 32u;
            const unsigned cb = atomicAdd(&n_chunks, nch);
            segs[slot] = FqLongSeg{wo, rel, n_in, tail, 0u};
            for (unsigned c = 0; c < nch; ++c) chunks[cb + c] = (uint16_t)(slot << 8 | c);
            continue;
        }
        uint64_t* o = out + wo;
        unsigned q = rel;
        for (unsigned j = 0; j + 1 < n_in; ++j, q += 32) o[j] = fq_cut_word(codes, q);
        uint64_t w = fq_cut_word(codes, q);
        if (tail < 32) w &= (1ull << (2 * tail)) - 1ull;
        o[n_in - 1] = w;
    }
    __syncthreads();
    // ---- phase 2b: long in-strip segments, 32 consecutive words per warp step
    const unsigned nc = n_chunks;
    for (unsigned k = warp; k < nc; k += kWarps) {
        const unsigned ch = chunks[k];
        const FqLongSeg sg = segs[ch >> 8];
        const unsigned j = (ch & 0xFFu) * 32u + lane;
        if (j < sg.count) {
            uint64_t w = fq_cut_word(codes, sg.rel + 32u * j);
            if (j + 1 == sg.count && sg.tail < 32) w &= (1ull << (2 * sg.tail)) - 1ull;
            out[sg.first + j] = w;
        }
    }
    // ---- phase 2c: the rest of the read that runs past the strip, straight from global memory
    if (spill_j0 != 0xFFFFFFFFu) {
        const unsigned long long r = spill_r;
        const unsigned long long s = __ldg(seq_off + r), len = __ldg(seq_len + r), wo = __ldg(word_off + r);
        const unsigned long long nw = (len + 31) >> 5;
        for (unsigned long long j = spill_j0 + tid; j < nw; j += kBThreads) {
            const unsigned long long a = s + 32ull * j;
            const unsigned nb = len - 32ull * j < 32ull ? (unsigned)(len - 32ull * j) : 32u;
            bool bad = false;
            out[wo + j] = fq_direct_word(bytes, n, a, nb, bad);
            if (bad) fq_report_range(bytes, a, a + nb, status);
        }
    }
}

// ---------------------------------------------------------------- 3'. encode, a thread per read (short reads) --------
// In FASTQ text of short reads less than half the bytes are sequence (150 of ~330 per record): the pack-then-cut kernel
// above loads and packs all of them.  Here a thread takes ONE read and fetches only the 32-byte blocks its sequence line
// touches, as 256-bit loads (LDG.256, sm_100): a block is exactly one DRAM sector, so nothing is fetched twice and nothing
// relies on L1; three blocks are in flight per thread before the first is packed (a 150-base read is five or six).  A block packs to one
// 64-bit code word; output word j is the 64-bit window 2 * (start mod 32) bits into code words j, j + 1.  Bytes of the
// first and last vector that lie outside the line are replaced by 'A' before packing (valid, code 00: they shift out at
// the front and are the zero padding at the back).  No shared memory, no barrier.  Records of 1-4 KiB keep the tiled kernel,
// longer ones take fastq_encode_long_kernel below.

// bytes [l, h) of the vector stay, the others become 'A'
__device__ __forceinline__ uint4 keep_bytes(uint4 v, int l, int h) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t m = byte_range_mask(l - 4 * i, h - 4 * i);
        w[i] = (w[i] & m) | (0x41414141u & ~m);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// one read: text[s, s + len), len > 0 -> o[0 .. ceil(len / 32)); kFqReadU blocks in flight
template <int kFqReadU>
__device__ __forceinline__ void fq_encode_read(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long s, unsigned long long len,
                                               uint64_t* __restrict__ o, unsigned long long* __restrict__ status) {
    const unsigned long long e = s + len, a0 = s & ~31ull;
    const unsigned sh2 = 2u * (unsigned)(s & 31ull);
    const unsigned long long nb = (e - a0 + 31) / 32, nw = (len + 31) / 32;   // blocks touched, words written (nb = nw or nw + 1)
    const bool wide = (reinterpret_cast<uintptr_t>(bytes) & 31u) == 0;        // 256-bit loads need the text itself 32-byte aligned
    uint64_t carry = 0;
    uint32_t bad = 0;
    for (unsigned long long b0 = 0; b0 < nb; b0 += kFqReadU) {
        uint4 lo[kFqReadU], hi[kFqReadU];
#pragma unroll
        for (int u = 0; u < kFqReadU; ++u) {
            const unsigned long long pos = a0 + 32 * (b0 + u);
            if (b0 + u >= nb) {
                lo[u] = hi[u] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
            } else if (wide && pos + 32 <= n) {
                const uint8x x = ld256<LD_NC_NOALLOC>(bytes + pos);
                lo[u] = x.lo;
                hi[u] = x.hi;
            } else {
                lo[u] = enc_load_cached(bytes, n, pos);
                hi[u] = enc_load_cached(bytes, n, pos + 16);
            }
        }
#pragma unroll
        for (int u = 0; u < kFqReadU; ++u) {
            const unsigned long long b = b0 + u;
            if (b >= nb) break;
            const unsigned long long pos = a0 + 32 * b;
            uint4 v0 = lo[u], v1 = hi[u];
            if (pos < s || pos + 16 > e) v0 = keep_bytes(v0, pos < s ? (int)(s - pos) : 0, pos + 16 > e ? (e > pos ? (int)(e - pos) : 0) : 16);
            if (pos + 16 < s || pos + 32 > e)
                v1 = keep_bytes(v1, pos + 16 < s ? (int)(s - pos - 16) : 0, pos + 32 > e ? (e > pos + 16 ? (int)(e - pos - 16) : 0) : 16);
            const uint64_t c = ((uint64_t)pack16(v1, bad) << 32) | pack16(v0, bad);
            if (b >= 1) o[b - 1] = sh2 ? (carry >> sh2) | (c << (64 - sh2)) : carry;   // word b - 1 < nw always (b <= nb - 1 <= nw)
            carry = c;
        }
    }
    if (nb == nw) o[nw - 1] = sh2 ? carry >> sh2 : carry;
    if (bad & kValidMask) fq_report_range(bytes, s, e, status);
}

template <int kFqReadU, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
fastq_encode_reads_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long n_reads,
                          const uint64_t* __restrict__ seq_off, const uint64_t* __restrict__ seq_len,
                          const uint64_t* __restrict__ word_off, uint64_t* __restrict__ out, unsigned long long* __restrict__ status) {
    const unsigned long long r = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned long long len = seq_len[r], s = seq_off[r], wo = word_off[r];
    // (the second condition is never true -- two offsets into buffers.  It makes the exit DEPEND on all three table entries:
    // otherwise ptxas sinks the loads of `s` and `wo` below the exit, and their round trip starts when `len` has arrived)
    if (len == 0 || (s & wo) == ~0ull) return;
    fq_encode_read<kFqReadU>(bytes, n, s, len, out + wo, status);
}

// ---------------------------------------------------------------- 3''. encode, a warp per read (long reads) ----------
// In FASTQ text half the bytes are quality values; the tiled kernel above loads and packs them with everything else
// (6.0 GB of text for 3.0 Gbases of 10 kbp reads).  Here a warp takes ONE read and walks its sequence line in chunks of 64
// output words: the (up to 130) aligned 16-byte vectors behind a chunk are fetched lane-consecutively -- all of a lane's
// loads before anything is packed --, packed into a per-warp code strip in shared memory (bytes of the first and last vector
// outside the chunk become 'A': valid, and code 00 is the zero padding of a ragged last word), and every lane cuts two words
// out of the strip: 512-byte coalesced loads, 256-byte coalesced stores, only sequence bytes cross the memory system.
// A read of more than kFqGiant bases is queued instead and cut by the whole grid (fastq_encode_giant_kernel): one warp
// would crawl through a chromosome on one line.
// Records between the two (1 - 4 KiB) take the same walk with G = 16 or 8 lanes per read -- two or four reads per warp, chunks
// of 2 G words, the group synchronising on its own lane mask -- instead of the tiled kernel.
constexpr int kFqChunkPerLane = 5;          // a chunk of 2 G words is 4 G + 2 vectors (+ 2 padding codes): five per lane
#ifndef BN_FQL_WARPS
#define BN_FQL_WARPS 4   // warps per CTA: 2 / 4 / 8 / 16 -> 0.693 / 0.693 / 0.706 / 0.741 ms (10 kbp), 0.989 / 0.992 / 1.045 / 1.163 ms (1 kbp)
#endif
constexpr int kFqLongWarps = BN_FQL_WARPS;
constexpr unsigned long long kFqGiant = kFqGiantBases;
template <int G>
struct FqGroup {
    static constexpr int kWords = 2 * G, kVecs = 4 * G + 2, kStrip = kVecs + 6;
    static_assert(G == 8 || G == 16 || G == 32, "lanes per read");
    static_assert(kFqChunkPerLane * G >= kVecs + 2, "five vectors per lane cover a chunk and its padding codes");
};

// chunks first, first + stride, ... of the read text[s, s + len) -> o[..], by the G lanes of this thread's group;
// strip = FqGroup<G>::kStrip codes owned by the group
template <int G>
__device__ __forceinline__ void fq_encode_chunks(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long s,
                                                 unsigned long long len, uint64_t* __restrict__ o, uint32_t* __restrict__ strip,
                                                 unsigned long long first, unsigned long long stride,
                                                 unsigned long long* __restrict__ status) {
    constexpr int kWords = FqGroup<G>::kWords;
    const unsigned lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));   // the lanes of this group
    const unsigned long long nw = (len + 31) / 32, e = s + len;
    for (unsigned long long w0 = first * kWords; w0 < nw; w0 += stride * kWords) {
        const unsigned cnt = nw - w0 < (unsigned long long)kWords ? (unsigned)(nw - w0) : (unsigned)kWords;
        const unsigned long long p0 = s + 32 * w0, p1 = p0 + 32ull * cnt < e ? p0 + 32ull * cnt : e;   // the chunk's bytes
        const unsigned long long a0 = p0 & ~15ull;
        // everything below in 32-bit offsets from a0 (ncu: the kernel is bound by instruction issue, and positions compared
        // on 64 bits were a good part of its 515 warp instructions per chunk)
        const unsigned rel0 = (unsigned)(p0 - a0), span = (unsigned)(p1 - a0), nvec = (span + 15u) >> 4;   // <= kVecs
        const uint4* src = reinterpret_cast<const uint4*>(bytes + a0);
        uint4 x[kFqChunkPerLane];
        if (a0 + 16ull * nvec <= n) {   // group-uniform: the chunk's vectors lie inside the text
#pragma unroll
            for (int j = 0; j < kFqChunkPerLane; ++j) {
                const unsigned v = gl + G * j;
                x[j] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
                if (v < nvec) x[j] = ld128<LD_NC_NOALLOC>(src + v);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kFqChunkPerLane; ++j) {
                const unsigned v = gl + G * j;
                x[j] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
                if (v < nvec) x[j] = enc_load_stream(bytes, n, a0 + 16ull * v);
            }
        }
        uint32_t bad = 0;
        // An inner chunk -- it starts on a vector boundary or inside the read, and ends inside the read -- is packed as it is:
        // the bytes of its end vectors that lie outside it are bases of the same read (an invalid one among them is found by
        // the neighbouring chunk too; here it costs one needless scan).  The read's first and last chunk mask their end vectors.
        if ((w0 != 0 || rel0 == 0) && p1 != e) {   // group-uniform
#pragma unroll
            for (int j = 0; j < kFqChunkPerLane; ++j) {
                const unsigned v = gl + G * j;
                if (v < nvec) strip[v] = pack16(x[j], bad);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kFqChunkPerLane; ++j) {
                const unsigned v = gl + G * j;
                if (v < nvec + 2) {   // two more codes (zero): the window of a ragged last word reaches past the last vector
                    const unsigned pos = 16u * v;
                    uint4 y = x[j];
                    if (pos < rel0 || pos + 16 > span)
                        y = keep_bytes(y, pos < rel0 ? (int)(rel0 - pos) : 0, pos + 16 > span ? (span > pos ? (int)(span - pos) : 0) : 16);
                    strip[v] = pack16(y, bad);
                }
            }
        }
        __syncwarp(gmask);
        uint64_t* oc = o + w0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const unsigned j = gl + G * k;
            if (j < cnt) oc[j] = fq_cut_word(strip, rel0 + 32u * j);
        }
        if (__any_sync(gmask, (bad & kValidMask) != 0u) && gl == 0) fq_report_range(bytes, p0, p1, status);
        __syncwarp(gmask);   // the strip is reused by the next chunk
    }
}

// G lanes per read
template <int G>
__global__ void __launch_bounds__(32 * kFqLongWarps, 32 / kFqLongWarps)   // 64 registers; G = 32: 3 / 4 / 5 CTAs per SM 0.918 / 0.894 / 1.012 ms
fastq_encode_long_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long n_reads,
                         const uint64_t* __restrict__ seq_off, const uint64_t* __restrict__ seq_len, const uint64_t* __restrict__ word_off,
                         uint64_t* __restrict__ out, unsigned long long* __restrict__ status, unsigned long long* __restrict__ giants) {
    constexpr int kGroups = 32 * kFqLongWarps / G;   // reads per CTA
    __shared__ uint32_t strip[kGroups][FqGroup<G>::kStrip];
    const unsigned group = threadIdx.x / G;
    const unsigned long long r = (unsigned long long)blockIdx.x * kGroups + group;
    if (r >= n_reads) return;
    const unsigned long long len = seq_len[r], s = seq_off[r], wo = word_off[r];
    if (len == 0 || (s & wo) == ~0ull) return;   // (the second term only keeps the three loads together, see fastq_encode_reads_kernel)
    if (len > kFqGiant) {
        if ((threadIdx.x & (G - 1)) == 0) giants[1 + atomicAdd(giants, 1ull)] = r;
        return;
    }
    fq_encode_chunks<G>(bytes, n, s, len, out + wo, strip[group], 0, 1, status);
}

// the queued giant reads, one after the other, the whole grid striding over the chunks of each
__global__ void __launch_bounds__(32 * kFqLongWarps)
fastq_encode_giant_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, const uint64_t* __restrict__ seq_off,
                          const uint64_t* __restrict__ seq_len, const uint64_t* __restrict__ word_off, uint64_t* __restrict__ out,
                          unsigned long long* __restrict__ status, const unsigned long long* __restrict__ giants) {
    __shared__ uint32_t strip[kFqLongWarps][FqGroup<32>::kStrip];
    const unsigned long long n_giants = giants[0];
    const unsigned long long warp = (unsigned long long)blockIdx.x * kFqLongWarps + (threadIdx.x >> 5), n_warps = (unsigned long long)gridDim.x * kFqLongWarps;
    for (unsigned long long i = 0; i < n_giants; ++i) {
        const unsigned long long r = giants[1 + i];
        fq_encode_chunks<32>(bytes, n, seq_off[r], seq_len[r], out + word_off[r], strip[threadIdx.x >> 5], warp, n_warps, status);
    }
}

// ---------------------------------------------------------------- launchers ----------------------------------------
// d_scratch (fastq_scratch_bytes): counts[n_tiles] | line_base[n_tiles + 1] | scan sums
// d_index_scratch (fastq_index_scratch_bytes): nl[4 n_reads] | scan sums

static inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

size_t fastq_scratch_bytes(size_t n_bytes) {
    const size_t t = fastq_tiles(n_bytes);
    // ... + the queue of giant reads of the long-read encode (a count, then at most n_bytes / kFqGiantBases entries)
    return align16(t * 8) + align16((t + 1) * 8) + align16(scan_scratch_bytes(t)) + 16 + t * kFqSlots * sizeof(uint32_t) +
           (2 + n_bytes / kFqGiantBases) * sizeof(unsigned long long);
}
size_t fastq_index_scratch_bytes(size_t n_reads) { return align16((n_reads ? n_reads : 1) * 32) + scan_scratch_bytes(n_reads ? n_reads : 1); }

struct FqScratch {
    unsigned long long* counts;
    uint64_t* line_base;
    unsigned long long* sums;
    unsigned* overflow;
    uint32_t* slots;
    unsigned long long* giants;
    unsigned long long n_tiles;
    FqScratch(void* p, size_t n_bytes) {
        n_tiles = fastq_tiles(n_bytes);
        char* c = static_cast<char*>(p);
        counts = reinterpret_cast<unsigned long long*>(c);
        c += align16(n_tiles * 8);
        line_base = reinterpret_cast<uint64_t*>(c);
        c += align16((n_tiles + 1) * 8);
        sums = reinterpret_cast<unsigned long long*>(c);
        c += align16(scan_scratch_bytes(n_tiles));
        overflow = reinterpret_cast<unsigned*>(c);
        slots = reinterpret_cast<uint32_t*>(c + 16);
        giants = reinterpret_cast<unsigned long long*>(slots + n_tiles * kFqSlots);
    }
};

cudaError_t launch_fastq_count(const DeviceInfo&, const uint8_t* d_bytes, size_t n_bytes, void* d_scratch, uint64_t* d_n_lines,
                               int fasta, cudaStream_t s) {
    const TextFormat fmt = text_format(fasta);
    if (n_bytes == 0) return cudaMemsetAsync(d_n_lines, 0, sizeof(uint64_t), s);
    const FqScratch sc(d_scratch, n_bytes);
    cudaError_t e = cudaMemsetAsync(sc.overflow, 0, 16, s);
    if (e != cudaSuccess) return e;
    // one tile per CTA: a resident grid striding over the tiles was measured at 1x and 2x the resident CTA count and is
    // 10-15 % slower (as for the codec kernels: the two dies finish at different times)
    const unsigned long long n_full = (unsigned long long)n_bytes / kFqTile;   // < n_tiles
    if (n_full) fastq_lines_kernel<false><<<(unsigned)n_full, kFqThreads, 0, s>>>(d_bytes, n_bytes, sc.counts, sc.slots, sc.overflow, 0, fmt.header);
    fastq_lines_kernel<true><<<(unsigned)(sc.n_tiles - n_full), kFqThreads, 0, s>>>(d_bytes, n_bytes, sc.counts, sc.slots, sc.overflow, n_full, fmt.header);
    launch_exclusive_scan(CountOfTile{sc.counts}, sc.n_tiles, sc.sums, sc.line_base, s);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(d_n_lines, sc.line_base + sc.n_tiles, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
}

cudaError_t launch_fastq_index(const DeviceInfo&, const uint8_t* d_bytes, size_t n_bytes, size_t n_reads, void* d_scratch,
                               void* d_index_scratch, uint64_t* d_seq_offsets, uint64_t* d_seq_lens, uint64_t* d_word_offsets,
                               unsigned long long* d_status, int fasta, cudaStream_t s) {
    const TextFormat fmt = text_format(fasta);
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, 2 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_bytes == 0) return cudaMemsetAsync(d_word_offsets, 0, sizeof(uint64_t), s);
    const FqScratch sc(d_scratch, n_bytes);
    uint64_t* nl = static_cast<uint64_t*>(d_index_scratch);
    // both forms are enqueued; the overflow flag left by the count pass lets exactly one of them work.  They run even
    // without a whole record: the faults of a partial one are found here.
    if (fasta)
        fastq_records_slots_kernel<1><<<(unsigned)ceil_div(sc.n_tiles, kFqRecWarps), 32 * kFqRecWarps, 0, s>>>(
            d_bytes, sc.line_base, sc.slots, sc.n_tiles, n_reads, d_seq_offsets, d_seq_lens, d_status, sc.overflow, fmt);
    else
        fastq_records_slots_kernel<2><<<(unsigned)ceil_div(sc.n_tiles, kFqRecWarps), 32 * kFqRecWarps, 0, s>>>(
            d_bytes, sc.line_base, sc.slots, sc.n_tiles, n_reads, d_seq_offsets, d_seq_lens, d_status, sc.overflow, fmt);
    const unsigned long long dense_grid = sc.n_tiles < 148ull * 8 ? sc.n_tiles : 148ull * 8;
    fastq_index_kernel<<<(unsigned)dense_grid, kFqThreads, 0, s>>>(d_bytes, n_bytes, sc.line_base, n_reads, nl, d_status, sc.overflow, sc.n_tiles, fmt);
    if (n_reads == 0) return cudaMemsetAsync(d_word_offsets, 0, sizeof(uint64_t), s);
    unsigned long long* sums2 = reinterpret_cast<unsigned long long*>(static_cast<char*>(d_index_scratch) + align16(n_reads * 32));
    const unsigned long long rec_blocks = ceil_div(n_reads, kThreads);
    fastq_records_kernel<<<(unsigned)(rec_blocks < 148ull * 8 ? rec_blocks : 148ull * 8), kThreads, 0, s>>>(nl, n_reads, d_seq_offsets, d_seq_lens,
                                                                                                      d_status, sc.overflow, fmt.shift);
    launch_exclusive_scan(WordsOfLen{d_seq_lens}, n_reads, sums2, d_word_offsets, s);
    return cudaGetLastError();
}

cudaError_t launch_fastq_encode(const DeviceInfo&, const uint8_t* d_bytes, size_t n_bytes, size_t n_reads, void* d_scratch,
                                const uint64_t* d_seq_offsets, const uint64_t* d_seq_lens, const uint64_t* d_word_offsets,
                                uint64_t* d_out_words, unsigned long long* d_status, int fasta, cudaStream_t s) {
    if (n_reads == 0 || n_bytes == 0) return cudaSuccess;
    const TextFormat fmt = text_format(fasta);
    const FqScratch sc(d_scratch, n_bytes);
    // tile / CTA shape: BN_FQ_VARIANT picks one of the measured shapes (profiles/r01_sweep_fastq.txt).  By default the
    // average record size decides how a read's two partial end vectors are validated: from per-vector maps kept in shared
    // memory (short reads: two end vectors per ~20 vectors of text; 1.59 ms against 1.86 on 150 bp reads), or by
    // fetching them again while phase 1 stays lighter (long reads: 1.21 ms against 1.51 on 10 kbp reads).
    static const int forced = [] {
        const char* v = getenv("BN_FQ_VARIANT");
        return v ? atoi(v) : -1;
    }();
    static const size_t short_max = [] {
        const char* v = getenv("BN_FQ_SHORT_MAX");
        return v ? (size_t)atoll(v) : (size_t)1024;
    }();
    if (forced < 0 && n_bytes / n_reads <= short_max) {   // short records: a thread per read, only the sequence bytes are fetched
        // (blocks in flight per thread, CTAs per SM) swept on 20 M x 150 bp: (6,3) 1.13 ms, (6,4) 1.19, (4,3) 1.23, (4,4) 1.08, (3,4) 1.10,
        // (3,5) 1.04, (2,6) 1.09, (6,2) 1.45 -- residency beats loads in flight per thread
        fastq_encode_reads_kernel<3, 5><<<(unsigned)ceil_div(n_reads, kThreads), kThreads, 0, s>>>(d_bytes, n_bytes, n_reads, d_seq_offsets,
                                                                                                  d_seq_lens, d_word_offsets, d_out_words, d_status);
        return cudaGetLastError();
    }
    static const size_t group_min = [] {
        const char* v = getenv("BN_FQ_GROUP_MIN");   // records above this many bytes on average take the lanes-per-read kernels
        return v ? (size_t)atoll(v) : (size_t)1024;
    }();
    if (forced < 0 && n_bytes / n_reads > group_min) {   // long records: G lanes per read, only the sequence bytes are fetched
        unsigned long long* giants = sc.giants;
        cudaError_t e = cudaMemsetAsync(giants, 0, sizeof(unsigned long long), s);
        if (e != cudaSuccess) return e;
        // words of an average sequence line: half of a FASTQ record is sequence, nearly all of a FASTA record
        const size_t avg_words = n_bytes / n_reads / (fasta ? 32 : 64);
#define BN_FQ_GROUPS(G)                                                                                                                  \
    fastq_encode_long_kernel<G><<<(unsigned)ceil_div(n_reads, 32 * kFqLongWarps / G), 32 * kFqLongWarps, 0, s>>>(                        \
        d_bytes, n_bytes, n_reads, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words, d_status, giants)
        if (avg_words > 48) BN_FQ_GROUPS(32);
        else if (avg_words > 20) BN_FQ_GROUPS(16);
        else BN_FQ_GROUPS(8);
#undef BN_FQ_GROUPS
        fastq_encode_giant_kernel<<<148 * 4, 32 * kFqLongWarps, 0, s>>>(d_bytes, n_bytes, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words, d_status,
                                                                        giants);
        return cudaGetLastError();
    }
    const int variant = forced >= 0 ? forced : (n_bytes / n_reads > 4096 ? 11 : 13);
#define BN_FQ_LAUNCH(TILE, THREADS, CTAS, ...)                                                                                         \
    fastq_encode_kernel<TILE, THREADS, CTAS, ##__VA_ARGS__><<<(unsigned)ceil_div(sc.n_tiles, TILE / kFqTile), THREADS, 0, s>>>(                        \
        d_bytes, n_bytes, sc.line_base, sc.n_tiles, n_reads, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words, d_status, fmt.shift)
    switch (variant) {
    case 1: BN_FQ_LAUNCH(65536, 256, 4); break;
    case 2: BN_FQ_LAUNCH(32768, 128, 8); break;
    case 3: BN_FQ_LAUNCH(32768, 128, 12); break;
    case 4: BN_FQ_LAUNCH(32768, 256, 6); break;
    case 5: BN_FQ_LAUNCH(65536, 128, 8, false); break;     // no maps: partial end vectors re-read as single sectors
    case 6: BN_FQ_LAUNCH(65536, 128, 10, false); break;
    case 7: BN_FQ_LAUNCH(49152, 128, 10, true); break;
    case 8: BN_FQ_LAUNCH(32768, 128, 10, true); break;
    case 9: BN_FQ_LAUNCH(65536, 128, 12, false); break;
    case 10: BN_FQ_LAUNCH(32768, 128, 12, false); break;
    case 11: BN_FQ_LAUNCH(49152, 128, 12, false); break;
    case 12: BN_FQ_LAUNCH(49152, 128, 8, true); break;
    case 13: BN_FQ_LAUNCH(49152, 128, 9, true); break;
    default: BN_FQ_LAUNCH(kFqEncTile, 128, 8); break;
    }
#undef BN_FQ_LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------- wrapped (multi-line) FASTA ------------------------
// A record = a '>' header line followed by any number of sequence lines (genome FASTA is wrapped at 60-80 columns); its
// sequence is the concatenation of those lines, so its bases are NOT contiguous in the text.  Done as: line table ->
// compaction -> the batch encode.
//   1. fastq_lines_kernel + scan (the count pass above)            -> number of lines, first line index of every tile
//   2. fastq_index_kernel in its dense form (forced)               -> nl[L] = position of line L's newline (| "a '\r' precedes it")
//   3. fasta_wrapped_lines_kernel, a thread per line               -> start, sequence length (0 for a header) and "is a header"
//      + a two-channel exclusive scan over the lines               -> H[L] = headers before line L, B[L] = sequence bytes before line L
//   4. fasta_wrapped_compact_kernel, a warp per line               -> header line: rec_off[H[L]] = B[L], hdr_off[H[L]] = start;
//                                                                     sequence line: its bytes to compact[B[L] ..)
//   5. launch_encode_batch(compact, rec_off)                        -> words, word offsets (every record on a fresh word)
// One more pass over the text than the one-line form and a copy of the sequence bytes: this is the convenience path for
// genome-style files, not a roofline row.  A text whose first line is not a header is a fault (record 0, FQ_BAD_HEADER).

__global__ void __launch_bounds__(kThreads)
fasta_wrapped_lines_kernel(const uint8_t* __restrict__ bytes, unsigned long long n, const uint64_t* __restrict__ nl, unsigned long long n_lines,
                           uint64_t* __restrict__ line_start, uint64_t* __restrict__ line_seq, uint8_t* __restrict__ line_hdr,
                           unsigned long long* __restrict__ status) {
    for (unsigned long long L = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; L < n_lines; L += (unsigned long long)gridDim.x * kThreads) {
        const unsigned long long e = nl[L], pos = e & ~kCrBit, cr = e >> 63;
        const unsigned long long s = L ? (nl[L - 1] & ~kCrBit) + 1 : 0ull;
        const bool hdr = s < n && bytes[s] == '>';
        line_start[L] = s;
        line_seq[L] = hdr ? 0ull : pos - cr - s;
        line_hdr[L] = hdr ? 1 : 0;
        if (L == 0 && !hdr) report_min(status + 1, FQ_BAD_HEADER);   // record 0: the text does not open with a header
    }
}

struct HeaderAndSeqLen {
    const uint8_t* line_hdr;
    const uint64_t* line_seq;
    __device__ __forceinline__ ulonglong2 operator()(unsigned long long L) const { return make_ulonglong2(line_hdr[L], line_seq[L]); }
};

__global__ void __launch_bounds__(kThreads)
fasta_wrapped_compact_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ line_start, const uint64_t* __restrict__ line_seq,
                             const uint8_t* __restrict__ line_hdr, const uint64_t* __restrict__ H, const uint64_t* __restrict__ B,
                             unsigned long long n_lines, uint8_t* __restrict__ compact, uint64_t* __restrict__ rec_off,
                             uint64_t* __restrict__ hdr_off) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    for (unsigned long long L = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); L < n_lines; L += n_warps) {
        const unsigned long long s = line_start[L], b = B[L];
        if (line_hdr[L]) {
            if (lane == 0) {
                rec_off[H[L]] = b;
                hdr_off[H[L]] = s;
            }
        } else {
            const unsigned long long len = line_seq[L];
            for (unsigned long long i = lane; i < len; i += 32) compact[b + i] = bytes[s + i];
        }
    }
}

// nl | line_start | line_seq | H | B (n_lines + 1 entries each) | line_hdr | scan sums | forced "dense" flag + a throw-away status pair
size_t fasta_wrapped_scratch_bytes(size_t n_lines) {
    const size_t m = n_lines + 1;
    return 5 * align16(m * 8) + align16(m) + align16(scan2_scratch_bytes(m)) + 48;
}

struct FwScratch {
    uint64_t *nl, *line_start, *line_seq, *H, *B;
    uint8_t* line_hdr;
    unsigned long long* sums;
    unsigned* force;
    unsigned long long* dummy_status;
    FwScratch(void* p, size_t n_lines) {
        const size_t m = n_lines + 1;
        char* c = static_cast<char*>(p);
        nl = reinterpret_cast<uint64_t*>(c), c += align16(m * 8);
        line_start = reinterpret_cast<uint64_t*>(c), c += align16(m * 8);
        line_seq = reinterpret_cast<uint64_t*>(c), c += align16(m * 8);
        H = reinterpret_cast<uint64_t*>(c), c += align16(m * 8);
        B = reinterpret_cast<uint64_t*>(c), c += align16(m * 8);
        line_hdr = reinterpret_cast<uint8_t*>(c), c += align16(m);
        sums = reinterpret_cast<unsigned long long*>(c), c += align16(scan2_scratch_bytes(m));
        force = reinterpret_cast<unsigned*>(c);
        dummy_status = reinterpret_cast<unsigned long long*>(c + 16);
    }
};

// after launch_fastq_count(..., fasta = 1): d_totals[0] = records, d_totals[1] = sequence bytes; d_status as launch_fastq_index
cudaError_t launch_fasta_wrapped_index(const DeviceInfo&, const uint8_t* d_bytes, size_t n_bytes, size_t n_lines, void* d_scratch,
                                       void* d_wscratch, uint64_t* d_totals, unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, 2 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_bytes == 0 || n_lines == 0) return cudaMemsetAsync(d_totals, 0, 2 * sizeof(uint64_t), s);
    const FqScratch sc(d_scratch, n_bytes);
    const FwScratch fw(d_wscratch, n_lines);
    e = cudaMemsetAsync(fw.force, 0x01, 4, s);   // non-zero: the dense index form runs whatever the count pass found
    if (e != cudaSuccess) return e;
    const unsigned long long dense_grid = sc.n_tiles < 148ull * 8 ? sc.n_tiles : 148ull * 8;
    fastq_index_kernel<<<(unsigned)dense_grid, kFqThreads, 0, s>>>(d_bytes, n_bytes, sc.line_base, n_lines, fw.nl, fw.dummy_status, fw.force,
                                                                   sc.n_tiles, TextFormat{0u, (uint32_t)'>'});
    const unsigned long long blocks = ceil_div(n_lines, kThreads);
    fasta_wrapped_lines_kernel<<<(unsigned)(blocks < 148ull * 16 ? blocks : 148ull * 16), kThreads, 0, s>>>(
        d_bytes, n_bytes, fw.nl, n_lines, fw.line_start, fw.line_seq, fw.line_hdr, d_status);
    launch_exclusive_scan2(HeaderAndSeqLen{fw.line_hdr, fw.line_seq}, n_lines, fw.sums, fw.H, fw.B, s);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d_totals, fw.H + n_lines, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(d_totals + 1, fw.B + n_lines, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
}

// d_compact: n_seq_bytes bytes (16-byte aligned); d_rec_off: n_records + 1 entries; d_hdr_off: n_records entries
cudaError_t launch_fasta_wrapped_compact(const DeviceInfo&, const uint8_t* d_bytes, size_t n_lines, void* d_wscratch, size_t n_records,
                                         uint8_t* d_compact, uint64_t* d_rec_off, uint64_t* d_hdr_off, cudaStream_t s) {
    if (n_lines == 0) return cudaMemsetAsync(d_rec_off, 0, sizeof(uint64_t), s);
    const FwScratch fw(d_wscratch, n_lines);
    const unsigned long long blocks = ceil_div(n_lines, kWarpsPerBlock);
    fasta_wrapped_compact_kernel<<<(unsigned)(blocks < 148ull * 32 ? blocks : 148ull * 32), kThreads, 0, s>>>(
        d_bytes, fw.line_start, fw.line_seq, fw.line_hdr, fw.H, fw.B, n_lines, d_compact, d_rec_off, d_hdr_off);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(d_rec_off + n_records, fw.B + n_lines, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
}

}  // namespace bn
