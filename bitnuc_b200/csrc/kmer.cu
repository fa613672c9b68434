// kmer.cu -- batched fixed-k records: as_2bit / from_2bit over n k-mers (k <= 32), sm_100a.
//
// Replaces the caller-side loops over the reference's single-word functions:
//   as_2bit:   /root/reference/src/utils/packing/mod.rs:81-110 (avx.rs:76-128, naive.rs:4-20)
//   from_2bit: /root/reference/src/utils/unpacking/mod.rs:119-147 (avx.rs:50-114, naive.rs:3-25)
//
// HBM-bound at (k + 8) bytes per record.  Three layouts per direction:
//   stride == 32 (padded records, 16-byte aligned): record r is vectors 2r, 2r+1 <-> word r, i.e. the
//       streaming codec layout with bytes >= k masked;
//   stride == k (tightly packed, e.g. 31-mers): a CTA packs its span of records as in the streaming encode into a
//       shared code strip and cuts each record out of it (three LDS + two funnel shifts);
//   stride <= 64: the CTA's byte span is staged through shared memory with coalesced 128-bit loads,
//       then each thread assembles its record with funnel shifts;
//   anything else: one thread per record with byte accesses (correct, not tuned).
#include "common.cuh"
#include "launch.cuh"

namespace bn {

// ============================================================================ as_2bit ========

// byte mask keeping the first `keep` (0..4) bytes of a word
__device__ __forceinline__ uint32_t keep_bytes(int keep) {
    return keep >= 4 ? 0xFFFFFFFFu : keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u);
}
// replace bytes [k - 4*word_index, ..) of word `w` (bytes 4*word_index.. of a record) by 'A'
__device__ __forceinline__ uint32_t mask_to_k(uint32_t w, int word_index, int k) {
    const uint32_t m = keep_bytes(k - 4 * word_index);
    return (w & m) | (0x41414141u & ~m);
}

__device__ __noinline__ void report_record_invalid(const uint8_t* rec, unsigned k, unsigned long long byte_offset,
                                                   unsigned long long* status) {
    for (unsigned i = 0; i < k; ++i) {
        const uint32_t b = rec[i];
        if (!byte_is_valid(b)) {
            report_invalid(status, byte_offset + i, b);
            return;
        }
    }
}

constexpr int kKmerU = 4;
constexpr int kKmerThreads = 512;

// Rare path of the padded kernel: first invalid byte among the kk bases of up to n_vec vectors (32 apart).
static __device__ __noinline__ void report_padded_invalid(const uint4* p, int n_vec, int kk, unsigned long long vec_index,
                                                          unsigned long long* status) {
    for (int j = 0; j < n_vec; ++j) {
        const uint8_t* rec = reinterpret_cast<const uint8_t*>(p + 32 * j);
        for (int i = 0; i < kk; ++i)
            if (!byte_is_valid(rec[i])) {
                report_invalid(status, (vec_index + 32ull * j) * 16ull + i, rec[i]);
                return;
            }
    }
}

// stride == 32, aligned: lane parity selects the low/high half of a record.
__global__ void __launch_bounds__(kKmerThreads, 2)
as_2bit_padded_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_vec, int k,
                      unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    const int half = (lane & 1) * 16;  // vectors 32*j + lane keep the lane's parity (ragged loop: THREADS is even too)
    const int kk = k - half < 0 ? 0 : (k - half > 16 ? 16 : k - half);  // bases of this half-record
    const uint32_t m0 = keep_bytes(kk), m1 = keep_bytes(kk - 4), m2 = keep_bytes(kk - 8), m3 = keep_bytes(kk - 12);
    constexpr unsigned kTile = 32 * kKmerU;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<kKmerThreads, 1, 1> walk(n_tiles);
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long v0 = t * kTile + lane;
        const uint4* p = in + v0;
        uint4 v[kKmerU];
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) v[j] = ld128<LD_PLAIN>(p + 32 * j);
        uint32_t bad = 0, r[kKmerU];
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) {  // bytes past k read as 'A': valid, code 0
            v[j].x = (v[j].x & m0) | (0x41414141u & ~m0);
            v[j].y = (v[j].y & m1) | (0x41414141u & ~m1);
            v[j].z = (v[j].z & m2) | (0x41414141u & ~m2);
            v[j].w = (v[j].w & m3) | (0x41414141u & ~m3);
            r[j] = pack16(v[j], bad);
        }
        uint32_t* q = out + v0;
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) st_stream_u32(q + 32 * j, r[j]);
        if (bad & kValidMask) report_padded_invalid(p, kKmerU, kk, v0, status);
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += kKmerThreads) {
            uint4 v = ld128<LD_PLAIN>(in + i);
            v.x = (v.x & m0) | (0x41414141u & ~m0);
            v.y = (v.y & m1) | (0x41414141u & ~m1);
            v.z = (v.z & m2) | (0x41414141u & ~m2);
            v.w = (v.w & m3) | (0x41414141u & ~m3);
            uint32_t bad = 0;
            out[i] = pack16(v, bad);
            if (bad & kValidMask) report_padded_invalid(in + i, 1, kk, i, status);
        }
    }
}

// Tightly packed records (stride == k): every byte of the buffer belongs to a record, so the bytes
// are packed exactly like the streaming encode (aligned 16-byte vectors, coalesced, validated as a whole)
// and each record is then assembled from the <= 3 vectors it straddles with two funnel shifts.

__device__ __noinline__ uint4 load_vector_at_edge(const uint8_t* recs, long long vbyte, unsigned long long total) {
    uint32_t w[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};  // 'A' outside the buffer
    for (int j = 0; j < 16; ++j) {
        const long long o = vbyte + j;
        if (o >= 0 && (unsigned long long)o < total)
            w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)recs[o] << (8 * (j & 3)));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __noinline__ void report_vector_invalid(uint4 v, long long vbyte, unsigned long long* status) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 16; ++j) {
        const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (!byte_is_valid(b)) {
            report_invalid(status, (unsigned long long)(vbyte + j), b);
            return;
        }
    }
}

// Pack, then cut: a CTA owns `rpc` consecutive records = one span of <= 64 KiB.  It packs the span as a pure stream
// (aligned 128-bit loads, U in flight per thread, validated as whole vectors) into a 16 KiB code strip in shared
// memory, then cuts the records out of the strip -- record j is the 2k-bit field starting 2 * (its byte offset) bits
// into the strip: three LDS + two funnel shifts -- one record per thread step, consecutive threads -> consecutive
// output words.  (A per-warp strip with shuffled / warp-private assembly was 83 % of the measured peak, this is
// 86-88 %; the first version, nine LDS.32 per record from staged raw bytes, 54 %.)
template <int THREADS, int U, int SPAN_KB = 64>
__global__ void __launch_bounds__(THREADS)
as_2bit_tight_kernel(const uint8_t* __restrict__ recs, unsigned long long n, unsigned k, unsigned rpc,
                         uint64_t* __restrict__ out, unsigned long long* __restrict__ status) {
    __shared__ uint32_t codes[SPAN_KB * 64 + 8];
    const unsigned tid = threadIdx.x;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * rpc;
    const unsigned cnt = (unsigned)(n - r0 < rpc ? n - r0 : rpc);
    const uint8_t* base = recs + r0 * k;
    const unsigned m0 = (unsigned)(reinterpret_cast<uintptr_t>(base) & 15u);
    const uint8_t* abase = base - m0;
    const unsigned nvec = (m0 + cnt * k + 15u) / 16u;                              // <= 4097
    const bool interior = abase >= recs && abase + 16ull * nvec <= recs + n * k;
    const uint4* src = reinterpret_cast<const uint4*>(abase);
    unsigned v = tid;
    if (interior) {
        for (; v + (U - 1) * THREADS < nvec; v += U * THREADS) {
            uint4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) x[j] = ld128<LD_NC_NOALLOC>(src + v + j * THREADS);
            uint32_t bad = 0;
#pragma unroll
            for (int j = 0; j < U; ++j) codes[v + j * THREADS] = pack16(x[j], bad);
            if (bad & kValidMask) {  // rare; unrolled so that x[] is never indexed dynamically (that would put it in local memory)
#pragma unroll
                for (int j = 0; j < U; ++j) report_vector_invalid(x[j], (long long)(abase - recs) + 16ll * (v + j * THREADS), status);
            }
        }
    }
    for (; v < nvec; v += THREADS) {
        const long long vbyte = (long long)(abase - recs) + 16ll * v;
        const uint4 x = vbyte >= 0 && (unsigned long long)vbyte + 16 <= n * k ? ld128<LD_NC_NOALLOC>(src + v)
                                                                              : load_vector_at_edge(recs, vbyte, n * k);
        uint32_t bad = 0;
        codes[v] = pack16(x, bad);
        if (bad & kValidMask) report_vector_invalid(x, vbyte, status);
    }
    __syncthreads();
    const uint32_t mlo = k >= 16 ? 0xFFFFFFFFu : (1u << (2 * k)) - 1u;
    const uint32_t mhi = k >= 32 ? 0xFFFFFFFFu : k <= 16 ? 0u : (1u << (2 * k - 32)) - 1u;
    uint2* o = reinterpret_cast<uint2*>(out + r0);
#pragma unroll 4
    for (unsigned j = tid; j < cnt; j += THREADS) {
        const unsigned rel = m0 + j * k;
        const unsigned vi = rel >> 4, sh = 2 * (rel & 15u);
        const uint32_t c0 = codes[vi], c1 = codes[vi + 1], c2 = codes[vi + 2];
        st_stream_v2(o + j, make_uint2(__funnelshift_r(c0, c1, sh) & mlo, __funnelshift_r(c1, c2, sh) & mhi));
    }
}

// stride <= 64: stage the CTA's span in shared memory, one thread per record.
constexpr int kStageRecords = kThreads;
constexpr int kStageMaxStride = 64;
constexpr int kStageBytes = kStageRecords * kStageMaxStride + 64;

__global__ void __launch_bounds__(kThreads)
as_2bit_staged_kernel(const uint8_t* __restrict__ recs, unsigned long long n, unsigned k, unsigned stride,
                      uint64_t* __restrict__ out, unsigned long long* __restrict__ status) {
    __shared__ __align__(16) uint8_t smem[kStageBytes];
    const unsigned long long total_bytes = (n - 1) * stride + k;
    {   // one group of kStageRecords records per CTA, handed out by the hardware CTA scheduler
        const unsigned long long r0 = (unsigned long long)blockIdx.x * kStageRecords;
        const unsigned cnt = (unsigned)(n - r0 < kStageRecords ? n - r0 : kStageRecords);
        const unsigned long long b0 = r0 * stride;                               // first byte of the span
        const unsigned long long b1 = b0 + (unsigned long long)(cnt - 1) * stride + k;  // one past the last
        const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(recs) + b0) & 15u);
        const uint8_t* base = recs + b0 - mis;                                    // 16-byte aligned
        const unsigned n_vec = (unsigned)((mis + (b1 - b0) + 15) / 16);
        for (unsigned i = threadIdx.x; i < n_vec; i += blockDim.x) {
            const long long lo = (long long)i * 16 - mis;  // span-relative offset of this vector
            uint4 v;
            if ((long long)b0 + lo >= 0 && (long long)b0 + lo + 16 <= (long long)total_bytes) {
                v = ld128<LD_PLAIN>(reinterpret_cast<const uint4*>(base) + i);
            } else {  // vector straddles the buffer edge: byte loads with bounds
                uint32_t w[4] = {0, 0, 0, 0};
                for (int j = 0; j < 16; ++j) {
                    const long long o = lo + j;
                    if ((long long)b0 + o >= 0 && (long long)b0 + o < (long long)total_bytes)
                        w[j >> 2] |= (uint32_t)recs[(long long)b0 + o] << (8 * (j & 3));
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            reinterpret_cast<uint4*>(smem)[i] = v;
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            const unsigned o = mis + threadIdx.x * stride;
            const uint32_t* ws = reinterpret_cast<const uint32_t*>(smem) + (o >> 2);
            const unsigned sh = 8 * (o & 3);
            uint32_t x[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) x[i] = ws[i];
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = mask_to_k(__funnelshift_r(x[i], x[i + 1], sh), i, (int)k);
            uint32_t bad = 0;
            const uint32_t lo = pack16(make_uint4(w[0], w[1], w[2], w[3]), bad);
            const uint32_t hi = pack16(make_uint4(w[4], w[5], w[6], w[7]), bad);
            out[r0 + threadIdx.x] = ((uint64_t)hi << 32) | lo;
            if (bad & kValidMask) report_record_invalid(smem + o, k, (r0 + threadIdx.x) * stride, status);
        }
    }
}

// any stride: one thread per record, byte loads.
__global__ void __launch_bounds__(kThreads)
as_2bit_generic_kernel(const uint8_t* __restrict__ recs, unsigned long long n, unsigned k, unsigned long long stride,
                       uint64_t* __restrict__ out, unsigned long long* __restrict__ status) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) {
        const uint8_t* rec = recs + r * stride;
        uint64_t packed = 0;
        for (unsigned i = 0; i < k; ++i) {
            const uint32_t b = rec[i];
            if (!byte_is_valid(b)) {
                report_invalid(status, r * stride + i, b);
                break;
            }
            packed |= (uint64_t)(((b >> 1) ^ (b >> 2)) & 3u) << (2 * i);
        }
        out[r] = packed;
    }
}

// ============================================================================ from_2bit ======

constexpr int kTightChunks = 4;  // 16-byte output chunks per thread, all loads issued before the first use

// Tightly packed records (stride == k, 16 <= k <= 32), 16-byte aligned output: one thread per
// 16-byte output chunk; a chunk spans at most two records.  A CTA owns kKmerThreads * kTightChunks
// consecutive chunks; (record, position) comes from one 64-bit division per CTA and one 32-bit division per thread,
// then advances.
__global__ void __launch_bounds__(kKmerThreads, 2)
from_2bit_tight_kernel(const uint64_t* __restrict__ packed, unsigned long long n, unsigned k,
                       uint8_t* __restrict__ out) {
    const unsigned long long total = n * k;
    const unsigned long long n_chunks = total / 16;
    const unsigned long long c0 = (unsigned long long)blockIdx.x * (kKmerThreads * kTightChunks) + threadIdx.x;
    // (record, position) of the CTA's first byte: one 64-bit division per CTA; the threads then only need a 32-bit one
    __shared__ unsigned long long r_cta;
    __shared__ unsigned p_cta;
    if (threadIdx.x == 0) {
        const unsigned long long b = (unsigned long long)blockIdx.x * (kKmerThreads * kTightChunks) * 16ull;
        r_cta = b / k;
        p_cta = (unsigned)(b % k);
    }
    __syncthreads();
    const unsigned t = p_cta + threadIdx.x * 16u;
    unsigned long long r = r_cta + t / k;
    unsigned pos = t % k;
    constexpr unsigned kStepBytes = kKmerThreads * 16;
    const unsigned dr = kStepBytes / k, dpos = kStepBytes % k;
    uint64_t w0[kTightChunks], w1[kTightChunks];
    unsigned p[kTightChunks];
#pragma unroll
    for (int it = 0; it < kTightChunks; ++it) {
        p[it] = pos;
        const bool live = c0 + (unsigned long long)it * kKmerThreads < n_chunks;
        w0[it] = live ? __ldg(packed + r) : 0ull;
        w1[it] = live && r + 1 < n ? __ldg(packed + r + 1) : 0ull;  // only used when the chunk crosses into record r+1
        r += dr;
        pos += dpos;
        if (pos >= k) {
            pos -= k;
            ++r;
        }
    }
#pragma unroll
    for (int it = 0; it < kTightChunks; ++it) {
        const unsigned long long c = c0 + (unsigned long long)it * kKmerThreads;
        if (c >= n_chunks) break;
        const unsigned avail = k - p[it];
        uint32_t x = (uint32_t)(w0[it] >> (2 * p[it]));
        if (avail < 16) x = (x & ((1u << (2 * avail)) - 1u)) | (uint32_t)(w1[it] << (2 * avail));
        st_stream_v4(reinterpret_cast<uint4*>(out) + c, decode16_prmt(x));
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {  // trailing < 16 bytes
        for (unsigned long long b = n_chunks * 16; b < total; ++b) {
            const uint64_t w = packed[b / k];
            out[b] = (uint8_t)(0x54474341u >> (8 * (unsigned)((w >> (2 * (b % k))) & 3u)));
        }
    }
}

// k = 31, tightly packed (the 31-mer batch of BASELINE configs[2]): 16 records are exactly 31 output chunks of 16 bytes
// (lcm(31, 16) = 496), so a warp takes groups of 16 records and lane l < 31 always owns chunk l of the group -- the record
// it starts in (16 l / 31), its first base there (16 l mod 31) and whether it spills into the next record are constants
// of the lane, computed once.  The generic tight kernel above re-derives (record, position) for every chunk and always
// fetches two words; here a chunk is one or two loads, two shifts and the 16-base decode.  Lane 31 idles (3 % of the lanes).
constexpr int kK31Groups = 4;   // groups per warp step, all loads issued before the first use
__global__ void __launch_bounds__(kKmerThreads, 2)
from_2bit_tight31_kernel(const uint64_t* __restrict__ packed, unsigned long long n, uint8_t* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long n_groups = n / 16;              // whole groups; the remainder goes through the tail below
    const unsigned rec = (16u * lane) / 31u, pos = (16u * lane) % 31u;
    const unsigned avail = 31u - pos;                        // bases of this chunk that come from its first record
    const bool spills = avail < 16u, active = lane < 31u;
    const unsigned sh0 = 2u * pos, sh1 = 2u * avail;
    const uint32_t keep = spills ? (1u << sh1) - 1u : 0xFFFFFFFFu;
    const TileWalk<kKmerThreads, 1, 1> walk(ceil_div(n_groups, kK31Groups));
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long g0 = t * kK31Groups;
        uint64_t w0[kK31Groups], w1[kK31Groups];
#pragma unroll
        for (int j = 0; j < kK31Groups; ++j) {
            const bool live = active && g0 + j < n_groups;
            const uint64_t* p = packed + (g0 + j) * 16 + rec;
            w0[j] = live ? __ldg(p) : 0ull;
            w1[j] = live && spills ? __ldg(p + 1) : 0ull;   // (a spilling chunk never starts in a group's last record)
        }
#pragma unroll
        for (int j = 0; j < kK31Groups; ++j) {
            if (active && g0 + j < n_groups) {
                const uint32_t x = ((uint32_t)(w0[j] >> sh0) & keep) | (spills ? (uint32_t)(w1[j] << sh1) : 0u);
                st_stream_v4(reinterpret_cast<uint4*>(out) + (g0 + j) * 31 + lane, decode16_prmt(x));
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {   // the records past the last whole group (< 16 of them)
        for (unsigned long long b = n_groups * 496; b < n * 31; ++b) {
            const uint64_t w = packed[b / 31];
            out[b] = (uint8_t)(0x54474341u >> (8 * (unsigned)((w >> (2 * (b % 31))) & 3u)));
        }
    }
}

// stride == 32, 16-byte aligned output: word r -> vectors 2r, 2r+1 (all 32 slots are written).
__global__ void __launch_bounds__(kKmerThreads)
from_2bit_padded_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, unsigned long long n_w32) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * kKmerU;
    const unsigned long long n_tiles = n_w32 / kTile;
    const TileWalk<kKmerThreads, 1, 1> walk(n_tiles);
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long i0 = t * kTile + lane;
        uint32_t w[kKmerU];
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) w[j] = ld32<LD_PLAIN>(in + i0 + 32 * j);
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) st_stream_v4(out + i0 + 32 * j, decode16_prmt(w[j]));
    }
    if (blockIdx.x == gridDim.x - 1)
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_w32; i += kKmerThreads)
            st_stream_v4(out + i, decode16_prmt(ld32<LD_PLAIN>(in + i)));
}

// any k / stride: one thread per record, byte stores of exactly k bytes.
__global__ void __launch_bounds__(kThreads)
from_2bit_generic_kernel(const uint64_t* __restrict__ packed, unsigned long long n, unsigned k,
                         uint8_t* __restrict__ out, unsigned long long stride) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) {
        const uint64_t w = packed[r];
        uint8_t* o = out + r * stride;
        for (unsigned i = 0; i < k; ++i) o[i] = (uint8_t)(0x54474341u >> (8 * (unsigned)((w >> (2 * i)) & 3u)));
    }
}

// ============================================================================ launchers ======

cudaError_t launch_as_2bit_batch(const DeviceInfo& di, const uint8_t* d_recs, size_t n, uint32_t k, size_t stride,
                                 uint64_t* d_out, unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || n == 0) return e;
    if (k == 0) return cudaMemsetAsync(d_out, 0, n * sizeof(uint64_t), s);
    if (stride == 32 && (reinterpret_cast<uintptr_t>(d_recs) & 15u) == 0) {
        const unsigned long long n_vec = 2ull * n;
        const unsigned long long ctas = TileWalk<kKmerThreads, 1, 1>::ctas(n_vec / (32 * kKmerU));
        as_2bit_padded_kernel<<<(unsigned)(ctas ? ctas : 1), kKmerThreads, 0, s>>>(
            reinterpret_cast<const uint4*>(d_recs), reinterpret_cast<uint32_t*>(d_out), n_vec, (int)k, d_status);
    } else if (stride == k) {
        const unsigned rpc = (65536u - 32u) / k < 2048u ? ((65536u - 32u) / k) / 256u * 256u : 2048u;   // records per CTA: span <= 64 KiB
        as_2bit_tight_kernel<256, 4><<<(unsigned)ceil_div(n, rpc), 256, 0, s>>>(d_recs, n, k, rpc, d_out, d_status);
    } else if (stride <= (size_t)kStageMaxStride) {
        as_2bit_staged_kernel<<<(unsigned)ceil_div(n, kStageRecords), kThreads, 0, s>>>(d_recs, n, k, (unsigned)stride, d_out,
                                                                                         d_status);
    } else {
        static const int per_sm = blocks_per_sm(as_2bit_generic_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        as_2bit_generic_kernel<<<grid_for(ceil_div(n, kThreads), resident), kThreads, 0, s>>>(d_recs, n, k, stride, d_out,
                                                                                               d_status);
    }
    return cudaGetLastError();
}

cudaError_t launch_from_2bit_batch(const DeviceInfo& di, const uint64_t* d_packed, size_t n, uint32_t k,
                                   uint8_t* d_out, size_t stride, cudaStream_t s) {
    if (n == 0 || k == 0) return cudaSuccess;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    if (aligned && stride == 32 && k == 32) {
        const unsigned long long n_w32 = 2ull * n;
        const unsigned long long ctas = TileWalk<kKmerThreads, 1, 1>::ctas(n_w32 / (32 * kKmerU));
        from_2bit_padded_kernel<<<(unsigned)(ctas ? ctas : 1), kKmerThreads, 0, s>>>(
            reinterpret_cast<const uint32_t*>(d_packed), reinterpret_cast<uint4*>(d_out), n_w32);
    } else if (aligned && stride == 31 && k == 31) {
        const unsigned long long steps = ceil_div(n / 16, kK31Groups);
        const unsigned long long ctas = TileWalk<kKmerThreads, 1, 1>::ctas(steps);
        from_2bit_tight31_kernel<<<(unsigned)(ctas ? ctas : 1), kKmerThreads, 0, s>>>(d_packed, n, d_out);
    } else if (aligned && stride == k && k >= 16) {
        const unsigned long long chunks = (unsigned long long)n * k / 16;
        from_2bit_tight_kernel<<<(unsigned)ceil_div(chunks + 1, kKmerThreads * kTightChunks), kKmerThreads, 0, s>>>(d_packed, n, k,
                                                                                                                 d_out);
    } else {
        static const int per_sm = blocks_per_sm(from_2bit_generic_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        from_2bit_generic_kernel<<<grid_for(ceil_div(n, kThreads), resident), kThreads, 0, s>>>(d_packed, n, k, d_out, stride);
    }
    return cudaGetLastError();
}

}  // namespace bn
