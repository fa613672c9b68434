// kmer.cu -- batched fixed-k records: as_2bit / from_2bit over n k-mers (k <= 32), sm_100a.
//
// Replaces the caller-side loops over the reference's single-word functions:
//   as_2bit:   /root/reference/src/utils/packing/mod.rs:81-110 (avx.rs:76-128, naive.rs:4-20)
//   from_2bit: /root/reference/src/utils/unpacking/mod.rs:119-147 (avx.rs:50-114, naive.rs:3-25)
//
// HBM-bound at (k + 8) bytes per record.  Three layouts per direction:
//   stride == 32 (padded records, 16-byte aligned): record r is vectors 2r, 2r+1 <-> word r, i.e. the
//       streaming codec layout with bytes >= k masked;
//   stride <= 64 (e.g. tightly packed 31-mers): the CTA's byte span is staged through shared memory
//       with coalesced 128-bit loads, then each thread assembles its record with funnel shifts;
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

// stride == 32, aligned: lane parity selects the low/high half of a record.
__global__ void __launch_bounds__(kKmerThreads)
as_2bit_padded_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_vec, int k,
                      unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    const int half = (lane & 1) * 16;  // vectors 32*j + lane keep the lane's parity (ragged loop: THREADS is even too)
    const uint32_t m0 = keep_bytes(k - half), m1 = keep_bytes(k - half - 4), m2 = keep_bytes(k - half - 8),
                   m3 = keep_bytes(k - half - 12);
    const int kk = k - half < 0 ? 0 : (k - half > 16 ? 16 : k - half);  // bases of this half-record
    constexpr unsigned kTile = 32 * kKmerU;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<kKmerThreads, 1, 1> walk(n_tiles);
    auto masked = [&](uint4 v) {
        v.x = (v.x & m0) | (0x41414141u & ~m0);
        v.y = (v.y & m1) | (0x41414141u & ~m1);
        v.z = (v.z & m2) | (0x41414141u & ~m2);
        v.w = (v.w & m3) | (0x41414141u & ~m3);
        return v;
    };
    auto report = [&](unsigned long long vi) {
        const uint8_t* rec = reinterpret_cast<const uint8_t*>(in + vi);
        for (int i = 0; i < kk; ++i)
            if (!byte_is_valid(rec[i])) {
                report_invalid(status, vi * 16ull + i, rec[i]);
                return true;
            }
        return false;
    };
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long v0 = t * kTile + lane;
        uint4 v[kKmerU];
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) v[j] = ld128<LD_PLAIN>(in + v0 + 32 * j);
        uint32_t bad = 0, r[kKmerU];
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) r[j] = pack16(masked(v[j]), bad);
#pragma unroll
        for (int j = 0; j < kKmerU; ++j) st_stream_u32(out + v0 + 32 * j, r[j]);
        if (bad & kValidMask)
            for (int j = 0; j < kKmerU; ++j)
                if (report(v0 + 32 * j)) break;
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += kKmerThreads) {
            uint32_t bad = 0;
            out[i] = pack16(masked(ld128<LD_PLAIN>(in + i)), bad);
            if (bad & kValidMask) report(i);
        }
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

constexpr int kTightChunks = 8;  // 16-byte output chunks per thread

// Tightly packed records (stride == k, 16 <= k <= 32), 16-byte aligned output: one thread per
// 16-byte output chunk; a chunk spans at most two records.  A CTA owns kKmerThreads * kTightChunks
// consecutive chunks; (record, position) is found by one division per thread, then advanced.
__global__ void __launch_bounds__(kKmerThreads)
from_2bit_tight_kernel(const uint64_t* __restrict__ packed, unsigned long long n, unsigned k,
                       uint8_t* __restrict__ out) {
    const unsigned long long total = n * k;
    const unsigned long long n_chunks = total / 16;
    const unsigned long long c0 = (unsigned long long)blockIdx.x * (kKmerThreads * kTightChunks) + threadIdx.x;
    unsigned long long r = (c0 * 16) / k;
    unsigned pos = (unsigned)((c0 * 16) % k);
    constexpr unsigned kStepBytes = kKmerThreads * 16;
    const unsigned long long dr = kStepBytes / k;
    const unsigned dpos = kStepBytes % k;
#pragma unroll 2
    for (int it = 0; it < kTightChunks; ++it) {
        const unsigned long long c = c0 + (unsigned long long)it * kKmerThreads;
        if (c >= n_chunks) break;
        const unsigned avail = k - pos;
        const uint64_t w0 = __ldg(packed + r);
        uint32_t x = (uint32_t)(w0 >> (2 * pos));
        if (avail < 16) {
            const uint64_t w1 = __ldg(packed + r + 1);  // exists: the chunk is full, so bytes follow
            x = (x & ((1u << (2 * avail)) - 1u)) | (uint32_t)(w1 << (2 * avail));
        }
        st_stream_v4(reinterpret_cast<uint4*>(out) + c, decode16_prmt(x));
        r += dr;
        pos += dpos;
        if (pos >= k) {
            pos -= k;
            ++r;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {  // trailing < 16 bytes
        for (unsigned long long b = n_chunks * 16; b < total; ++b) {
            const uint64_t w = packed[b / k];
            out[b] = (uint8_t)(0x54474341u >> (8 * (unsigned)((w >> (2 * (b % k))) & 3u)));
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
    } else if (stride <= (size_t)kStageMaxStride) {
        as_2bit_staged_kernel<<<(unsigned)ceil_div(n, kStageRecords), kThreads, 0, s>>>(d_recs, n, k, (unsigned)stride, d_out,
                                                                                         d_status);
    } else {
        static const int resident = resident_blocks(as_2bit_generic_kernel, kThreads, di);
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
    } else if (aligned && stride == k && k >= 16) {
        const unsigned long long chunks = (unsigned long long)n * k / 16;
        from_2bit_tight_kernel<<<(unsigned)ceil_div(chunks + 1, kKmerThreads * kTightChunks), kKmerThreads, 0, s>>>(d_packed, n, k,
                                                                                                                 d_out);
    } else {
        static const int resident = resident_blocks(from_2bit_generic_kernel, kThreads, di);
        from_2bit_generic_kernel<<<grid_for(ceil_div(n, kThreads), resident), kThreads, 0, s>>>(d_packed, n, k, d_out, stride);
    }
    return cudaGetLastError();
}

}  // namespace bn
