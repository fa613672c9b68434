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

// stride == 32, aligned: lane parity selects the low/high half of a record.
template <int U>
__global__ void __launch_bounds__(kThreads)
as_2bit_padded_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_vec, int k,
                      unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    const int half = (lane & 1) * 16;  // vectors 32*j + lane keep the lane's parity
    const uint32_t m0 = keep_bytes(k - half), m1 = keep_bytes(k - half - 4), m2 = keep_bytes(k - half - 8),
                   m3 = keep_bytes(k - half - 12);
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long warp = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = ceil_div(n_vec, kTile);
    for (unsigned long long t = warp; t < n_tiles; t += n_warps) {
        const unsigned long long v0 = t * kTile + lane;
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j)
            v[j] = v0 + 32 * j < n_vec ? ld_stream_v4(in + v0 + 32 * j) : make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        uint32_t bad = 0, r[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            v[j].x = (v[j].x & m0) | (0x41414141u & ~m0);
            v[j].y = (v[j].y & m1) | (0x41414141u & ~m1);
            v[j].z = (v[j].z & m2) | (0x41414141u & ~m2);
            v[j].w = (v[j].w & m3) | (0x41414141u & ~m3);
            r[j] = pack16(v[j], bad);
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (v0 + 32 * j < n_vec) st_stream_u32(out + v0 + 32 * j, r[j]);
        if (bad & kValidMask) {
            for (int j = 0; j < U; ++j) {
                const unsigned long long vi = v0 + 32 * j;
                if (vi >= n_vec) break;
                const int kk = k - half < 0 ? 0 : (k - half > 16 ? 16 : k - half);
                const uint8_t* rec = reinterpret_cast<const uint8_t*>(in + vi);
                bool found = false;
                for (int i = 0; i < kk; ++i)
                    if (!byte_is_valid(rec[i])) {
                        report_invalid(status, vi * 16ull + i, rec[i]);
                        found = true;
                        break;
                    }
                if (found) break;
            }
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
    const unsigned long long n_groups = ceil_div(n, kStageRecords);
    for (unsigned long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const unsigned long long r0 = g * kStageRecords;
        const unsigned cnt = (unsigned)(n - r0 < kStageRecords ? n - r0 : kStageRecords);
        const unsigned long long b0 = r0 * stride;                               // first byte of the span
        const unsigned long long b1 = b0 + (unsigned long long)(cnt - 1) * stride + k;  // one past the last
        const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(recs) + b0) & 15u);
        const uint8_t* base = recs + b0 - mis;                                    // 16-byte aligned
        const unsigned n_vec = (unsigned)((mis + (b1 - b0) + 15) / 16);
        __syncthreads();  // previous group's readers are done
        for (unsigned i = threadIdx.x; i < n_vec; i += blockDim.x) {
            const long long lo = (long long)i * 16 - mis;  // span-relative offset of this vector
            uint4 v;
            if ((long long)b0 + lo >= 0 && (long long)b0 + lo + 16 <= (long long)total_bytes) {
                v = ld_stream_v4(reinterpret_cast<const uint4*>(base) + i);
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

constexpr int kLutWords = 256 * 32;  // per-lane replicated 256-entry table, see codec.cu

__device__ __forceinline__ void lut_init(uint32_t* lut) {
    for (int i = threadIdx.x; i < kLutWords; i += blockDim.x) lut[i] = ascii4_of_byte((uint32_t)i >> 5);
    __syncthreads();
}
__device__ __forceinline__ uint4 lut_decode16(uint32_t w, const uint32_t* lut_lane) {
    return make_uint4(lut_lane[(w & 0xFFu) << 5], lut_lane[((w >> 8) & 0xFFu) << 5],
                      lut_lane[((w >> 16) & 0xFFu) << 5], lut_lane[(w >> 24) << 5]);
}

// Tightly packed records (stride == k, 16 <= k <= 32), 16-byte aligned output: one thread per
// 16-byte output chunk; a chunk spans at most two records.
__global__ void __launch_bounds__(kThreads)
from_2bit_tight_kernel(const uint64_t* __restrict__ packed, unsigned long long n, unsigned k,
                       uint8_t* __restrict__ out) {
    __shared__ uint32_t lut[kLutWords];
    lut_init(lut);
    const uint32_t* lut_lane = lut + (threadIdx.x & 31);
    const unsigned long long total = n * k;
    const unsigned long long n_chunks = total / 16;
    const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long first = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    // (record, position) of byte 16*first, advanced incrementally by 16*T bytes per round
    unsigned long long r = (first * 16) / k;
    unsigned pos = (unsigned)((first * 16) % k);
    const unsigned long long dr = (T * 16) / k;
    const unsigned dpos = (unsigned)((T * 16) % k);
    for (unsigned long long c = first; c < n_chunks; c += T) {
        const unsigned avail = k - pos;
        const uint64_t w0 = __ldg(packed + r);
        uint32_t x = (uint32_t)(w0 >> (2 * pos));
        if (avail < 16) {
            const uint64_t w1 = __ldg(packed + r + 1);  // exists: the chunk is full, so bytes follow
            x = (x & ((1u << (2 * avail)) - 1u)) | (uint32_t)(w1 << (2 * avail));
        }
        st_stream_v4(reinterpret_cast<uint4*>(out) + c, lut_decode16(x, lut_lane));
        r += dr;
        pos += dpos;
        if (pos >= k) {
            pos -= k;
            ++r;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // trailing < 16 bytes
        for (unsigned long long b = n_chunks * 16; b < total; ++b) {
            const uint64_t w = packed[b / k];
            out[b] = (uint8_t)(0x54474341u >> (8 * (unsigned)((w >> (2 * (b % k))) & 3u)));
        }
    }
}

// stride == 32, 16-byte aligned output: word r -> vectors 2r, 2r+1 (all 32 slots are written).
template <int U>
__global__ void __launch_bounds__(kThreads)
from_2bit_padded_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, unsigned long long n_w32) {
    __shared__ uint32_t lut[kLutWords];
    lut_init(lut);
    const unsigned lane = threadIdx.x & 31;
    const uint32_t* lut_lane = lut + lane;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long warp = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = ceil_div(n_w32, kTile);
    for (unsigned long long t = warp; t < n_tiles; t += n_warps) {
        const unsigned long long i0 = t * kTile + lane;
        uint32_t w[U];
#pragma unroll
        for (int j = 0; j < U; ++j) w[j] = i0 + 32 * j < n_w32 ? ld_stream_u32(in + i0 + 32 * j) : 0u;
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (i0 + 32 * j < n_w32) st_stream_v4(out + i0 + 32 * j, lut_decode16(w[j], lut_lane));
    }
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
        constexpr int U = 4;
        static const int resident = resident_blocks(as_2bit_padded_kernel<U>, kThreads, di);
        const unsigned long long n_vec = 2ull * n;
        as_2bit_padded_kernel<U><<<grid_for(ceil_div(ceil_div(n_vec, 32 * U), kWarpsPerBlock), resident), kThreads, 0, s>>>(
            reinterpret_cast<const uint4*>(d_recs), reinterpret_cast<uint32_t*>(d_out), n_vec, (int)k, d_status);
    } else if (stride <= (size_t)kStageMaxStride) {
        static const int resident = resident_blocks(as_2bit_staged_kernel, kThreads, di);
        as_2bit_staged_kernel<<<grid_for(ceil_div(n, kStageRecords), resident), kThreads, 0, s>>>(
            d_recs, n, k, (unsigned)stride, d_out, d_status);
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
        constexpr int U = 4;
        static const int resident = resident_blocks(from_2bit_padded_kernel<U>, kThreads, di);
        const unsigned long long n_w32 = 2ull * n;
        from_2bit_padded_kernel<U><<<grid_for(ceil_div(ceil_div(n_w32, 32 * U), kWarpsPerBlock), resident), kThreads, 0, s>>>(
            reinterpret_cast<const uint32_t*>(d_packed), reinterpret_cast<uint4*>(d_out), n_w32);
    } else if (aligned && stride == k && k >= 16) {
        static const int resident = resident_blocks(from_2bit_tight_kernel, kThreads, di);
        const unsigned long long chunks = (unsigned long long)n * k / 16;
        from_2bit_tight_kernel<<<grid_for(ceil_div(chunks + 1, kThreads), resident), kThreads, 0, s>>>(d_packed, n, k, d_out);
    } else {
        static const int resident = resident_blocks(from_2bit_generic_kernel, kThreads, di);
        from_2bit_generic_kernel<<<grid_for(ceil_div(n, kThreads), resident), kThreads, 0, s>>>(d_packed, n, k, d_out, stride);
    }
    return cudaGetLastError();
}

}  // namespace bn
