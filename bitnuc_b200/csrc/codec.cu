// codec.cu -- ASCII <-> 2-bit streaming codec kernels (sm_100a).
//
// Replaces the reference's per-ISA kernels behind encode/decode:
//   encode: /root/reference/src/utils/packing/avx.rs:76-151 (naive.rs:4-43 defines the results)
//   decode: /root/reference/src/utils/unpacking/avx.rs:26-33,117-153 (naive.rs:3-25)
//
// Both kernels are HBM-bound at 1.25 algorithmic bytes per base (encode 1 B read + 0.25 B write,
// decode the reverse).  Data layout: the ASCII stream is viewed as 16-byte vectors, the packed
// stream as 32-bit words; vector i <-> word i, so lane l of a warp always touches element
// tile_base + 32*j + l and every global access is a fully coalesced 512-byte (vector) or 128-byte
// (word) warp transaction.  Grids are persistent: sm_count x resident CTAs, each warp walking
// tiles of 32*U vectors with all U loads issued before the first use.
#include "common.cuh"
#include "launch.cuh"

namespace bn {

// ============================================================================ encode =========

// Rare path: re-read this thread's vectors (address order = j order) and report the first invalid byte.
__device__ __noinline__ void report_first_invalid(const uint4* p, int n_vec, unsigned long long vec_index,
                                                  unsigned long long* status) {
    for (int j = 0; j < n_vec; ++j) {
        const uint4 v = p[32 * j];
        const int idx = first_invalid16(v);
        if (idx < 16) {
            const uint32_t w = idx < 4 ? v.x : idx < 8 ? v.y : idx < 12 ? v.z : v.w;
            report_invalid(status, (vec_index + 32ull * j) * 16ull + idx, w >> (8 * (idx & 3)));
            return;
        }
    }
}

template <int U>
__device__ __forceinline__ void encode_tile(const uint4* __restrict__ in, uint32_t* __restrict__ out,
                                            unsigned long long vec0, unsigned lane,
                                            unsigned long long* status) {
    const uint4* p = in + vec0 + lane;
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = ld_stream_v4(p + 32 * j);
    uint32_t bad = 0, r[U];
#pragma unroll
    for (int j = 0; j < U; ++j) r[j] = pack16(v[j], bad);
    uint32_t* q = out + vec0 + lane;
#pragma unroll
    for (int j = 0; j < U; ++j) st_stream_u32(q + 32 * j, r[j]);
    if (bad & kValidMask) report_first_invalid(p, U, vec0 + lane, status);
}

// n_vec full 16-byte vectors, then `tail` (< 16) trailing bytes; out32 holds 2*ceil(n/32) words.
template <int U>
__global__ void __launch_bounds__(kThreads)
encode_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_vec,
              unsigned tail, unsigned long long total32, unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long warp = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = n_vec / kTile;

    for (unsigned long long t = warp; t < n_tiles; t += n_warps) encode_tile<U>(in, out, t * kTile, lane, status);

    // ragged end: the vectors after the last full tile, one vector per lane per round
    if (warp == n_tiles % n_warps) {
        for (unsigned long long i = n_tiles * kTile + lane; i < n_vec; i += 32) {
            const uint4 v = ld_stream_v4(in + i);
            uint32_t bad = 0;
            out[i] = pack16(v, bad);
            if (bad & kValidMask) report_first_invalid(in + i, 1, i, status);
        }
    }
    // trailing bytes (< 16) and the zero padding of the last 64-bit word
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(in + n_vec);
        unsigned long long w = n_vec;
        if (tail) {
            uint32_t packed = 0;
            for (unsigned i = 0; i < tail; ++i) {
                const uint32_t b = bytes[i];
                if (!byte_is_valid(b)) {
                    report_invalid(status, n_vec * 16ull + i, b);
                    break;
                }
                packed |= (((b >> 1) ^ (b >> 2)) & 3u) << (2 * i);
            }
            out[w++] = packed;
        }
        for (; w < total32; ++w) out[w] = 0;
    }
}

// Fallback for input pointers that are not 16-byte aligned: one thread per 16 bases, byte loads.
__global__ void __launch_bounds__(kThreads)
encode_unaligned_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n,
                        unsigned long long total32, unsigned long long* __restrict__ status) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total32; w += stride) {
        uint32_t packed = 0;
        const unsigned long long b0 = w * 16ull;
        for (unsigned i = 0; i < 16 && b0 + i < n; ++i) {
            const uint32_t b = in[b0 + i];
            if (!byte_is_valid(b)) {
                report_invalid(status, b0 + i, b);
                break;
            }
            packed |= (((b >> 1) ^ (b >> 2)) & 3u) << (2 * i);
        }
        out[w] = packed;
    }
}

// ============================================================================ decode =========
// One packed byte (4 bases) indexes a 256-entry table of 4-ASCII-byte words held in shared memory.
// The table is replicated once per lane (entry e of lane l at word e*32 + l) so that the 32 lanes
// of a warp always hit 32 different banks whatever the data: every LDS is conflict-free.

constexpr int kLutWords = 256 * 32;

__device__ __forceinline__ void decode_lut_init(uint32_t* lut) {
    for (int i = threadIdx.x; i < kLutWords; i += blockDim.x) lut[i] = ascii4_of_byte((uint32_t)i >> 5);
    __syncthreads();
}

__device__ __forceinline__ uint4 decode16(uint32_t w, const uint32_t* lut_lane) {
    uint4 o;
    o.x = lut_lane[(w & 0xFFu) << 5];
    o.y = lut_lane[((w >> 8) & 0xFFu) << 5];
    o.z = lut_lane[((w >> 16) & 0xFFu) << 5];
    o.w = lut_lane[(w >> 24) << 5];
    return o;
}

// n_w32 full 32-bit words (16 bases each), then `tail` (< 16) bases from one more word.
template <int U>
__global__ void __launch_bounds__(kThreads)
decode_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, unsigned long long n_w32, unsigned tail) {
    __shared__ uint32_t lut[kLutWords];
    decode_lut_init(lut);
    const unsigned lane = threadIdx.x & 31;
    const uint32_t* lut_lane = lut + lane;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long warp = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = n_w32 / kTile;

    for (unsigned long long t = warp; t < n_tiles; t += n_warps) {
        const uint32_t* p = in + t * kTile + lane;
        uint32_t w[U];
#pragma unroll
        for (int j = 0; j < U; ++j) w[j] = ld_stream_u32(p + 32 * j);
        uint4* q = out + t * kTile + lane;
#pragma unroll
        for (int j = 0; j < U; ++j) st_stream_v4(q + 32 * j, decode16(w[j], lut_lane));
    }
    if (warp == n_tiles % n_warps) {
        for (unsigned long long i = n_tiles * kTile + lane; i < n_w32; i += 32)
            st_stream_v4(out + i, decode16(ld_stream_u32(in + i), lut_lane));
    }
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t w = in[n_w32];
        uint8_t* o = reinterpret_cast<uint8_t*>(out + n_w32);
        for (unsigned i = 0; i < tail; ++i) o[i] = (uint8_t)(0x54474341u >> (8 * ((w >> (2 * i)) & 3u)));
    }
}

// Fallback for output pointers that are not 16-byte aligned: one thread per 4 bases, byte stores.
__global__ void __launch_bounds__(kThreads)
decode_unaligned_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, unsigned long long n_bases) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_groups = (n_bases + 3) / 4;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint32_t a = ascii4_of_byte(in[g]);
        for (unsigned i = 0; i < 4 && g * 4 + i < n_bases; ++i) out[g * 4 + i] = (uint8_t)(a >> (8 * i));
    }
}

// ============================================================================ launchers ======

constexpr int kEncodeU = 4;
constexpr int kDecodeU = 4;

cudaError_t launch_encode(const DeviceInfo& di, const uint8_t* d_seq, size_t n, uint64_t* d_out,
                          unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || n == 0) return e;
    const unsigned long long total32 = 2ull * ((n + 31) / 32);
    uint32_t* out32 = reinterpret_cast<uint32_t*>(d_out);
    if (reinterpret_cast<uintptr_t>(d_seq) & 15u) {
        static const int resident = resident_blocks(encode_unaligned_kernel, kThreads, di);
        encode_unaligned_kernel<<<grid_for((total32 + kThreads - 1) / kThreads, resident), kThreads, 0, s>>>(
            d_seq, out32, n, total32, d_status);
        return cudaGetLastError();
    }
    static const int resident = resident_blocks(encode_kernel<kEncodeU>, kThreads, di);
    const unsigned long long n_vec = n / 16;
    const unsigned long long tiles = n_vec / (32 * kEncodeU) + 1;
    encode_kernel<kEncodeU><<<grid_for((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock, resident), kThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(d_seq), out32, n_vec, (unsigned)(n % 16), total32, d_status);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DeviceInfo& di, const uint64_t* d_words, size_t n_bases, uint8_t* d_out,
                          cudaStream_t s) {
    if (n_bases == 0) return cudaSuccess;
    if (reinterpret_cast<uintptr_t>(d_out) & 15u) {
        static const int resident = resident_blocks(decode_unaligned_kernel, kThreads, di);
        const unsigned long long groups = (n_bases + 3) / 4;
        decode_unaligned_kernel<<<grid_for((groups + kThreads - 1) / kThreads, resident), kThreads, 0, s>>>(
            reinterpret_cast<const uint8_t*>(d_words), d_out, n_bases);
        return cudaGetLastError();
    }
    static const int resident = resident_blocks(decode_kernel<kDecodeU>, kThreads, di);
    const unsigned long long n_w32 = n_bases / 16;
    const unsigned long long tiles = n_w32 / (32 * kDecodeU) + 1;
    decode_kernel<kDecodeU><<<grid_for((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock, resident), kThreads, 0, s>>>(
        reinterpret_cast<const uint32_t*>(d_words), reinterpret_cast<uint4*>(d_out), n_w32, (unsigned)(n_bases % 16));
    return cudaGetLastError();
}

}  // namespace bn
