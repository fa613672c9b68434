// codec.cu -- launchers of the ASCII <-> 2-bit streaming codec kernels (sm_100a).
//
// The kernels live in codec_kernels.cuh (shared with tools/tune_codec.cu).  Both are HBM-bound at
// 1.25 algorithmic bytes per base (encode 1 B read + 0.25 B write, decode the reverse).
//
// Configuration chosen from the on-device sweeps in profiles/r01_tune_codec_sweep{1,2}.txt:
//   * one CTA per chunk of tiles handed out by the hardware CTA scheduler (dynamic balance between
//     the two dies) instead of a persistent statically partitioned grid: +13 % on both kernels;
//   * 512 threads, U = 4 vectors/words in flight per thread, one tile per warp;
//   * loads through the read-only path with L1 allocation, stores with the streaming (.cs) policy;
//   * decode without any table in memory: nibble spread + PRMT as a 4-entry byte LUT (no shared
//     memory, no per-CTA table initialisation, full occupancy).
#include "codec_kernels.cuh"
#include "launch.cuh"

namespace bn {

constexpr int kCodecU = 4;
constexpr int kCodecThreads = 512;
constexpr int kCodecT = 1;

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

cudaError_t launch_encode(const DeviceInfo& di, const uint8_t* d_seq, size_t n, uint64_t* d_out,
                          unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || n == 0) return e;
    const unsigned long long total32 = 2ull * ((n + 31) / 32);
    uint32_t* out32 = reinterpret_cast<uint32_t*>(d_out);
    if (reinterpret_cast<uintptr_t>(d_seq) & 15u) {
        static const int per_sm = blocks_per_sm(encode_unaligned_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        encode_unaligned_kernel<<<grid_for((total32 + kThreads - 1) / kThreads, resident), kThreads, 0, s>>>(
            d_seq, out32, n, total32, d_status);
        return cudaGetLastError();
    }
    const unsigned long long n_vec = n / 16;
    const unsigned long long ctas = TileWalk<kCodecThreads, 1, kCodecT>::ctas(n_vec / (32 * kCodecU));
    encode_kernel<kCodecU, kCodecThreads, 1, kCodecT, LD_PLAIN, ST_CS><<<(unsigned)(ctas ? ctas : 1), kCodecThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(d_seq), out32, n_vec, (unsigned)(n % 16), total32, d_status);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DeviceInfo& di, const uint64_t* d_words, size_t n_bases, uint8_t* d_out,
                          cudaStream_t s) {
    if (n_bases == 0) return cudaSuccess;
    if (reinterpret_cast<uintptr_t>(d_out) & 15u) {
        static const int per_sm = blocks_per_sm(decode_unaligned_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
        const unsigned long long groups = (n_bases + 3) / 4;
        decode_unaligned_kernel<<<grid_for((groups + kThreads - 1) / kThreads, resident), kThreads, 0, s>>>(
            reinterpret_cast<const uint8_t*>(d_words), d_out, n_bases);
        return cudaGetLastError();
    }
    const unsigned long long n_w32 = n_bases / 16;
    const unsigned long long ctas = TileWalk<kCodecThreads, 1, kCodecT>::ctas(n_w32 / (32 * kCodecU));
    decode_kernel<kCodecU, kCodecThreads, 1, kCodecT, LD_PLAIN, ST_CS, 2><<<(unsigned)(ctas ? ctas : 1), kCodecThreads, 0, s>>>(
        reinterpret_cast<const uint32_t*>(d_words), reinterpret_cast<uint4*>(d_out), n_w32, (unsigned)(n_bases % 16));
    return cudaGetLastError();
}

}  // namespace bn
