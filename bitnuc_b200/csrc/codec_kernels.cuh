// codec_kernels.cuh -- the ASCII <-> 2-bit streaming codec kernels, parametrised so that the
// production build (codec.cu) and the tuning harness (tools/tune_codec.cu) compile the same code.
//
// Replaces the reference's per-ISA kernels behind encode/decode:
//   encode: /root/reference/src/utils/packing/avx.rs:76-151 (naive.rs:4-43 defines the results)
//   decode: /root/reference/src/utils/unpacking/avx.rs:26-33,117-153 (naive.rs:3-25)
//
// Data layout: the ASCII stream is viewed as 16-byte vectors, the packed stream as 32-bit words;
// vector i <-> word i, so lane l of a warp always touches element tile_base + 32*j + l and every
// global access is a fully coalesced 512-byte (vector) or 128-byte (word) warp transaction.
// A "tile" is 32*U vectors (U loads in flight per thread, all issued before the first use).
//
// Template knobs
//   U        vectors / words per thread per tile
//   THREADS  CTA size
//   SCHED    0: persistent grid, tiles strided over all warps (static partition)
//            1: one CTA per chunk of (THREADS/32)*T tiles, handed out by the hardware CTA scheduler
//               (dynamic load balance across the two dies)
//   T        tiles per warp when SCHED == 1
//   LP / SP  load / store cache policy
//   DEC      decode flavour: 0 per-lane replicated 256-entry LUT in shared memory (conflict-free),
//            1 single 1 KB LUT, 2 register-only (nibble spread + PRMT as a 4-entry byte LUT)
#pragma once

#include "common.cuh"

namespace bn {

// ============================================================================ encode =========

// Rare path: re-read this thread's vectors (address order = j order) and report the first invalid byte.
static __device__ __noinline__ void report_first_invalid(const uint4* p, int n_vec, unsigned long long vec_index,
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

// n_vec full 16-byte vectors, then `tail` (< 16) trailing bytes; out32 holds total32 = 2*ceil(n/32) words.
template <int U, int THREADS, int SCHED, int T, int LP, int SP>
__global__ void __launch_bounds__(THREADS)
encode_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_vec, unsigned tail,
              unsigned long long total32, unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = n_vec / kTile;
    const TileWalk<THREADS, SCHED, T> walk(n_tiles);

    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long vec0 = t * kTile;
        const uint4* p = in + vec0 + lane;
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld128<LP>(p + 32 * j);
        uint32_t bad = 0, r[U];
#pragma unroll
        for (int j = 0; j < U; ++j) r[j] = pack16(v[j], bad);
        uint32_t* q = out + vec0 + lane;
#pragma unroll
        for (int j = 0; j < U; ++j) st32<SP>(q + 32 * j, r[j]);
        if (bad & kValidMask) report_first_invalid(p, U, vec0 + lane, status);
    }

    // ragged end (last CTA): the vectors after the last full tile, the trailing bytes (< 16) and the
    // zero padding of the last 64-bit word
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_vec; i += THREADS) {
            const uint4 v = ld128<LP>(in + i);
            uint32_t bad = 0;
            out[i] = pack16(v, bad);
            if (bad & kValidMask) report_first_invalid(in + i, 1, i, status);
        }
        if (threadIdx.x == 0) {
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
}

// ============================================================================ decode =========

constexpr int kLutWords = 256 * 32;

template <int DEC>
__device__ __forceinline__ uint4 decode16(uint32_t w, const uint32_t* lut_lane) {
    if (DEC == 2) return decode16_prmt(w);
    constexpr int kShift = DEC == 0 ? 5 : 0;  // replicated: entry e of lane l at word e*32 + l
    return make_uint4(lut_lane[(w & 0xFFu) << kShift], lut_lane[((w >> 8) & 0xFFu) << kShift],
                      lut_lane[((w >> 16) & 0xFFu) << kShift], lut_lane[(w >> 24) << kShift]);
}

// n_w32 full 32-bit words (16 bases each), then `tail` (< 16) bases from one more word.
template <int U, int THREADS, int SCHED, int T, int LP, int SP, int DEC>
__global__ void __launch_bounds__(THREADS)
decode_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, unsigned long long n_w32, unsigned tail) {
    __shared__ uint32_t lut[DEC == 0 ? kLutWords : DEC == 1 ? 256 : 1];
    const unsigned lane = threadIdx.x & 31;
    if (DEC == 0) {
        for (int i = threadIdx.x; i < kLutWords; i += THREADS) lut[i] = ascii4_of_byte((uint32_t)i >> 5);
        __syncthreads();
    } else if (DEC == 1) {
        for (int i = threadIdx.x; i < 256; i += THREADS) lut[i] = ascii4_of_byte((uint32_t)i);
        __syncthreads();
    }
    const uint32_t* lut_lane = DEC == 0 ? lut + lane : lut;
    constexpr unsigned kTile = 32 * U;
    const unsigned long long n_tiles = n_w32 / kTile;
    const TileWalk<THREADS, SCHED, T> walk(n_tiles);

    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const uint32_t* p = in + t * kTile + lane;
        uint32_t w[U];
#pragma unroll
        for (int j = 0; j < U; ++j) w[j] = ld32<LP>(p + 32 * j);
        uint4* q = out + t * kTile + lane;
#pragma unroll
        for (int j = 0; j < U; ++j) st128<SP>(q + 32 * j, decode16<DEC>(w[j], lut_lane));
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (unsigned long long i = n_tiles * kTile + threadIdx.x; i < n_w32; i += THREADS)
            st128<SP>(out + i, decode16<DEC>(ld32<LP>(in + i), lut_lane));
        if (tail && threadIdx.x == 0) {
            const uint32_t w = in[n_w32];
            uint8_t* o = reinterpret_cast<uint8_t*>(out + n_w32);
            for (unsigned i = 0; i < tail; ++i) o[i] = (uint8_t)(0x54474341u >> (8 * ((w >> (2 * i)) & 3u)));
        }
    }
}

// ============================================================================ 256-bit variants ==
// Same work with one 64-bit packed word <-> 32 ASCII bytes per lane step: 256-bit loads (encode) / stores (decode),
// i.e. one whole 32-byte sector per lane on the ASCII side.  The ASCII buffer must be 32-byte aligned.
// The ragged end (words after the last full tile, trailing bases, zero padding) is left to the 128-bit kernels'
// tail code: these kernels only take n_w64 full words.
// Measured 2 % SLOWER than the 128-bit kernels on B200 (profiles/r01_tune_codec_sweep3.txt: encode 6930 vs 7085 GB/s,
// decode 6457 vs 6598 GB/s), so the production launchers in codec.cu do not use them; they stay here for the
// tuning harness.

template <int U, int THREADS, int T, int LP, int SP>
__global__ void __launch_bounds__(THREADS)
encode256_kernel(const uint8_t* __restrict__ in, uint2* __restrict__ out, unsigned long long n_tiles,
                 unsigned long long* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * U;   // 64-bit words per tile
    const TileWalk<THREADS, 1, T> walk(n_tiles);
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long w0 = t * kTile + lane;
        const uint8_t* p = in + 32ull * w0;
        uint8x v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld256<LP>(p + 1024ull * j);
        uint32_t bad = 0;
        uint2 r[U];
#pragma unroll
        for (int j = 0; j < U; ++j) r[j] = make_uint2(pack16(v[j].lo, bad), pack16(v[j].hi, bad));
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (SP == ST_CS) st_stream_v2(out + w0 + 32 * j, r[j]);
            else out[w0 + 32 * j] = r[j];
        }
        if (bad & kValidMask) {
#pragma unroll 1
            for (int j = 0; j < U; ++j)  // vectors 2*(w0 + 32 j) and the one after it
                report_first_invalid(reinterpret_cast<const uint4*>(p + 1024ull * j), 1, 2 * (w0 + 32ull * j), status),
                report_first_invalid(reinterpret_cast<const uint4*>(p + 1024ull * j) + 1, 1, 2 * (w0 + 32ull * j) + 1, status);
        }
    }
}

template <int U, int THREADS, int T, int LP, int SP>
__global__ void __launch_bounds__(THREADS)
decode256_kernel(const uint2* __restrict__ in, uint8_t* __restrict__ out, unsigned long long n_tiles) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kTile = 32 * U;   // 64-bit words per tile
    const TileWalk<THREADS, 1, T> walk(n_tiles);
    for (unsigned long long t = walk.first; t < walk.end; t += walk.step) {
        const unsigned long long w0 = t * kTile + lane;
        uint2 w[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (LP == LD_NC_NOALLOC) w[j] = ld_stream_v2(in + w0 + 32 * j);
            else w[j] = __ldg(in + w0 + 32 * j);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) st256<SP>(out + 32ull * (w0 + 32 * j), decode16_prmt(w[j].x), decode16_prmt(w[j].y));
    }
}

}  // namespace bn
