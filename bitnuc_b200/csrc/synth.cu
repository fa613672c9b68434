// synth.cu -- counter-based synthetic nucleotide streams generated on the device (sm_100a).
//
// The reference's tests draw "nucgen-style" input: i.i.d. uniform upper-case A/C/G/T from an
// unseeded RNG (/root/reference/src/utils/mod.rs:114-121).  Here word j of stream s is
// splitmix64((seed ^ s*golden) + j) and base 32j+i is "ACGT"[(W >> 2i) & 3], so the
// CPU-side checker, any GPU and any shard produce identical bytes without a transfer, and the expected
// encode of a generated stream is the word stream itself.
#include "common.cuh"
#include "launch.cuh"

namespace bn {

__device__ __forceinline__ unsigned long long synth_word(unsigned long long base, unsigned long long j) {
    return splitmix64(base + j);
}

__global__ void __launch_bounds__(kThreads)
synth_words_kernel(unsigned long long base, unsigned long long first_word, unsigned long long n_words, uint64_t* __restrict__ out) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += step)
        out[i] = synth_word(base, first_word + i);
}

// one thread per 32-base word; full words are written as two 128-bit stores when `out` is aligned
__global__ void __launch_bounds__(kThreads)
synth_ascii_kernel(unsigned long long base, unsigned long long first_word, unsigned long long n, uint8_t* __restrict__ out,
                   int aligned) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_words = ceil_div(n, 32);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += step) {
        const unsigned long long w = synth_word(base, first_word + i);
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = ascii4_of_byte((uint32_t)(w >> (8 * k)) & 0xFFu);
        if (aligned && i * 32 + 32 <= n) {
            uint4* o = reinterpret_cast<uint4*>(out + i * 32);
            o[0] = make_uint4(a[0], a[1], a[2], a[3]);
            o[1] = make_uint4(a[4], a[5], a[6], a[7]);
        } else {
            for (unsigned b = 0; b < 32 && i * 32 + b < n; ++b) out[i * 32 + b] = (uint8_t)(a[b >> 2] >> (8 * (b & 3)));
        }
    }
}

static unsigned long long stream_base(uint64_t seed, uint64_t stream_id) {
    return seed ^ (stream_id * 0x9E3779B97F4A7C15ull);
}

cudaError_t launch_synth_words(const DeviceInfo& di, uint64_t seed, uint64_t stream_id, uint64_t first_word,
                               size_t n_words, uint64_t* d_out, cudaStream_t s) {
    if (n_words == 0) return cudaSuccess;
    static const int per_sm = blocks_per_sm(synth_words_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
    synth_words_kernel<<<grid_for(ceil_div(n_words, kThreads), resident), kThreads, 0, s>>>(stream_base(seed, stream_id),
                                                                                             first_word, n_words, d_out);
    return cudaGetLastError();
}

cudaError_t launch_synth_ascii(const DeviceInfo& di, uint64_t seed, uint64_t stream_id, uint64_t first_base,
                               size_t n, uint8_t* d_out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    static const int per_sm = blocks_per_sm(synth_ascii_kernel, kThreads);
    const int resident = per_sm * di.sm_count;
    const int aligned = (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    synth_ascii_kernel<<<grid_for(ceil_div(ceil_div(n, 32), kThreads), resident), kThreads, 0, s>>>(
        stream_base(seed, stream_id), first_base / 32, n, d_out, aligned);
    return cudaGetLastError();
}

}  // namespace bn
