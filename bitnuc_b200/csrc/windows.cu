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

cudaError_t launch_kmer_windows(const DeviceInfo&, const uint8_t* d_seq, size_t n, uint32_t k, uint64_t* d_out,
                                unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || k == 0 || n < k) return e;
    const unsigned long long n_win = n - k + 1;
    kmer_windows_kernel<<<(unsigned)ceil_div(n_win, kWinTile), kWinThreads, 0, s>>>(d_seq, n, k, d_out, d_status);
    return cudaGetLastError();
}

}  // namespace bn
