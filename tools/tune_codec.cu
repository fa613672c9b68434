// tune_codec.cu -- on-GPU sweep of the codec kernel variants (development tool, not shipped).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I bitnuc_b200/csrc tools/tune_codec.cu -o tools/tune_codec
// Prints, per variant: isolated encode / decode time (same kernel back to back) and the time of the
// alternating encode->decode step that bench.py measures, all with CUDA events, plus GB/s at the
// algorithmic 1.25 B/base.  Every variant's output is checked against the first one.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "codec_kernels.cuh"

using namespace bn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void gen_kernel(uint8_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = (uint8_t)(0x54474341u >> (8 * (splitmix64(i >> 5) >> (2 * (i & 31)) & 3)));
}
__global__ void diff_kernel(const uint4* a, const uint4* b, size_t n, unsigned long long* count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (; i < n; i += step) { uint4 x = a[i], y = b[i]; c += (x.x != y.x) + (x.y != y.y) + (x.z != y.z) + (x.w != y.w); }
    if (c) atomicAdd(count, c);
}

struct Variant {
    std::string name;
    std::function<void(const uint8_t*, size_t, uint64_t*, unsigned long long*, cudaStream_t)> enc;
    std::function<void(const uint64_t*, size_t, uint8_t*, cudaStream_t)> dec;
};

static int g_sms = 148;

template <int U, int THREADS, int SCHED, int T, int LP, int SP>
void run_encode(const uint8_t* in, size_t n, uint64_t* out, unsigned long long* status, cudaStream_t s) {
    auto k = encode_kernel<U, THREADS, SCHED, T, LP, SP>;
    const unsigned long long n_vec = n / 16, n_tiles = n_vec / (32 * U);
    unsigned grid;
    if (SCHED == 0) {
        static int per_sm = 0;
        if (!per_sm) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, 0));
        grid = per_sm * g_sms;
    } else {
        grid = (unsigned)TileWalk<THREADS, SCHED, T>::ctas(n_tiles);
        if (grid == 0) grid = 1;
    }
    cudaMemsetAsync(status, 0xFF, 8, s);
    k<<<grid, THREADS, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint32_t*>(out), n_vec, (unsigned)(n % 16),
                               2ull * ((n + 31) / 32), status);
}
template <int U, int THREADS, int SCHED, int T, int LP, int SP, int DEC>
void run_decode(const uint64_t* in, size_t n, uint8_t* out, cudaStream_t s) {
    auto k = decode_kernel<U, THREADS, SCHED, T, LP, SP, DEC>;
    const unsigned long long n_w32 = n / 16, n_tiles = n_w32 / (32 * U);
    unsigned grid;
    if (SCHED == 0) {
        static int per_sm = 0;
        if (!per_sm) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, 0));
        grid = per_sm * g_sms;
    } else {
        grid = (unsigned)TileWalk<THREADS, SCHED, T>::ctas(n_tiles);
        if (grid == 0) grid = 1;
    }
    k<<<grid, THREADS, 0, s>>>(reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint4*>(out), n_w32, (unsigned)(n % 16));
}

// 256-bit variants: full tiles by the wide kernel, the ragged rest by the 128-bit kernel on the remaining range
template <int U, int THREADS, int T, int LP, int SP>
void run_encode256(const uint8_t* in, size_t n, uint64_t* out, unsigned long long* status, cudaStream_t s) {
    const unsigned long long n_tiles = (n / 32) / (32 * U);
    const size_t done = (size_t)n_tiles * 32 * U * 32;   // bases handled by the wide kernel
    cudaMemsetAsync(status, 0xFF, 8, s);
    if (n_tiles)
        encode256_kernel<U, THREADS, T, LP, SP><<<(unsigned)TileWalk<THREADS, 1, T>::ctas(n_tiles), THREADS, 0, s>>>(
            in, reinterpret_cast<uint2*>(out), n_tiles, status);
    if (n > done) {
        const size_t rest = n - done;
        encode_kernel<4, 512, 1, 1, LD_PLAIN, ST_CS><<<(unsigned)std::max<unsigned long long>(1, TileWalk<512, 1, 1>::ctas((rest / 16) / 128)), 512, 0, s>>>(
            reinterpret_cast<const uint4*>(in + done), reinterpret_cast<uint32_t*>(out + done / 32), rest / 16, (unsigned)(rest % 16),
            2ull * ((rest + 31) / 32), status);
    }
}
template <int U, int THREADS, int T, int LP, int SP>
void run_decode256(const uint64_t* in, size_t n, uint8_t* out, cudaStream_t s) {
    const unsigned long long n_tiles = (n / 32) / (32 * U);
    const size_t done = (size_t)n_tiles * 32 * U * 32;
    if (n_tiles)
        decode256_kernel<U, THREADS, T, LP, SP><<<(unsigned)TileWalk<THREADS, 1, T>::ctas(n_tiles), THREADS, 0, s>>>(
            reinterpret_cast<const uint2*>(in), out, n_tiles);
    if (n > done) {
        const size_t rest = n - done;
        decode_kernel<4, 512, 1, 1, LD_PLAIN, ST_CS, 2><<<(unsigned)std::max<unsigned long long>(1, TileWalk<512, 1, 1>::ctas((rest / 16) / 128)), 512, 0, s>>>(
            reinterpret_cast<const uint32_t*>(in + done / 32), reinterpret_cast<uint4*>(out + done), rest / 16, (unsigned)(rest % 16));
    }
}
#define W(name, U, TH, T, LP, SP) Variant{name, run_encode256<U, TH, T, LP, SP>, run_decode256<U, TH, T, LP, SP>}

#define V(name, U, TH, SCHED, T, LP, SP, DEC) \
    Variant{name, run_encode<U, TH, SCHED, T, LP, SP>, run_decode<U, TH, SCHED, T, LP, SP, DEC>}

int main(int argc, char** argv) {
    size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000000ull;
    int reps = argc > 2 ? atoi(argv[2]) : 20;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, n = %zu bases, reps = %d\n", prop.name, g_sms, n, reps);
    uint8_t *asc, *back, *back_ref;
    uint64_t *words, *words_ref;
    unsigned long long *status, *count;
    const size_t nw = (n + 31) / 32;
    CK(cudaMalloc(&asc, n + 64)); CK(cudaMalloc(&back, n + 64)); CK(cudaMalloc(&back_ref, n + 64));
    CK(cudaMalloc(&words, nw * 8 + 64)); CK(cudaMalloc(&words_ref, nw * 8 + 64));
    CK(cudaMalloc(&status, 8)); CK(cudaMalloc(&count, 8));
    gen_kernel<<<g_sms * 8, 256>>>(asc, n);
    CK(cudaDeviceSynchronize());

    std::vector<Variant> vs = {
        // name                                  U  TH  SCHED T  LP             SP       DEC
        V("prmt   U4 t512 cta T1  plain/cs  ", 4, 512, 1, 1, LD_PLAIN, ST_CS, 2),
        W("wide256 U2 t512 T1 plain/cs      ", 2, 512, 1, LD_PLAIN, ST_CS),
        W("wide256 U2 t256 T1 plain/cs      ", 2, 256, 1, LD_PLAIN, ST_CS),
        W("wide256 U4 t256 T1 plain/cs      ", 4, 256, 1, LD_PLAIN, ST_CS),
        W("wide256 U4 t512 T1 plain/cs      ", 4, 512, 1, LD_PLAIN, ST_CS),
        W("wide256 U1 t512 T1 plain/cs      ", 1, 512, 1, LD_PLAIN, ST_CS),
        W("wide256 U1 t1024 T1 plain/cs     ", 1, 1024, 1, LD_PLAIN, ST_CS),
        W("wide256 U2 t512 T1 nc/cs         ", 2, 512, 1, LD_NC_NOALLOC, ST_CS),
        W("wide256 U2 t512 T1 plain/plain   ", 2, 512, 1, LD_PLAIN, ST_PLAIN),
        W("wide256 U2 t512 T1 plain/noalloc ", 2, 512, 1, LD_PLAIN, ST_NOALLOC),
        W("wide256 U2 t512 T2 plain/cs      ", 2, 512, 2, LD_PLAIN, ST_CS),
        W("wide256 U2 t128 T1 plain/cs      ", 2, 128, 1, LD_PLAIN, ST_CS),
    };

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_ms = [&](std::function<void()> f) {
        for (int i = 0; i < 3; ++i) f();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms / reps;
    };
    const double gb = 1.25 * (double)n / 1e9;
    printf("%-36s %9s %8s %9s %8s %9s %8s  %s\n", "variant", "enc ms", "GB/s", "dec ms", "GB/s", "step ms", "GB/s", "check");
    for (size_t i = 0; i < vs.size(); ++i) {
        auto& v = vs[i];
        uint64_t* w = i == 0 ? words_ref : words;
        uint8_t* b = i == 0 ? back_ref : back;
        CK(cudaMemset(w, 0xA5, nw * 8)); CK(cudaMemset(b, 0, n));
        float enc = time_ms([&] { v.enc(asc, n, w, status, 0); });
        float dec = time_ms([&] { v.dec(w, n, b, 0); });
        float step = time_ms([&] { v.enc(asc, n, w, status, 0); v.dec(w, n, b, 0); });
        CK(cudaGetLastError());
        unsigned long long bad = 0, st = 0;
        if (i > 0) {
            CK(cudaMemset(count, 0, 8));
            diff_kernel<<<g_sms * 8, 256>>>((const uint4*)words_ref, (const uint4*)words, nw * 8 / 16, count);
            diff_kernel<<<g_sms * 8, 256>>>((const uint4*)back_ref, (const uint4*)back, n / 16, count);
            CK(cudaMemcpy(&bad, count, 8, cudaMemcpyDeviceToHost));
        } else {
            CK(cudaMemset(count, 0, 8));
            diff_kernel<<<g_sms * 8, 256>>>((const uint4*)asc, (const uint4*)back_ref, n / 16, count);
            CK(cudaMemcpy(&bad, count, 8, cudaMemcpyDeviceToHost));
        }
        CK(cudaMemcpy(&st, status, 8, cudaMemcpyDeviceToHost));
        printf("%-36s %9.4f %8.0f %9.4f %8.0f %9.4f %8.0f  %s\n", v.name.c_str(), enc, gb / enc * 1e3, dec, gb / dec * 1e3, step,
               2 * gb / step * 1e3, (bad == 0 && st == ~0ull) ? "ok" : "MISMATCH");
        fflush(stdout);
    }
    return 0;
}
