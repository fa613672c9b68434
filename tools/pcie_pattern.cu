// pcie_pattern.cu -- what the PCIe link gives for the COPY PATTERN of the pipelined end-to-end leg, without any kernel:
// thread A moves (big up, small down) chunks like bn_encode, thread B (small up, big down) chunks like bn_decode, each
// through `stages` rotating streams (a chunk's download follows its upload on the same stream, a stage is reused once
// its chunk has completed).  Variants: chunk size, pipeline depth, and whether the small copies ride their own stream
// ahead of the big ones.  Prints GB/s moved in each direction: the denominator for bench.py's e2e at N = 1.
//   nvcc -O2 -o tools/pcie_pattern tools/pcie_pattern.cu && tools/pcie_pattern
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e = (x);                                                              \
        if (e != cudaSuccess) {                                                           \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            std::exit(1);                                                                 \
        }                                                                                 \
    } while (0)

struct Side {
    size_t up_total, down_total;   // bytes per step
    char *h_up, *h_down, *d_up, *d_down;
};

// one call: n_chunks chunks, chunk c = (up_total / n_chunks) up then (down_total / n_chunks) down on stream c % stages
static void run_call(const Side& s, int n_chunks, int stages, cudaStream_t* st, cudaEvent_t* ev) {
    const size_t up = s.up_total / n_chunks, down = s.down_total / n_chunks;
    for (int c = 0; c < n_chunks; ++c) {
        const int k = c % stages;
        if (c >= stages) CK(cudaEventSynchronize(ev[k]));
        CK(cudaMemcpyAsync(s.d_up + (size_t)k * up, s.h_up + (size_t)c * up, up, cudaMemcpyHostToDevice, st[k]));
        CK(cudaMemcpyAsync(s.h_down + (size_t)c * down, s.d_down + (size_t)k * down, down, cudaMemcpyDeviceToHost, st[k]));
        CK(cudaEventRecord(ev[k], st[k]));
    }
    for (int k = 0; k < stages && k < n_chunks; ++k) CK(cudaEventSynchronize(ev[k]));
}

// variant 2: uploads on one stream, downloads on another (a download waits for its chunk's upload through an event); the
// SMALL direction runs `deep` chunks ahead of / behind the big one (its ring of stage buffers is cheap), the big direction
// keeps `stages` buffers -- so a small copy queued behind the other call's big copies has `deep` chunk-times to get through
static void run_call_split(const Side& s, int n_chunks, int stages, int deep, cudaStream_t up_st, cudaStream_t down_st, cudaEvent_t* up_ev,
                           cudaEvent_t* down_ev) {
    const size_t up = s.up_total / n_chunks, down = s.down_total / n_chunks;
    const bool up_small = s.up_total < s.down_total;
    const int up_ring = up_small ? deep : stages, down_ring = up_small ? stages : deep;
    int issued_up = 0, issued_down = 0;
    while (issued_down < n_chunks) {
        // uploads run ahead as far as their ring allows: upload c reuses the buffer of chunk c - up_ring, whose download was issued
        // on down_st before (the "kernel" that consumed it precedes that download), so it waits for that download's upload event chain
        while (issued_up < n_chunks && issued_up < issued_down + up_ring) {
            const int c = issued_up, k = c % up_ring;
            CK(cudaMemcpyAsync(s.d_up + (size_t)k * up, s.h_up + (size_t)c * up, up, cudaMemcpyHostToDevice, up_st));
            CK(cudaEventRecord(up_ev[c % 64], up_st));
            ++issued_up;
        }
        const int c = issued_down, k = c % down_ring;
        if (c >= down_ring) CK(cudaEventSynchronize(down_ev[k]));          // the out buffer's previous download has completed
        CK(cudaStreamWaitEvent(down_st, up_ev[c % 64], 0));
        CK(cudaMemcpyAsync(s.h_down + (size_t)c * down, s.d_down + (size_t)k * down, down, cudaMemcpyDeviceToHost, down_st));
        CK(cudaEventRecord(down_ev[k], down_st));
        ++issued_down;
    }
    CK(cudaStreamSynchronize(down_st));
}

int main() {
    const size_t big = 1000000000, small = 250000000;
    Side enc{big, small, nullptr, nullptr, nullptr, nullptr}, dec{small, big, nullptr, nullptr, nullptr, nullptr};
    for (Side* s : {&enc, &dec}) {
        CK(cudaHostAlloc(&s->h_up, s->up_total, cudaHostAllocPortable));
        CK(cudaHostAlloc(&s->h_down, s->down_total, cudaHostAllocPortable));
        CK(cudaMalloc(&s->d_up, s->up_total));
        CK(cudaMalloc(&s->d_down, s->down_total));
        memset(s->h_up, 1, s->up_total);
        memset(s->h_down, 1, s->down_total);
    }
    const int kMaxStages = 16;
    cudaStream_t sa[kMaxStages], sb[kMaxStages];
    cudaEvent_t ea[kMaxStages], eb[kMaxStages];
    for (int i = 0; i < kMaxStages; ++i) {
        CK(cudaStreamCreateWithFlags(&sa[i], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&sb[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ea[i], cudaEventDisableTiming | cudaEventBlockingSync));
        CK(cudaEventCreateWithFlags(&eb[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    const int steps = 6;
    {
        cudaEvent_t ua[64], ub[64], da[kMaxStages], db[kMaxStages];
        for (int i = 0; i < 64; ++i) {
            CK(cudaEventCreateWithFlags(&ua[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ub[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kMaxStages; ++i) {
            CK(cudaEventCreateWithFlags(&da[i], cudaEventDisableTiming | cudaEventBlockingSync));
            CK(cudaEventCreateWithFlags(&db[i], cudaEventDisableTiming | cudaEventBlockingSync));
        }
        for (int deep : {3, 8, 16}) {
            for (int n_chunks : {16, 32}) {
                auto both = [&](bool a, bool b) {
                    CK(cudaDeviceSynchronize());
                    const auto t0 = std::chrono::steady_clock::now();
                    std::thread ta([&] { if (a) { CK(cudaSetDevice(0)); for (int i = 0; i < steps; ++i) run_call_split(enc, n_chunks, 3, deep, sa[0], sa[1], ua, da); } });
                    std::thread tb([&] { if (b) { CK(cudaSetDevice(0)); for (int i = 0; i < steps; ++i) run_call_split(dec, n_chunks, 3, deep, sb[0], sb[1], ub, db); } });
                    ta.join();
                    tb.join();
                    CK(cudaDeviceSynchronize());
                    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
                };
                both(true, true);
                const double t_ab = both(true, true), t_a = both(true, false), t_b = both(false, true);
                std::printf("{\"variant\": \"split streams\", \"big_ring\": 3, \"small_ring\": %d, \"chunks_per_call\": %d, \"both_ms\": %.2f, "
                            "\"both_gbs_each_way\": %.1f, \"encode_alone_ms\": %.2f, \"decode_alone_ms\": %.2f}\n",
                            deep, n_chunks, t_ab * 1e3, (big + small) / t_ab / 1e9, t_a * 1e3, t_b * 1e3);
                std::fflush(stdout);
            }
        }
    }
    for (int stages : {3}) {
        for (int n_chunks : {16}) {
            auto both = [&](bool a, bool b) {
                CK(cudaDeviceSynchronize());
                const auto t0 = std::chrono::steady_clock::now();
                std::thread ta([&] { if (a) { CK(cudaSetDevice(0)); for (int i = 0; i < steps; ++i) run_call(enc, n_chunks, stages, sa, ea); } });
                std::thread tb([&] { if (b) { CK(cudaSetDevice(0)); for (int i = 0; i < steps; ++i) run_call(dec, n_chunks, stages, sb, eb); } });
                ta.join();
                tb.join();
                CK(cudaDeviceSynchronize());
                return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
            };
            both(true, true);
            const double t_ab = both(true, true), t_a = both(true, false), t_b = both(false, true);
            std::printf("{\"stages\": %d, \"chunks_per_call\": %d, \"chunk_mib_big\": %.1f, \"both_ms\": %.2f, \"both_gbs_each_way\": %.1f, "
                        "\"encode_alone_ms\": %.2f, \"decode_alone_ms\": %.2f}\n",
                        stages, n_chunks, big / (double)n_chunks / (1 << 20), t_ab * 1e3, (big + small) / t_ab / 1e9, t_a * 1e3, t_b * 1e3);
            std::fflush(stdout);
        }
    }
    return 0;
}
