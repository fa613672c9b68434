// ctx.h -- the context object behind the C ABI and the helpers api.cu and multi.cu share (internal to libbitnuc_cuda.so).
#pragma once

#include "../../include/bitnuc_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <cstdlib>
#include <vector>

#include "kernels.h"

using bn::DeviceInfo;


constexpr int kStages = 3;                       // pipeline depth of the host-pointer codec calls
constexpr size_t kDefaultChunk = 64ull << 20;    // ASCII bytes per stage
constexpr int kSlots = 8;                        // reusable device scratch buffers
constexpr unsigned long long kNoError = ~0ull;

struct Buffer {
    void* p = nullptr;
    size_t cap = 0;
};
struct HostBuffer {  // pinned
    void* p = nullptr;
    size_t cap = 0;
};


struct bn_ctx {
    DeviceInfo di;
    cudaStream_t stream = nullptr;               // context stream (device-pointer calls default to it)
    cudaStream_t stage_stream[kStages] = {};
    cudaEvent_t stage_done[kStages] = {};
    Buffer stage_in[kStages], stage_out[kStages];
    Buffer stage_aux[kStages][4];                // batch calls: offsets, word offsets, per-read status, scratch
    HostBuffer hstage_in[kStages][2], hstage_out[kStages];   // pinned bounce buffers for pageable caller memory
    Buffer slot[kSlots];
    unsigned long long* d_words = nullptr;       // 5 kStages + 4 device status / accumulator words
    unsigned long long* h_words = nullptr;       // pinned mirror
    size_t chunk = kDefaultChunk;
    // bn_fastq_scan -> bn_fastq_encode: the uploaded text and its index stay resident between the two calls
    Buffer fq[7];                                // text, scratch, index scratch, seq offsets, seq lens, word offsets, out words
    const void* fq_text = nullptr;
    size_t fq_bytes = 0, fq_reads = 0, fq_words = 0;
    bool fq_valid = false;
    int fq_fasta = 0;
    int compat = BN_COMPAT_X86_64;
    // bn_ctx_set_timing / bn_last_kernel_ms: CUDA events around the launches of the last device-pointer call
    bool timing = false, timed = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::mutex mu;
};


// True when the calling thread already has a CUDA context bound (it chose a device at some point).  A fresh thread
// reports device 0 without having asked for it, and "restoring" that would create a primary context on GPU 0 --
// hundreds of milliseconds, on a GPU that may belong to another rank.  Asked through the driver API, resolved at
// run time so that the library has no link-time dependency on libcuda.
inline bool thread_has_context() {
    using Fn = int (*)(void**);
    static const Fn fn = [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        return h ? reinterpret_cast<Fn>(dlsym(h, "cuCtxGetCurrent")) : nullptr;
    }();
    void* cur = nullptr;
    return fn == nullptr || (fn(&cur) == 0 && cur != nullptr);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        const bool bound = thread_has_context();
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        if (!bound || prev == dev) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int set_err(bn_error_t* err, int code, uint64_t a = 0, uint64_t b = 0, uint64_t c = 0) {
    if (err) {
        std::memset(err, 0, sizeof(*err));
        err->code = code;
        err->a = a;
        err->b = b;
        err->c = c;
        if (code == BN_INVALID_BASE) err->base = (uint8_t)a;
    }
    return code;
}

inline int cuda_fail(bn_error_t* err, cudaError_t e) {
    set_err(err, BN_ERR_CUDA);
    if (err) err->cuda_error = (int32_t)e;
    cudaGetLastError();  // clear the sticky-less error state
    return BN_ERR_CUDA;
}

inline int invalid_base(bn_error_t* err, unsigned long long key, uint64_t base_offset) {
    set_err(err, BN_INVALID_BASE, key & 0xFFu);
    if (err) err->offset = (key >> 8) + base_offset;
    return BN_INVALID_BASE;
}

#define BN_CUDA(expr)                                   \
    do {                                                \
        cudaError_t e__ = (expr);                       \
        if (e__ != cudaSuccess) return cuda_fail(err, e__); \
    } while (0)

inline cudaError_t ensure(Buffer& b, size_t bytes) {
    if (bytes <= b.cap) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

inline cudaError_t ensure_host(HostBuffer& b, size_t bytes) {
    if (bytes <= b.cap) return cudaSuccess;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = (bytes + 4095) & ~(size_t)4095;
    cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

// Caller memory that is neither pinned nor device memory: cudaMemcpyAsync on it is staged by the driver through one
// thread (~10 GB/s here).  The host-pointer calls bounce such buffers through their own pinned stage buffers with a
// multi-threaded memcpy instead, which keeps the PCIe pipeline fed at several times that rate.
inline bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

inline void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    static const unsigned hw = [] {   // BN_MEMCPY_THREADS overrides the default of min(8, hardware threads)
        const char* v = std::getenv("BN_MEMCPY_THREADS");
        const unsigned want = v ? (unsigned)std::atoi(v) : 8u;
        return std::max(1u, std::min(want ? want : 8u, std::thread::hardware_concurrency()));
    }();
    const size_t kMinSlice = 4u << 20;
    const unsigned t = (unsigned)std::min<size_t>(hw, bytes / kMinSlice);
    if (t <= 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t slice = ((bytes + t - 1) / t + 4095) & ~(size_t)4095;
    std::vector<std::thread> workers;
    workers.reserve(t - 1);
    for (unsigned i = 1; i < t; ++i) {
        const size_t off = i * slice;
        if (off >= bytes) break;
        workers.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(slice, bytes - off)); });
    }
    std::memcpy(dst, src, std::min(slice, bytes));
    for (auto& w : workers) w.join();
}

// Declared right after the context lock in every host-pointer call.  A CUDA error in the middle of a pipelined call
// returns early while copies issued for earlier chunks may still be reading the caller's input, writing its output or
// using the stages' pinned bounce buffers (which a later call may free): wait for all of them before the caller gets
// its memory back and before the lock is released.  On the normal path everything has retired already and this is
// four no-op synchronisations.
struct StageDrain {
    bn_ctx* ctx;
    ~StageDrain() {
        for (int s = 0; s < kStages; ++s)
            if (ctx->stage_stream[s]) cudaStreamSynchronize(ctx->stage_stream[s]);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
};

// Declared in a device-pointer call after its argument checks: with bn_ctx_set_timing(ctx, 1) the launches the call
// enqueues are bracketed by two events on the same stream, read back by bn_last_kernel_ms.  Off by default (an event
// record costs a few microseconds of stream time).
struct LaunchTimer {
    bn_ctx* ctx;
    cudaStream_t s;
    LaunchTimer(bn_ctx* c, cudaStream_t st) : ctx(c), s(st) {
        if (ctx->timing) ctx->timed = cudaEventRecord(ctx->t0, s) == cudaSuccess;
    }
    ~LaunchTimer() {
        if (ctx->timing && ctx->timed) ctx->timed = cudaEventRecord(ctx->t1, s) == cudaSuccess;
    }
};

inline cudaStream_t pick(bn_ctx* ctx, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : ctx->stream; }


