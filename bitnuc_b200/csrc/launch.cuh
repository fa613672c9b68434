// launch.cuh -- grid sizing shared by the launchers: persistent grids of sm_count x resident CTAs.
#pragma once

#include "kernels.h"

namespace bn {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

// CTAs of `kernel` resident per SM.  A property of the kernel on sm_100a (the only target), so it is cached per
// process by the launchers; the persistent grid is this times the SM count of the context's own device (contexts on
// different devices of one host must not share a cached grid size).
template <typename K>
static int blocks_per_sm(K kernel, int threads, size_t dyn_smem = 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm;
}

static inline unsigned grid_for(unsigned long long work_blocks, int resident) {
    if (work_blocks < 1) work_blocks = 1;
    return (unsigned)(work_blocks < (unsigned long long)resident ? work_blocks : (unsigned long long)resident);
}

__host__ __device__ static inline unsigned long long ceil_div(unsigned long long a, unsigned long long b) { return (a + b - 1) / b; }

}  // namespace bn
