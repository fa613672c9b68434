// common.cuh -- device helpers shared by the bitnuc sm_100a kernels.
//
// Everything on this path is HBM-bound byte/integer streaming: no tensor cores, no reuse.  The
// helpers below are the streaming loads/stores (128-bit, L1 no-allocate) and the word-parallel
// nucleotide arithmetic (4 bases per 32-bit register).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

constexpr int kSMs = 148;                   // B200: 2 dies x 74 SMs
constexpr unsigned long long kNoError = ~0ull;

// ---------------------------------------------------------------- streaming memory access ----

__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_v2(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(uint4* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_v2(uint2* p, uint2 v) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// ---------------------------------------------------------------- cache policies, tile walk ---

enum { LD_NC_NOALLOC = 0, LD_PLAIN = 1, LD_CS = 2, LD_LU = 3 };
enum { ST_CS = 0, ST_PLAIN = 1, ST_WT = 2, ST_NOALLOC = 3 };

template <int LP>
__device__ __forceinline__ uint4 ld128(const uint4* p) {
    uint4 r;
    if (LP == LD_NC_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (LP == LD_PLAIN)
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (LP == LD_CS)
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int LP>
__device__ __forceinline__ uint32_t ld32(const uint32_t* p) {
    uint32_t r;
    if (LP == LD_NC_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    else if (LP == LD_PLAIN)
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    else if (LP == LD_CS)
        asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(r) : "l"(p));
    else
        asm volatile("ld.global.lu.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <int SP>
__device__ __forceinline__ void st32(uint32_t* p, uint32_t v) {
    if (SP == ST_CS)
        asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if (SP == ST_PLAIN)
        asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if (SP == ST_WT)
        asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else
        asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int SP>
__device__ __forceinline__ void st128(uint4* p, uint4 v) {
    if (SP == ST_CS)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (SP == ST_PLAIN)
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (SP == ST_WT)
        asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit global accesses (sm_100+: LDG.256 / STG.256): one full 32-byte sector per lane.
struct uint8x {
    uint4 lo, hi;
};
template <int LP>
__device__ __forceinline__ uint8x ld256(const void* p) {
    uint8x r;
    if (LP == LD_NC_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w) : "l"(p));
    else
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w) : "l"(p));
    return r;
}
template <int SP>
__device__ __forceinline__ void st256(void* p, uint4 a, uint4 b) {
    if (SP == ST_CS)
        asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                     "r"(b.z), "r"(b.w) : "memory");
    else if (SP == ST_NOALLOC)
        asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                     "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    else
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                     "r"(b.z), "r"(b.w) : "memory");
}

// Maps (CTA, warp, round) to tile indices for the two scheduling modes.
template <int THREADS, int SCHED, int T>
struct TileWalk {
    static constexpr unsigned kWarps = THREADS / 32;
    unsigned long long first, step, end;
    __device__ __forceinline__ TileWalk(unsigned long long n_tiles) {
        const unsigned warp = threadIdx.x >> 5;
        if (SCHED == 0) {
            first = (unsigned long long)blockIdx.x * kWarps + warp;
            step = (unsigned long long)gridDim.x * kWarps;
            end = n_tiles;
        } else {  // CTA b owns tiles [b*kWarps*T, (b+1)*kWarps*T): round i, warp w -> b*kWarps*T + i*kWarps + w
            first = (unsigned long long)blockIdx.x * (kWarps * T) + warp;
            step = kWarps;
            const unsigned long long stop = (unsigned long long)(blockIdx.x + 1) * (kWarps * T);
            end = stop < n_tiles ? stop : n_tiles;
        }
    }
    // grid size for n_tiles
    static unsigned long long ctas(unsigned long long n_tiles) { return (n_tiles + kWarps * T - 1) / (kWarps * T); }
};

// ---------------------------------------------------------------- nucleotide arithmetic ------
// ASCII -> 2-bit code, 4 bases per 32-bit word w (little-endian: byte k = base k):
//   code  = ((b >> 1) ^ (b >> 2)) & 3        A/a=0 C/c=1 G/g=2 T/t=3   (bit 5 = case is ignored)
//   valid <=> b7=0, b6=1, b3=0 and (b4,b0) = (1,0) when b2&~b1 (the T/t column) else (0,1)
// The validity test is independent of the code mapping ('N' also maps to a code), and is
// accumulated as an OR of mismatches that is masked and tested once per thread.

constexpr uint32_t kValidMask = 0xD9D9D9D9u;   // bits 7,6,4,3,0 of every byte

// Returns a word whose TOP byte holds the four codes of w packed LSB-first (c0 | c1<<2 | c2<<4 |
// c3<<6); lower bits are garbage.  `bad` accumulates validity mismatches (test with kValidMask).
__device__ __forceinline__ uint32_t pack4_top(uint32_t w, uint32_t& bad) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t code = (s1 ^ s2) & 0x03030303u;
    const uint32_t tcol = s2 & ~s1 & 0x01010101u;          // 1 in bytes whose (b2,b1) = (1,0)
    const uint32_t expect = tcol * 0x11u + 0x41414141u;    // 0x52 in the T column, else 0x41
    bad |= w ^ expect;
    return code * 0x01041040u;                             // gather: no two terms collide below bit 32
}

// 16 ASCII bases (one 128-bit load) -> 32 bits of packed codes.
__device__ __forceinline__ uint32_t pack16(uint4 v, uint32_t& bad) {
    const uint32_t p0 = pack4_top(v.x, bad), p1 = pack4_top(v.y, bad);
    const uint32_t p2 = pack4_top(v.z, bad), p3 = pack4_top(v.w, bad);
    const uint32_t lo = __byte_perm(p0, p1, 0x7373);       // [p0.3, p1.3, p0.3, p1.3]
    const uint32_t hi = __byte_perm(p2, p3, 0x7373);
    return __byte_perm(lo, hi, 0x5410);                    // [p0.3, p1.3, p2.3, p3.3]
}

__device__ __forceinline__ bool byte_is_valid(uint32_t b) {
    const uint32_t l = b | 0x20u;
    return l == 'a' || l == 'c' || l == 'g' || l == 't';
}

// Index (0..15) of the first invalid byte in a 16-byte vector, or 16.
static __device__ __noinline__ int first_invalid16(uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        if (!byte_is_valid(b)) return i;
    }
    return 16;
}

// status word = min over invalid bases of (offset << 8 | byte)
__device__ __forceinline__ void report_invalid(unsigned long long* status, unsigned long long offset, uint32_t byte) {
    const unsigned long long key = (offset << 8) | (unsigned long long)(byte & 0xFFu);
    if (key < ld_volatile_u64(status)) atomicMin(status, key);
}

// 2-bit code -> ASCII, one packed byte (4 bases) -> one 32-bit word of 4 ASCII bytes.
__device__ __forceinline__ uint32_t ascii4_of_byte(uint32_t e) {
    constexpr uint32_t kAcgt = 0x54474341u;  // 'A','C','G','T' as bytes 0..3
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) r |= ((kAcgt >> (8 * ((e >> (2 * k)) & 3u))) & 0xFFu) << (8 * k);
    return r;
}

// Register-only decode: no table in memory at all.  Each half-word (8 bases) is spread so that every 2-bit code
// sits in its own nibble, then PRMT with the constant "ACGT" word acts as a 4-entry byte LUT whose
// selector nibbles are the codes.
__device__ __forceinline__ uint32_t spread_codes(uint32_t t /* [b_lo, 0, b_hi, 0] */) {
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    return (t | (t << 2)) & 0x33333333u;
}
__device__ __forceinline__ uint4 decode16_prmt(uint32_t w) {
    constexpr uint32_t kAcgt = 0x54474341u;  // 'A','C','G','T'
    const uint32_t lo = spread_codes(__byte_perm(w, 0, 0x4140));
    const uint32_t hi = spread_codes(__byte_perm(w, 0, 0x4342));
    return make_uint4(__byte_perm(kAcgt, kAcgt, lo), __byte_perm(kAcgt, kAcgt, lo >> 16),
                      __byte_perm(kAcgt, kAcgt, hi), __byte_perm(kAcgt, kAcgt, hi >> 16));
}

// per-base mismatch mask of two packed words: bit 2i set iff base i differs
__device__ __forceinline__ uint32_t mismatch_mask(uint32_t x /* u ^ v */) {
    return (x | (x >> 1)) & 0x55555555u;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------- reductions -----------------

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of `v` over the block, valid in thread 0.  `scratch` holds >= 32 values.
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum_u64(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    unsigned long long r = 0;
    if (warp == 0) {
        r = lane < (int)((blockDim.x + 31) >> 5) ? scratch[lane] : 0ull;
        r = warp_sum_u64(r);
    }
    __syncthreads();
    return r;
}

}  // namespace bn
