// gather.cu -- PackedSequence::get / PackedSequence::slice over a batch of queries (sm_100a).  SURVEY.md 8(f) rank 2.
//
// Replaces the caller-side loops over /root/reference/src/sequence.rs:116-135 (get: one 2-bit poke per call) and
// sequence.rs:198-212 (slice: a Vec filled by per-base get()): query q names a read of a packed batch (words at
// word_offsets[read], lens[read] bases) and a base index / a base range; the bases come out as upper-case ASCII
// without decoding the whole read.
//   get:   index >= len            -> IndexOutOfBounds{index, len}      (sequence.rs:117-122)
//   slice: start > end || end > len -> InvalidRange{start, end, len}     (sequence.rs:199-205)
// The first failing query in index order is reported (status word = min failing query index); failing queries
// produce no output (slice) or 0 (get).
//
// slice: an exclusive scan of (end - start) places every query's bytes; then 8 lanes per query walk the range four
// bases (one packed byte -> one 32-bit word of ASCII, PRMT as a 4-entry LUT) at a time: 32 contiguous bytes per
// 8-lane step, byte stores only for the unaligned head and tail of a range.
#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

struct SliceLen {
    const uint64_t *lens, *q_read, *q_start, *q_end;
    unsigned long long n_reads;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long q) const {
        const unsigned long long r = q_read[q], s = q_start[q], e = q_end[q];
        if (r >= n_reads || s > e || e > lens[r]) return 0;  // reported as an error, takes no room
        return e - s;
    }
};

// ASCII of the 4 bases starting at base b of the read whose words (viewed as 32-bit halves) start at w32:
// a funnel shift of two halves, the low byte spread to one code per nibble, PRMT as a 4-entry LUT.
__device__ __forceinline__ uint32_t ascii4_at(const uint32_t* __restrict__ w32, unsigned long long b) {
    const unsigned long long i = b >> 4;
    const unsigned sh = 2u * (unsigned)(b & 15u);
    const uint32_t lo = __ldg(w32 + i);
    const uint32_t hi = sh > 24 ? __ldg(w32 + i + 1) : 0u;  // only reached when all 4 bases exist, so that half does too
    uint32_t t = __funnelshift_r(lo, hi, sh) & 0xFFu;
    t = (t | (t << 4)) & 0x0F0Fu;
    t = (t | (t << 2)) & 0x3333u;
    return __byte_perm(0x54474341u, 0u, t);
}
__device__ __forceinline__ uint8_t ascii1_at(const uint32_t* __restrict__ w32, unsigned long long b) {
    return (uint8_t)(0x54474341u >> (8u * ((__ldg(w32 + (b >> 4)) >> (2u * (unsigned)(b & 15u))) & 3u)));
}

constexpr int kSliceLanes = 8;

__global__ void __launch_bounds__(kThreads)
slice_batch_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets, const uint64_t* __restrict__ lens,
                   unsigned long long n_reads, const uint64_t* __restrict__ q_read, const uint64_t* __restrict__ q_start,
                   const uint64_t* __restrict__ q_end, unsigned long long nq, uint8_t* __restrict__ out,
                   const uint64_t* __restrict__ out_offsets, unsigned long long* __restrict__ status) {
    const unsigned sub = threadIdx.x % kSliceLanes;
    const unsigned long long q = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) / kSliceLanes;
    if (q >= nq) return;
    const unsigned long long r = q_read[q], s = q_start[q], e = q_end[q];
    if (r >= n_reads || s > e || e > lens[r]) {
        if (sub == 0 && q < ld_volatile_u64(status)) atomicMin(status, q);
        return;
    }
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(words + word_offsets[r]);
    uint8_t* o = out + out_offsets[q];
    const unsigned long long n = e - s;
    // head: bytes up to the first 4-byte aligned output address
    const unsigned head = (unsigned)min((unsigned long long)((4u - (unsigned)(reinterpret_cast<uintptr_t>(o) & 3u)) & 3u), n);
    if (sub < head) o[sub] = ascii1_at(w32, s + sub);
    // body: one aligned 32-bit store of 4 bases per lane step
    const unsigned long long body = (n - head) / 4;
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o + head);
    const unsigned long long b0 = s + head;
    for (unsigned long long i = sub; i < body; i += kSliceLanes) o32[i] = ascii4_at(w32, b0 + 4 * i);
    // tail: the last (< 4) bytes
    const unsigned long long done = head + 4 * body;
    if (done + sub < n) o[done + sub] = ascii1_at(w32, s + done + sub);
}

__global__ void __launch_bounds__(kThreads)
get_batch_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets, const uint64_t* __restrict__ lens,
                 unsigned long long n_reads, const uint64_t* __restrict__ q_read, const uint64_t* __restrict__ q_index,
                 unsigned long long nq, uint8_t* __restrict__ out, unsigned long long* __restrict__ status) {
    const unsigned long long q = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (q >= nq) return;
    const unsigned long long r = q_read[q], i = q_index[q];
    if (r >= n_reads || i >= lens[r]) {
        out[q] = 0;
        if (q < ld_volatile_u64(status)) atomicMin(status, q);
        return;
    }
    const uint64_t x = __ldg(words + word_offsets[r] + (i >> 5));
    out[q] = (uint8_t)(0x54474341u >> (8 * ((x >> (2 * (i & 31))) & 3)));
}

size_t slice_batch_scratch_bytes(size_t nq) { return scan_scratch_bytes(nq); }

cudaError_t launch_slice_batch(const DeviceInfo&, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                               size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_start, const uint64_t* d_q_end, size_t nq,
                               uint8_t* d_out, uint64_t* d_out_offsets, unsigned long long* d_status, void* d_scratch, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (nq == 0) return cudaMemsetAsync(d_out_offsets, 0, sizeof(uint64_t), s);
    launch_exclusive_scan(SliceLen{d_lens, d_q_read, d_q_start, d_q_end, n_reads}, nq, static_cast<unsigned long long*>(d_scratch),
                          d_out_offsets, s);
    slice_batch_kernel<<<(unsigned)ceil_div((unsigned long long)nq * kSliceLanes, kThreads), kThreads, 0, s>>>(
        d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_start, d_q_end, nq, d_out, d_out_offsets, d_status);
    return cudaGetLastError();
}

cudaError_t launch_get_batch(const DeviceInfo&, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                             size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_index, size_t nq, uint8_t* d_out,
                             unsigned long long* d_status, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess || nq == 0) return e;
    get_batch_kernel<<<(unsigned)ceil_div(nq, kThreads), kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_index, nq,
                                                                          d_out, d_status);
    return cudaGetLastError();
}

}  // namespace bn
