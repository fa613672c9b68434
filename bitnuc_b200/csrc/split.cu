// split.cu -- split_packed over a batch of packed reads (sm_100a).  SURVEY.md 8(f) rank 1.
//
// Replaces the caller-side loop over /root/reference/src/utils/functions/split.rs:14-102: read r
// (lens[r] bases, `ebuf` = words[word_offsets[r] .. word_offsets[r+1]), normally ceil(lens[r]/32) words) is
// split at base idx[r] into a left buffer of idx/32 + 1 words (an all-zero extra word when idx % 32 == 0,
// split.rs:51,72-77) and a right buffer of ebuf.len() - idx/32 words (the loop at split.rs:83-94 pushes one
// word per remaining input word); idx == 0 / idx == len copy ebuf through (split.rs:33-42); an empty ebuf
// gives two empty outputs (split.rs:45-47).  The reference's sequential carry loop (split.rs:83-94) makes right word j
//     (ebuf[c+j] >> shift) | (ebuf[c+j-1] << (64 - shift))      c = idx/32, shift = 2*(idx%32), j >= 1
// i.e. it carries the PREVIOUS word's low bits upward; that is reproduced bit for bit (for reads longer
// than 32 bases split off a word boundary it is not the suffix of the read).  The final
// `carry != 0 && rbuf.len() < right_chunks` push (split.rs:97-99) can never fire when ebuf holds at least
// ceil(len/32) words, because ceil(len/32) - idx/32 >= ceil((len-idx)/32); a shorter ebuf is rejected here
// (the reference either panics at split.rs:77 or returns a truncated right half).
//
// Every output word depends on at most two input words: one thread per read here (reads on this path
// are short: barcode | insert splits, benches/functions_benchmark.rs:59 uses 30-280 bases).
// Status word = min over failing reads of (read index << 1 | kind): kind 0 = idx > len ->
// IndexOutOfBounds{idx, len} (split.rs:22-27), kind 1 = 0 < idx < len and a non-empty ebuf shorter than ceil(len/32) words.
#include "common.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace bn {

// (left, right) word counts of read r: they depend on (ebuf.len(), len, idx) only
struct SplitShape {
    const uint64_t *word_offsets, *lens, *idx;
    __device__ __forceinline__ ulonglong2 operator()(unsigned long long r) const {
        const unsigned long long slen = lens[r], i = idx[r], nw = word_offsets[r + 1] - word_offsets[r];
        if (i > slen || (i && i < slen && nw && nw < (slen + 31) / 32)) return make_ulonglong2(0, 0);  // an error, takes no room
        if (i == 0) return make_ulonglong2(0, nw);
        if (i == slen) return make_ulonglong2(nw, 0);
        if (nw == 0) return make_ulonglong2(0, 0);
        return make_ulonglong2(i / 32 + 1, nw - i / 32);
    }
};

__global__ void __launch_bounds__(kThreads)
split_packed_kernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_offsets,
                    const uint64_t* __restrict__ lens, const uint64_t* __restrict__ idx, unsigned long long n_reads,
                    uint64_t* __restrict__ left, const uint64_t* __restrict__ left_offsets, uint64_t* __restrict__ right,
                    const uint64_t* __restrict__ right_offsets, unsigned long long* __restrict__ status) {
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned long long wo = word_offsets[r];
    const unsigned long long slen = lens[r], i = idx[r], nw = word_offsets[r + 1] - wo;
    const bool oob = i > slen;
    if (oob || (i && i < slen && nw && nw < (slen + 31) / 32)) {
        const unsigned long long key = r << 1 | (oob ? 0ull : 1ull);
        if (key < ld_volatile_u64(status)) atomicMin(status, key);
        return;
    }
    const uint64_t* w = words + wo;
    uint64_t* lo = left + left_offsets[r];
    uint64_t* ro = right + right_offsets[r];
    if (i == 0) {
        for (unsigned long long j = 0; j < nw; ++j) ro[j] = w[j];
    } else if (i == slen) {
        for (unsigned long long j = 0; j < nw; ++j) lo[j] = w[j];
    } else if (nw) {
        const unsigned long long c = i / 32;
        const unsigned sh = 2 * (unsigned)(i % 32);
        for (unsigned long long j = 0; j < c; ++j) lo[j] = w[j];
        lo[c] = sh ? w[c] & ((1ull << sh) - 1ull) : 0ull;
        uint64_t prev = 0;
        for (unsigned long long j = 0; c + j < nw; ++j) {
            const uint64_t cur = w[c + j];
            ro[j] = (cur >> sh) | (sh && j ? prev << (64 - sh) : 0ull);
            prev = cur;
        }
    }
}

size_t split_packed_scratch_bytes(size_t n_reads) { return scan2_scratch_bytes(n_reads); }

cudaError_t launch_split_packed_batch(const DeviceInfo&, const uint64_t* d_words, const uint64_t* d_word_offsets,
                                      const uint64_t* d_lens, const uint64_t* d_idx, size_t n_reads, uint64_t* d_left,
                                      uint64_t* d_left_offsets, uint64_t* d_right, uint64_t* d_right_offsets,
                                      unsigned long long* d_status, void* d_scratch, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (n_reads == 0) {
        e = cudaMemsetAsync(d_left_offsets, 0, sizeof(uint64_t), s);
        return e != cudaSuccess ? e : cudaMemsetAsync(d_right_offsets, 0, sizeof(uint64_t), s);
    }
    unsigned long long* sums = static_cast<unsigned long long*>(d_scratch);
    launch_exclusive_scan2(SplitShape{d_word_offsets, d_lens, d_idx}, n_reads, sums, d_left_offsets, d_right_offsets, s);
    split_packed_kernel<<<(unsigned)ceil_div(n_reads, kThreads), kThreads, 0, s>>>(d_words, d_word_offsets, d_lens, d_idx, n_reads, d_left,
                                                                                    d_left_offsets, d_right, d_right_offsets, d_status);
    return cudaGetLastError();
}

}  // namespace bn
