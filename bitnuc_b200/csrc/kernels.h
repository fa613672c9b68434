// kernels.h -- host-side launchers of the sm_100a kernels (internal to libbitnuc_cuda.so).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bn {

// Launch geometry shared by the streaming kernels: persistent grid of sm_count * blocks_per_sm CTAs.
struct DeviceInfo {
    int device = 0;
    int sm_count = 148;
};

// codec.cu
cudaError_t launch_encode(const DeviceInfo& di, const uint8_t* d_seq, size_t n, uint64_t* d_out,
                          unsigned long long* d_status, cudaStream_t s);
cudaError_t launch_decode(const DeviceInfo& di, const uint64_t* d_words, size_t n_bases, uint8_t* d_out,
                          cudaStream_t s);

// kmer.cu
cudaError_t launch_as_2bit_batch(const DeviceInfo& di, const uint8_t* d_recs, size_t n, uint32_t k, size_t stride,
                                 uint64_t* d_out, unsigned long long* d_status, cudaStream_t s);
cudaError_t launch_from_2bit_batch(const DeviceInfo& di, const uint64_t* d_packed, size_t n, uint32_t k,
                                   uint8_t* d_out, size_t stride, cudaStream_t s);

// hamming.cu
cudaError_t launch_hdist(const DeviceInfo& di, const uint64_t* d_a, const uint64_t* d_b, size_t n_bases,
                         unsigned long long* d_total, cudaStream_t s);
cudaError_t launch_hdist_pairs(const DeviceInfo& di, const uint64_t* d_u, const uint64_t* d_v, size_t n_pairs,
                               uint32_t len, uint32_t* d_out, cudaStream_t s);

// counts.cu
cudaError_t launch_base_counts(const DeviceInfo& di, const uint64_t* d_words, size_t n_bases,
                               unsigned long long* d_counts, double* d_gc, cudaStream_t s);
cudaError_t launch_base_counts_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets,
                                     const uint64_t* d_lens, size_t n_reads, size_t fixed_len, size_t n_words_hint,
                                     unsigned long long* d_counts4,
                                     double* d_gc, unsigned long long* d_totals, cudaStream_t s);

// batch.cu
size_t encode_batch_scratch_bytes(size_t n_reads, size_t n_bytes);
cudaError_t launch_encode_batch(const DeviceInfo& di, const uint8_t* d_bytes, const uint64_t* d_offsets,
                                size_t n_reads, size_t n_bytes, uint64_t* d_out_words, uint64_t* d_out_word_offsets,
                                uint32_t* d_read_status, unsigned long long* d_status, void* d_scratch,
                                cudaStream_t s);

// split.cu
size_t split_packed_scratch_bytes(size_t n_reads);
cudaError_t launch_split_packed_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets,
                                      const uint64_t* d_lens, const uint64_t* d_idx, size_t n_reads, uint64_t* d_left,
                                      uint64_t* d_left_offsets, uint64_t* d_right, uint64_t* d_right_offsets,
                                      unsigned long long* d_status, void* d_scratch, cudaStream_t s);

// gather.cu
size_t slice_batch_scratch_bytes(size_t nq);
cudaError_t launch_slice_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                               size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_start, const uint64_t* d_q_end, size_t nq,
                               uint8_t* d_out, uint64_t* d_out_offsets, unsigned long long* d_status, void* d_scratch, cudaStream_t s);
cudaError_t launch_get_batch(const DeviceInfo& di, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                             size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_index, size_t nq, uint8_t* d_out,
                             unsigned long long* d_status, cudaStream_t s);

// windows.cu
cudaError_t launch_kmer_windows(const DeviceInfo& di, const uint8_t* d_seq, size_t n, uint32_t k, uint64_t* d_out,
                                unsigned long long* d_status, cudaStream_t s);

size_t kmer_windows_batch_scratch_bytes(size_t n_reads, size_t n_bytes);
cudaError_t launch_kmer_windows_batch(const DeviceInfo& di, const uint8_t* d_bytes, const uint64_t* d_offsets, size_t n_reads, size_t n_bytes,
                                      uint32_t k, uint64_t* d_out, uint64_t* d_out_offsets, unsigned long long* d_status, void* d_scratch,
                                      cudaStream_t s);

// fastq.cu
size_t fastq_scratch_bytes(size_t n_bytes);
size_t fastq_index_scratch_bytes(size_t n_reads);
// fasta = 0: FASTQ (four lines per record, '@'); 1: FASTA with one sequence line per record (two lines, '>')
cudaError_t launch_fastq_count(const DeviceInfo& di, const uint8_t* d_bytes, size_t n_bytes, void* d_scratch, uint64_t* d_n_lines,
                               int fasta, cudaStream_t s);
cudaError_t launch_fastq_index(const DeviceInfo& di, const uint8_t* d_bytes, size_t n_bytes, size_t n_reads, void* d_scratch,
                               void* d_index_scratch, uint64_t* d_seq_offsets, uint64_t* d_seq_lens, uint64_t* d_word_offsets,
                               unsigned long long* d_status, int fasta, cudaStream_t s);
cudaError_t launch_fastq_encode(const DeviceInfo& di, const uint8_t* d_bytes, size_t n_bytes, size_t n_reads, void* d_scratch,
                                const uint64_t* d_seq_offsets, const uint64_t* d_seq_lens, const uint64_t* d_word_offsets,
                                uint64_t* d_out_words, unsigned long long* d_status, int fasta, cudaStream_t s);
// wrapped (multi-line) FASTA: after launch_fastq_count(fasta = 1) -- line table, then compaction of the sequence bytes; the
// caller runs launch_encode_batch on (d_compact, d_rec_off)
size_t fasta_wrapped_scratch_bytes(size_t n_lines);
cudaError_t launch_fasta_wrapped_index(const DeviceInfo& di, const uint8_t* d_bytes, size_t n_bytes, size_t n_lines, void* d_scratch,
                                       void* d_wscratch, uint64_t* d_totals, unsigned long long* d_status, cudaStream_t s);
cudaError_t launch_fasta_wrapped_compact(const DeviceInfo& di, const uint8_t* d_bytes, size_t n_lines, void* d_wscratch, size_t n_records,
                                         uint8_t* d_compact, uint64_t* d_rec_off, uint64_t* d_hdr_off, cudaStream_t s);

// synth.cu
cudaError_t launch_synth_words(const DeviceInfo& di, uint64_t seed, uint64_t stream_id, uint64_t first_word,
                               size_t n_words, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_synth_ascii(const DeviceInfo& di, uint64_t seed, uint64_t stream_id, uint64_t first_base,
                               size_t n, uint8_t* d_out, cudaStream_t s);

}  // namespace bn
