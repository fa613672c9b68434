//! Raw `extern "C"` bindings, 1:1 with `include/bitnuc_cuda.h`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct bn_error_t {
    pub code: i32,
    pub base: u8,
    pub pad_: [u8; 3],
    pub a: u64,
    pub b: u64,
    pub c: u64,
    pub offset: u64,
    pub record: u64,
    pub cuda_error: i32,
    pub pad2_: i32,
}

#[repr(C)]
pub struct bn_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct bn_multi {
    _private: [u8; 0],
}

pub const BN_OK: c_int = 0;
pub const BN_ERR_EMPTY_ENCODE: c_int = -3;
pub const BN_ERR_COLLECTIVE: c_int = -6;
pub const BN_REDUCE_NCCL: c_int = 0;
pub const BN_REDUCE_P2P: c_int = 1;

extern "C" {
    pub fn bn_abi_version() -> c_int;
    pub fn bn_device_count() -> c_int;
    pub fn bn_error_string(err: *const bn_error_t, buf: *mut c_char, cap: usize) -> c_int;
    pub fn bn_ctx_create(device: c_int, out: *mut *mut bn_ctx) -> c_int;
    pub fn bn_ctx_destroy(ctx: *mut bn_ctx);
    pub fn bn_ctx_device(ctx: *const bn_ctx) -> c_int;
    pub fn bn_ctx_stream(ctx: *const bn_ctx) -> *mut c_void;
    pub fn bn_ctx_synchronize(ctx: *mut bn_ctx) -> c_int;
    pub fn bn_ctx_set_chunk_bytes(ctx: *mut bn_ctx, bytes: usize) -> c_int;
    pub fn bn_dev_alloc(ctx: *mut bn_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn bn_dev_free(ctx: *mut bn_ctx, ptr: *mut c_void) -> c_int;
    pub fn bn_host_alloc(ctx: *mut bn_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn bn_host_free(ctx: *mut bn_ctx, ptr: *mut c_void) -> c_int;
    pub fn bn_copy_h2d(ctx: *mut bn_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn bn_copy_d2h(ctx: *mut bn_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;

    pub fn bn_encode(ctx: *mut bn_ctx, seq: *const u8, n: usize, out: *mut u64, n_words: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_decode(ctx: *mut bn_ctx, words: *const u64, n_words: usize, n_bases: usize, out: *mut u8, err: *mut bn_error_t) -> c_int;
    pub fn bn_as_2bit_batch(ctx: *mut bn_ctx, recs: *const u8, n: usize, k: u32, stride: usize, out: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_from_2bit_batch(ctx: *mut bn_ctx, packed: *const u64, n: usize, k: u32, out: *mut u8, stride: usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_hdist(ctx: *mut bn_ctx, a: *const u64, n_words_a: usize, b: *const u64, n_words_b: usize, n_bases: usize, total: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_hdist_pairs(ctx: *mut bn_ctx, u: *const u64, v: *const u64, n_pairs: usize, len: u32, out: *mut u32, err: *mut bn_error_t) -> c_int;
    pub fn bn_base_counts(ctx: *mut bn_ctx, words: *const u64, n_words: usize, n_bases: usize, counts: *mut u64, gc: *mut f64, err: *mut bn_error_t) -> c_int;
    pub fn bn_base_counts_batch(ctx: *mut bn_ctx, words: *const u64, n_words: usize, word_offsets: *const u64, lens: *const u64, n_reads: usize, counts4: *mut u64, gc: *mut f64, totals: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_encode_batch(ctx: *mut bn_ctx, bytes: *const u8, offsets: *const u64, n_reads: usize, out_words: *mut u64, out_word_offsets: *mut u64, read_status: *mut u32, err: *mut bn_error_t) -> c_int;

    pub fn bn_encode_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_seq: *const u8, n: usize, d_out: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_decode_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, n_words: usize, n_bases: usize, d_out: *mut u8) -> c_int;
    pub fn bn_as_2bit_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_recs: *const u8, n: usize, k: u32, stride: usize, d_out: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_from_2bit_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_packed: *const u64, n: usize, k: u32, d_out: *mut u8, stride: usize) -> c_int;
    pub fn bn_hdist_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_a: *const u64, d_b: *const u64, n_bases: usize, d_total: *mut u64) -> c_int;
    pub fn bn_hdist_pairs_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_u: *const u64, d_v: *const u64, n_pairs: usize, len: u32, d_out: *mut u32) -> c_int;
    pub fn bn_base_counts_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, n_bases: usize, d_counts: *mut u64, d_gc: *mut f64) -> c_int;
    pub fn bn_base_counts_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, n_words: usize, d_word_offsets: *const u64, d_lens: *const u64, n_reads: usize, d_counts4: *mut u64, d_gc: *mut f64, d_totals: *mut u64) -> c_int;
    pub fn bn_base_counts_fixed_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, n_reads: usize, read_len: usize, d_counts4: *mut u64, d_gc: *mut f64, d_totals: *mut u64) -> c_int;
    pub fn bn_encode_batch_scratch_bytes(n_reads: usize, n_bytes: usize) -> usize;
    pub fn bn_encode_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_bytes: *const u8, d_offsets: *const u64, n_reads: usize, n_bytes: usize, d_out_words: *mut u64, d_out_word_offsets: *mut u64, d_read_status: *mut u32, d_status: *mut u64, d_scratch: *mut c_void) -> c_int;
    pub fn bn_split_packed_batch(ctx: *mut bn_ctx, words: *const u64, n_words: usize, word_offsets: *const u64, lens: *const u64, idx: *const u64, n_reads: usize, left: *mut u64, left_offsets: *mut u64, right: *mut u64, right_offsets: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_split_packed_scratch_bytes(n_reads: usize) -> usize;
    pub fn bn_split_packed_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, d_word_offsets: *const u64, d_lens: *const u64, d_idx: *const u64, n_reads: usize, d_left: *mut u64, d_left_offsets: *mut u64, d_right: *mut u64, d_right_offsets: *mut u64, d_status: *mut u64, d_scratch: *mut c_void) -> c_int;
    pub fn bn_fastq_scan(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_reads: *mut usize, n_words: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_fastq_encode(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_reads: usize, n_words: usize, out_words: *mut u64, out_word_offsets: *mut u64, seq_offsets: *mut u64, seq_lens: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_fasta_scan(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_reads: *mut usize, n_words: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_fasta_encode(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_reads: usize, n_words: usize, out_words: *mut u64, out_word_offsets: *mut u64, seq_offsets: *mut u64, seq_lens: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_fasta_wrapped_scan(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_records: *mut usize, n_bases: *mut usize, n_words: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_fasta_wrapped_encode(ctx: *mut bn_ctx, text: *const u8, n_bytes: usize, n_records: usize, n_words: usize, out_words: *mut u64, out_word_offsets: *mut u64, header_offsets: *mut u64, seq_lens: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_fasta_count_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, d_scratch: *mut c_void, d_n_lines: *mut u64) -> c_int;
    pub fn bn_fasta_index_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, n_reads: usize, d_scratch: *mut c_void, d_index_scratch: *mut c_void, d_seq_offsets: *mut u64, d_seq_lens: *mut u64, d_word_offsets: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_fasta_encode_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, n_reads: usize, d_scratch: *mut c_void, d_seq_offsets: *const u64, d_seq_lens: *const u64, d_word_offsets: *const u64, d_out_words: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_fasta_status_fetch(ctx: *mut bn_ctx, stream: *mut c_void, d_status: *const u64, n_lines: u64, d_seq_offsets: *const u64, n_reads: usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_ctx_set_compat(ctx: *mut bn_ctx, mode: c_int) -> c_int;
    pub fn bn_ctx_compat(ctx: *const bn_ctx) -> c_int;
    pub fn bn_fastq_scratch_bytes(n_bytes: usize) -> usize;
    pub fn bn_fastq_index_scratch_bytes(n_reads: usize) -> usize;
    pub fn bn_fastq_count_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, d_scratch: *mut c_void, d_n_lines: *mut u64) -> c_int;
    pub fn bn_fastq_index_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, n_reads: usize, d_scratch: *mut c_void, d_index_scratch: *mut c_void, d_seq_offsets: *mut u64, d_seq_lens: *mut u64, d_word_offsets: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_fastq_encode_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_text: *const u8, n_bytes: usize, n_reads: usize, d_scratch: *mut c_void, d_seq_offsets: *const u64, d_seq_lens: *const u64, d_word_offsets: *const u64, d_out_words: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_fastq_status_fetch(ctx: *mut bn_ctx, stream: *mut c_void, d_status: *const u64, n_lines: u64, d_seq_offsets: *const u64, n_reads: usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_kmers(ctx: *mut bn_ctx, seq: *const u8, n: usize, k: u32, out: *mut u64, n_out: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_kmers_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_seq: *const u8, n: usize, k: u32, d_out: *mut u64, d_status: *mut u64) -> c_int;
    pub fn bn_kmers_batch(ctx: *mut bn_ctx, bytes: *const u8, offsets: *const u64, n_reads: usize, k: u32, out: *mut u64, out_cap: usize, out_offsets: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_kmers_batch_scratch_bytes(n_reads: usize, n_bytes: usize) -> usize;
    pub fn bn_kmers_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_bytes: *const u8, d_offsets: *const u64, n_reads: usize, n_bytes: usize, k: u32, d_out: *mut u64, d_out_offsets: *mut u64, d_status: *mut u64, d_scratch: *mut c_void) -> c_int;
    pub fn bn_slice_batch(ctx: *mut bn_ctx, words: *const u64, n_words: usize, word_offsets: *const u64, lens: *const u64, n_reads: usize, q_read: *const u64, q_start: *const u64, q_end: *const u64, nq: usize, out: *mut u8, out_cap: usize, out_offsets: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_get_batch(ctx: *mut bn_ctx, words: *const u64, n_words: usize, word_offsets: *const u64, lens: *const u64, n_reads: usize, q_read: *const u64, q_index: *const u64, nq: usize, out: *mut u8, err: *mut bn_error_t) -> c_int;
    pub fn bn_slice_batch_scratch_bytes(nq: usize) -> usize;
    pub fn bn_slice_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, d_word_offsets: *const u64, d_lens: *const u64, n_reads: usize, d_q_read: *const u64, d_q_start: *const u64, d_q_end: *const u64, nq: usize, d_out: *mut u8, d_out_offsets: *mut u64, d_status: *mut u64, d_scratch: *mut c_void) -> c_int;
    pub fn bn_get_batch_dev(ctx: *mut bn_ctx, stream: *mut c_void, d_words: *const u64, d_word_offsets: *const u64, d_lens: *const u64, n_reads: usize, d_q_read: *const u64, d_q_index: *const u64, nq: usize, d_out: *mut u8, d_status: *mut u64) -> c_int;
    pub fn bn_status_fetch(ctx: *mut bn_ctx, stream: *mut c_void, d_status: *const u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_synth_words_dev(ctx: *mut bn_ctx, stream: *mut c_void, seed: u64, stream_id: u64, first_word: u64, n_words: usize, d_out: *mut u64) -> c_int;
    pub fn bn_synth_ascii_dev(ctx: *mut bn_ctx, stream: *mut c_void, seed: u64, stream_id: u64, first_base: u64, n: usize, d_out: *mut u8) -> c_int;
    pub fn bn_ctx_set_timing(ctx: *mut bn_ctx, on: c_int) -> c_int;
    pub fn bn_last_kernel_ms(ctx: *mut bn_ctx, ms: *mut f32) -> c_int;

    // multi-GPU: one process, N devices (include/bitnuc_cuda.h, "multi-GPU")
    pub fn bn_multi_create(devs: *const c_int, n: c_int, reduce: c_int, out: *mut *mut bn_multi) -> c_int;
    pub fn bn_multi_destroy(m: *mut bn_multi);
    pub fn bn_multi_size(m: *const bn_multi) -> c_int;
    pub fn bn_multi_ctx(m: *mut bn_multi, i: c_int) -> *mut bn_ctx;
    pub fn bn_multi_reduce(m: *const bn_multi) -> c_int;
    pub fn bn_multi_nccl_version(m: *const bn_multi) -> c_int;
    pub fn bn_multi_set_chunk_bytes(m: *mut bn_multi, bytes: usize) -> c_int;
    pub fn bn_multi_synchronize(m: *mut bn_multi) -> c_int;
    pub fn bn_multi_shard_units(m: *const bn_multi, n_units: usize, align: usize, starts: *mut usize) -> c_int;
    pub fn bn_multi_shard_reads(m: *const bn_multi, offsets: *const u64, n_reads: usize, starts: *mut usize) -> c_int;
    pub fn bn_multi_encode(m: *mut bn_multi, seq: *const u8, n: usize, out: *mut u64, n_words: *mut usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_decode(m: *mut bn_multi, words: *const u64, n_words: usize, n_bases: usize, out: *mut u8, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_as_2bit_batch(m: *mut bn_multi, recs: *const u8, n: usize, k: u32, stride: usize, out: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_from_2bit_batch(m: *mut bn_multi, packed: *const u64, n: usize, k: u32, out: *mut u8, stride: usize, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_hdist(m: *mut bn_multi, a: *const u64, n_words_a: usize, b: *const u64, n_words_b: usize, n_bases: usize, total: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_hdist_pairs(m: *mut bn_multi, u: *const u64, v: *const u64, n_pairs: usize, len: u32, out: *mut u32, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_base_counts(m: *mut bn_multi, words: *const u64, n_words: usize, n_bases: usize, counts: *mut u64, gc: *mut f64, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_base_counts_batch(m: *mut bn_multi, words: *const u64, n_words: usize, word_offsets: *const u64, lens: *const u64, n_reads: usize, counts4: *mut u64, gc: *mut f64, totals: *mut u64, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_encode_batch(m: *mut bn_multi, bytes: *const u8, offsets: *const u64, n_reads: usize, out_words: *mut u64, out_word_offsets: *mut u64, read_status: *mut u32, err: *mut bn_error_t) -> c_int;
    pub fn bn_multi_base_counts_dev(m: *mut bn_multi, d_words: *const *const u64, n_bases: *const usize, d_counts: *const *mut u64, d_gc: *const *mut f64) -> c_int;
    pub fn bn_multi_base_counts_fixed_dev(m: *mut bn_multi, d_words: *const *const u64, n_reads: *const usize, read_len: usize, d_counts4: *const *mut u64, d_gc_reads: *const *mut f64, d_totals: *const *mut u64, d_gc: *const *mut f64) -> c_int;
    pub fn bn_multi_hdist_dev(m: *mut bn_multi, d_a: *const *const u64, d_b: *const *const u64, n_bases: *const usize, d_total: *const *mut u64) -> c_int;
    pub fn bn_multi_allreduce_u64_dev(m: *mut bn_multi, d_buf: *const *mut u64, count: c_int) -> c_int;
    pub fn bn_multi_last_ms(m: *mut bn_multi, ms: *mut f32) -> c_int;
}
