/*
 * bitnuc_cuda.h -- C ABI of the B200-native bitnuc hot path (libbitnuc_cuda.so).
 *
 * This is the drop-in boundary.  The reference (drbh/bitnuc, a pure-Rust crate) has no FFI of its
 * own: its boundary is the set of safe-Rust functions re-exported at /root/reference/src/lib.rs:214-220.
 * Each entry point below names the reference function it replaces (paths relative to
 * /root/reference); the `bitnuc-cuda` Rust crate (rust/bitnuc-cuda) binds these
 * symbols 1:1 and re-creates the reference signatures on top of them.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  `stream` arguments are a cudaStream_t
 *     passed as void* (NULL = the context's own stream).
 *   - every function returns a bn_status: 0 = Ok, 1..6 = the six NucleotideError variants in
 *     declaration order (src/error.rs:4-18), negative = failure outside the reference's vocabulary.
 *   - `bn_error_t *err` (may be NULL) receives the variant payload.
 *   - Packed layout is the reference's: base i of a word at bits [2i,2i+1], A=00 C=01 G=10 T=11,
 *     32 bases per uint64_t, zero-padded tail (src/utils/packing/mod.rs:13-20, src/lib.rs:96-98).
 *   - There is no CPU fallback: every call runs CUDA kernels built for sm_100a and fails with
 *     BN_ERR_CUDA when no such device is usable.
 *   - Host-pointer calls are synchronous and include the PCIe copies.  Any host memory works: pinned buffers
 *     (bn_host_alloc, cudaHostRegister) run at the PCIe link rate, pageable ones are bounced through the
 *     context's pinned stage buffers by a multi-threaded memcpy.  `_dev` calls take device
 *     pointers, only enqueue work on `stream`, and report validation results through a device-side
 *     status word that is read back later with bn_status_fetch.
 *   - A context is bound to one device and is internally synchronised: it may be shared by host
 *     threads, calls on one context serialise.  Use one context per thread for concurrency.
 */
#ifndef BITNUC_CUDA_H
#define BITNUC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_ABI_VERSION 2

typedef enum bn_status {
    BN_OK = 0,
    BN_INVALID_BASE = 1,        /* NucleotideError::InvalidBase(u8)                 src/error.rs:5  */
    BN_SEQUENCE_TOO_LONG = 2,   /* NucleotideError::SequenceTooLong(usize)          src/error.rs:6  */
    BN_INVALID_LENGTH = 3,      /* NucleotideError::InvalidLength(usize)            src/error.rs:7  */
    BN_INDEX_OUT_OF_BOUNDS = 4, /* NucleotideError::IndexOutOfBounds{index,length}  src/error.rs:8  */
    BN_INVALID_RANGE = 5,       /* NucleotideError::InvalidRange{start,end,length}  src/error.rs:12 */
    BN_UNSUPPORTED = 6,         /* NucleotideError::Unsupported                     src/error.rs:17 */
    BN_ERR_CUDA = -1,           /* CUDA runtime failure; err->cuda_error holds the cudaError_t */
    BN_ERR_ARGUMENT = -2,       /* NULL / misaligned / inconsistent arguments */
    BN_ERR_EMPTY_ENCODE = -3,   /* encode of an empty sequence: the reference panics
                                   (src/utils/packing/avx.rs:138); bindings should panic too */
    BN_ERR_NOMEM = -4,
    BN_ERR_FASTQ = -5,          /* malformed FASTQ / FASTA text (bn_fastq_*, bn_fasta_*): err->record = the record, err->a = bn_fastq_fault */
    BN_ERR_COLLECTIVE = -6      /* bn_multi_*: NCCL missing / failed (err->cuda_error = ncclResult_t) or a peer never arrived */
} bn_status;

typedef enum bn_fastq_fault {
    BN_FASTQ_BAD_HEADER = 1,          /* the record's first line does not start with '@' (FASTA: '>') */
    BN_FASTQ_BAD_SEPARATOR = 2,       /* line 4r+2 does not start with '+' */
    BN_FASTQ_BAD_QUALITY_LENGTH = 3,  /* line 4r+3 is not as long as the sequence */
    BN_FASTQ_TRUNCATED = 4            /* the text ends inside record r */
} bn_fastq_fault;

typedef struct bn_error {
    int32_t code;        /* bn_status */
    uint8_t base;        /* InvalidBase: the offending byte (printed as a decimal integer) */
    uint8_t pad_[3];
    uint64_t a, b, c;    /* payload fields in declaration order (length | index,length | start,end,length) */
    uint64_t offset;     /* InvalidBase: byte offset of the first invalid base in the input */
    uint64_t record;     /* batched calls: index of the first failing record / read */
    int32_t cuda_error;  /* BN_ERR_CUDA: cudaError_t */
    int32_t pad2_;
} bn_error_t;

typedef struct bn_ctx bn_ctx;

/* ------------------------------------------------------------------ library / context ------- */

int bn_abi_version(void);
/* Number of usable CUDA devices (0 when none; never negative). */
int bn_device_count(void);
/* Display string of an error (src/error.rs:20-45).  Returns bytes written excluding the NUL. */
int bn_error_string(const bn_error_t *err, char *buf, size_t cap);

int bn_ctx_create(int device, bn_ctx **out);
void bn_ctx_destroy(bn_ctx *ctx);
int bn_ctx_device(const bn_ctx *ctx);
void *bn_ctx_stream(const bn_ctx *ctx);        /* the context's cudaStream_t */
int bn_ctx_synchronize(bn_ctx *ctx);
/* Staging chunk (bytes of ASCII per pipeline stage) used by the host-pointer calls. 0 = default. */
int bn_ctx_set_chunk_bytes(bn_ctx *ctx, size_t bytes);

/* Which of the reference's per-ISA paths is mirrored where they disagree.  BN_COMPAT_X86_64 (default):
 * src/utils/packing/avx.rs, src/utils/unpacking/avx.rs -- what every other comment in this header describes.
 * BN_COMPAT_AARCH64: src/utils/packing/aarch64.rs:173-219 -- bn_encode reports, for an invalid byte inside a whole
 * 32-base block, the FIRST byte of that block as err->base (`InvalidBase(*ip)`, aarch64.rs:194-196; err->offset stays
 * the offending byte's offset); an invalid byte in the ragged tail, or in a sequence shorter than 32, is reported
 * itself (aarch64.rs:208-214, :223-227); bn_encode of an empty sequence is Ok with one zero word (as_2bit(b"") = 0 is
 * pushed, aarch64.rs:223-227) instead of BN_ERR_EMPTY_ENCODE.  The Vec semantics that differ (append vs overwrite,
 * src/utils/unpacking/aarch64.rs:127-130) belong to the language bindings, which read the mode back with bn_ctx_compat. */
typedef enum bn_compat { BN_COMPAT_X86_64 = 0, BN_COMPAT_AARCH64 = 1 } bn_compat;
int bn_ctx_set_compat(bn_ctx *ctx, int mode);
int bn_ctx_compat(const bn_ctx *ctx);

/* Device timing of the device-pointer calls (SURVEY.md 8b `bn_last_kernel_ms`).  With timing on, every *_dev call on
 * this context brackets the launches it enqueues with two CUDA events on the stream it uses; bn_last_kernel_ms waits
 * for the last such call and returns its duration (the kernels only: no copies, no host time).  Off by default.
 * BN_ERR_ARGUMENT when timing is off or nothing has been timed yet. */
int bn_ctx_set_timing(bn_ctx *ctx, int on);
int bn_last_kernel_ms(bn_ctx *ctx, float *ms);

/* Memory helpers for callers without their own CUDA runtime binding. */
int bn_dev_alloc(bn_ctx *ctx, size_t bytes, void **out);
int bn_dev_free(bn_ctx *ctx, void *ptr);
int bn_host_alloc(bn_ctx *ctx, size_t bytes, void **out);   /* pinned, 256-byte aligned */
int bn_host_free(bn_ctx *ctx, void *ptr);
int bn_copy_h2d(bn_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);   /* synchronous */
int bn_copy_d2h(bn_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);   /* synchronous */

/* ------------------------------------------------------------------ host-pointer calls ------ */

/* bitnuc::encode / encode_alloc (src/utils/mod.rs:22-25,38-42 -> src/utils/packing/avx.rs:130-151).
 * out needs ceil(n/32) words.  On BN_INVALID_BASE, *n_words = offset/32 = the words the reference
 * leaves in ebuf (those of the chunks before the failing chunk); out[0..*n_words) is valid.
 * n == 0 -> BN_ERR_EMPTY_ENCODE. */
int bn_encode(bn_ctx *ctx, const uint8_t *seq, size_t n, uint64_t *out, size_t *n_words, bn_error_t *err);

/* bitnuc::decode (src/utils/mod.rs:60-62 -> src/utils/unpacking/avx.rs:117-153).  Writes n_bases
 * bytes at out (the binding appends them to the caller's Vec).  n_words < ceil(n_bases/32) ->
 * BN_INVALID_LENGTH(n_bases) (src/utils/unpacking/mod.rs:42-45); n_bases == 0 writes nothing. */
int bn_decode(bn_ctx *ctx, const uint64_t *words, size_t n_words, size_t n_bases, uint8_t *out, bn_error_t *err);

/* bitnuc::as_2bit over n records (src/utils/packing/mod.rs:81-110; the caller's `for kmer in ..`
 * loop with `?`, README.md:52-56).  Record r = recs[r*stride .. r*stride+k).  k > 32 ->
 * BN_SEQUENCE_TOO_LONG(k) before any content is looked at; otherwise the first failing record in
 * index order reports BN_INVALID_BASE (err->record, err->offset = byte offset inside recs).
 * k == 0 gives 0 for every record. stride >= k. */
int bn_as_2bit_batch(bn_ctx *ctx, const uint8_t *recs, size_t n, uint32_t k, size_t stride, uint64_t *out, bn_error_t *err);

/* bitnuc::from_2bit over n words (src/utils/unpacking/mod.rs:119-147).  Writes the low k bases of
 * packed[r] at out[r*stride .. r*stride+k); bytes between records are not touched.  k > 32 ->
 * BN_INVALID_LENGTH(k). */
int bn_from_2bit_batch(bn_ctx *ctx, const uint64_t *packed, size_t n, uint32_t k, uint8_t *out, size_t stride, bn_error_t *err);

/* bitnuc::hdist (src/utils/functions/hamming/multi.rs:122-160).  *total is the exact count; the
 * reference's u32 result is (uint32_t)*total (its accumulator wraps in release builds, :130).
 * n_words_a or n_words_b < ceil(n_bases/32) -> BN_INVALID_LENGTH(n_bases). */
int bn_hdist(bn_ctx *ctx, const uint64_t *a, size_t n_words_a, const uint64_t *b, size_t n_words_b, size_t n_bases, uint64_t *total, bn_error_t *err);

/* bitnuc::hdist_scalar over n pairs (src/utils/functions/hamming/scalar.rs:11-48):
 * out[i] = mismatches among the low len bases of (u[i], v[i]).  len > 32 -> BN_INVALID_LENGTH(len). */
int bn_hdist_pairs(bn_ctx *ctx, const uint64_t *u, const uint64_t *v, size_t n_pairs, uint32_t len, uint32_t *out, bn_error_t *err);

/* BaseCount::base_counts + GCContent::gc_content of one packed sequence
 * (src/utils/analysis.rs:19-39, :3-17).  counts = [A,C,G,T] over the n_bases valid bases only
 * (tail padding is not counted as A); *gc = (gc as f64 / len as f64) * 100.0, 0.0 when empty.
 * n_words < ceil(n_bases/32) -> BN_INVALID_LENGTH(n_bases).  gc may be NULL. */
int bn_base_counts(bn_ctx *ctx, const uint64_t *words, size_t n_words, size_t n_bases, uint64_t counts[4], double *gc, bn_error_t *err);

/* The same per read for a batch of packed reads: read r = words[word_offsets[r] ..) holding lens[r]
 * bases.  counts4 = n_reads x [A,C,G,T] (may be NULL), gc = n_reads doubles (may be NULL),
 * totals[4] = sum over reads (may be NULL). */
int bn_base_counts_batch(bn_ctx *ctx, const uint64_t *words, size_t n_words, const uint64_t *word_offsets, const uint64_t *lens, size_t n_reads, uint64_t *counts4, double *gc, uint64_t totals[4], bn_error_t *err);

/* PackedSequence::new over a batch of variable-length reads (src/sequence.rs:40-52 -> encode):
 * read r = bytes[offsets[r] .. offsets[r+1]); every read starts on a fresh word.
 * out_word_offsets[n_reads+1] receives the exclusive prefix sum of ceil(len/32) (empty reads take
 * no words, src/sequence.rs:42-46); out_words needs out_word_offsets[n_reads] words (an upper bound
 * is (offsets[n_reads]-offsets[0])/32 + n_reads).  The first invalid base in input order reports
 * BN_INVALID_BASE with err->record = read index, err->b = position in the read, err->offset = byte
 * offset in bytes.  read_status (may be NULL) receives per read the position of its first invalid
 * base or UINT32_MAX when the read is clean (the non-short-circuit variant). */
int bn_encode_batch(bn_ctx *ctx, const uint8_t *bytes, const uint64_t *offsets, size_t n_reads, uint64_t *out_words, uint64_t *out_word_offsets, uint32_t *read_status, bn_error_t *err);

/* bitnuc::split_packed over a batch of packed reads (src/utils/functions/split.rs:14-102): read r
 * (lens[r] bases, ebuf = words[word_offsets[r] .. word_offsets[r+1]); word_offsets has n_reads+1 entries, e.g.
 * the out_word_offsets of bn_encode_batch) is split at base idx[r].  left needs n_words + n_reads words,
 * right needs n_words words; left_offsets / right_offsets [n_reads+1] receive the exclusive prefix sums of
 * the per-read word counts (left: 0, ebuf.len() or idx/32 + 1; right: ebuf.len(), 0 or ebuf.len() - idx/32).
 * The reference's carry order is reproduced bit for bit (see csrc/split.cu).
 * idx[r] > lens[r] -> BN_INDEX_OUT_OF_BOUNDS{idx, len} for the first such read (err->record); a non-empty
 * ebuf shorter than ceil(len/32) words (the reference panics or truncates) -> BN_INVALID_LENGTH(len). */
int bn_split_packed_batch(bn_ctx *ctx, const uint64_t *words, size_t n_words, const uint64_t *word_offsets, const uint64_t *lens, const uint64_t *idx, size_t n_reads, uint64_t *left, uint64_t *left_offsets, uint64_t *right, uint64_t *right_offsets, bn_error_t *err);

/* Every k-mer of a sequence: `for w in seq.windows(k) { as_2bit(w)? }` (README.md:160-180 -> src/utils/packing/mod.rs:81-110).
 * out[i] = as_2bit(seq[i .. i+k]) for i in [0, n-k]; *n_out = n - k + 1 (0 when n < k: no window exists and nothing is
 * looked at).  k > 32 (and n >= k) -> BN_SEQUENCE_TOO_LONG(k); the first window holding an invalid byte ->
 * BN_INVALID_BASE (err->offset = the byte's offset, err->record = the window, max(0, offset - k + 1)); k == 0
 * (slice::windows panics) -> BN_ERR_ARGUMENT. */
int bn_kmers(bn_ctx *ctx, const uint8_t *seq, size_t n, uint32_t k, uint64_t *out, size_t *n_out, bn_error_t *err);

/* The same per read of a batch (reads = bytes[offsets[r] .. offsets[r+1]), as in bn_encode_batch): windows never cross
 * a read, a read shorter than k has none (and its bytes are never looked at).  out receives the windows of all reads
 * back to back (out_cap words available), out_offsets[n_reads+1] the exclusive prefix sums of max(0, len - k + 1).
 * The first failing window in (read, position) order -> BN_INVALID_BASE (err->record = the read, err->b = the byte's
 * position inside it, err->offset = its offset in bytes); out_cap too small -> BN_ERR_ARGUMENT. */
int bn_kmers_batch(bn_ctx *ctx, const uint8_t *bytes, const uint64_t *offsets, size_t n_reads, uint32_t k, uint64_t *out, size_t out_cap, uint64_t *out_offsets, bn_error_t *err);

/* PackedSequence::slice over a batch of queries (src/sequence.rs:198-212): query q = bases [q_start[q], q_end[q]) of
 * read q_read[q] (lens[r] bases at words[word_offsets[r]]).  out receives the upper-case ASCII of all queries back to
 * back (out_cap bytes available), out_offsets[nq+1] the exclusive prefix sums of the range lengths.
 * start > end || end > len -> BN_INVALID_RANGE{start, end, len} for the first such query (err->record);
 * q_read[q] >= n_reads or out_cap too small -> BN_ERR_ARGUMENT. */
int bn_slice_batch(bn_ctx *ctx, const uint64_t *words, size_t n_words, const uint64_t *word_offsets, const uint64_t *lens, size_t n_reads, const uint64_t *q_read, const uint64_t *q_start, const uint64_t *q_end, size_t nq, uint8_t *out, size_t out_cap, uint64_t *out_offsets, bn_error_t *err);

/* PackedSequence::get over a batch of queries (src/sequence.rs:116-135): out[q] = base q_index[q] of read q_read[q].
 * index >= len -> BN_INDEX_OUT_OF_BOUNDS{index, len} for the first such query (err->record). */
int bn_get_batch(bn_ctx *ctx, const uint64_t *words, size_t n_words, const uint64_t *word_offsets, const uint64_t *lens, size_t n_reads, const uint64_t *q_read, const uint64_t *q_index, size_t nq, uint8_t *out, bn_error_t *err);

/* FASTQ text -> record offsets -> per-read packed words, parsed on the device (the reference has no
 * parser, README.md:160-180 shows the caller's loop `for record in reader { PackedSequence::new(record.seq())? }`
 * that this replaces, src/sequence.rs:40-52).  Strict four-line records: '@' header, sequence, '+' separator,
 * quality as long as the sequence; "\n" or "\r\n" line ends; the last newline may be missing.
 * bn_fastq_scan uploads the text, finds the records and returns the sizes the caller must allocate:
 * *n_reads, and *n_words = sum of ceil(len/32).  A malformed record -> BN_ERR_FASTQ (first in file order).
 * bn_fastq_encode (same text, right after the scan: the context still holds the upload and the index) fills
 * out_words[n_words], out_word_offsets[n_reads+1] (read r = out_words[out_word_offsets[r] ..), every read on a fresh
 * word), seq_offsets[n_reads] (byte offset of each sequence line in the text) and seq_lens[n_reads]; any of the last
 * three may be NULL.  The first byte outside ACGTacgt in file order -> BN_INVALID_BASE (err->record = the read,
 * err->b = its position inside the read, err->offset = its offset in the text). */
int bn_fastq_scan(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t *n_reads, size_t *n_words, bn_error_t *err);
int bn_fastq_encode(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t n_reads, size_t n_words, uint64_t *out_words, uint64_t *out_word_offsets, uint64_t *seq_offsets, uint64_t *seq_lens, bn_error_t *err);
/* The same for FASTA text with ONE sequence line per record (the form read processors emit: '>' header line, sequence
 * line): two-line records, same outputs, same errors (BN_FASTQ_BAD_HEADER / BN_FASTQ_TRUNCATED are the faults that can
 * occur).  A sequence wrapped over several lines is not this format (its second line is reported as a bad header): see
 * bn_fasta_wrapped_* below. */
int bn_fasta_scan(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t *n_reads, size_t *n_words, bn_error_t *err);
int bn_fasta_encode(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t n_reads, size_t n_words, uint64_t *out_words, uint64_t *out_word_offsets, uint64_t *seq_offsets, uint64_t *seq_lens, bn_error_t *err);
/* WRAPPED (multi-line) FASTA, the genome-file form: a '>' header line, then any number of sequence lines per record; a
 * record's sequence is the concatenation of its lines ("\n" or "\r\n", empty lines add nothing, the last newline may be
 * missing).  The caller's loop is the same `PackedSequence::new(record.seq())` (README.md:160-180, src/sequence.rs:40-52)
 * -- a FASTA reader hands it the joined sequence.  bn_fasta_wrapped_scan uploads the text and returns *n_records,
 * *n_bases (all sequence bytes) and *n_words = sum of ceil(len/32); a text that does not open with a header ->
 * BN_ERR_FASTQ (record 0, BN_FASTQ_BAD_HEADER); the first byte outside ACGTacgt in file order -> BN_INVALID_BASE
 * (err->record = the record, err->b = its position inside the record's sequence, err->offset = its position in the
 * concatenation of all sequences).  bn_fasta_wrapped_encode (same text, right after the scan) fills out_words[n_words],
 * out_word_offsets[n_records+1], header_offsets[n_records] (byte offset of every header line in the text) and
 * seq_lens[n_records]; any of the last three may be NULL. */
int bn_fasta_wrapped_scan(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t *n_records, size_t *n_bases, size_t *n_words, bn_error_t *err);
int bn_fasta_wrapped_encode(bn_ctx *ctx, const uint8_t *text, size_t n_bytes, size_t n_records, size_t n_words, uint64_t *out_words, uint64_t *out_word_offsets, uint64_t *header_offsets, uint64_t *seq_lens, bn_error_t *err);

/* ------------------------------------------------------------------ device-pointer calls ---- */
/* All of these only enqueue work on `stream` (NULL = context stream) and never synchronise.
 * Pointers are device pointers.  ASCII and packed buffers should be 16-byte aligned: that selects the fast kernels
 * (misaligned pointers are served by slower byte- / word-granular kernels; packed buffers always need 8 bytes).
 * d_status is one device uint64_t that the call resets and the kernel updates with the smallest
 * (offset << 8 | byte) of any invalid base; read it back with bn_status_fetch. */

int bn_encode_dev(bn_ctx *ctx, void *stream, const uint8_t *d_seq, size_t n, uint64_t *d_out, uint64_t *d_status);
int bn_decode_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, size_t n_words, size_t n_bases, uint8_t *d_out);
int bn_as_2bit_batch_dev(bn_ctx *ctx, void *stream, const uint8_t *d_recs, size_t n, uint32_t k, size_t stride, uint64_t *d_out, uint64_t *d_status);
int bn_from_2bit_batch_dev(bn_ctx *ctx, void *stream, const uint64_t *d_packed, size_t n, uint32_t k, uint8_t *d_out, size_t stride);
/* *d_total (device uint64_t) is reset by the call, then accumulated. */
int bn_hdist_dev(bn_ctx *ctx, void *stream, const uint64_t *d_a, const uint64_t *d_b, size_t n_bases, uint64_t *d_total);
int bn_hdist_pairs_dev(bn_ctx *ctx, void *stream, const uint64_t *d_u, const uint64_t *d_v, size_t n_pairs, uint32_t len, uint32_t *d_out);
/* d_counts = 4 device uint64_t [A,C,G,T]; d_gc = device double or NULL. */
int bn_base_counts_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, size_t n_bases, uint64_t *d_counts, double *d_gc);
/* d_totals (4 device uint64_t, may be NULL) is reset, then accumulated over the reads.  n_words (the
 * size of d_words) only selects the thread-per-read or warp-per-read kernel; 0 = unknown. */
int bn_base_counts_batch_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, size_t n_words, const uint64_t *d_word_offsets, const uint64_t *d_lens, size_t n_reads, uint64_t *d_counts4, double *d_gc, uint64_t *d_totals);
/* Fixed-length reads (e.g. 10 M x 150 bp): read r = d_words[r*ceil(read_len/32) ..), read_len bases each.
 * Same outputs as bn_base_counts_batch_dev without the 16 bytes per read of offset/length arrays. */
int bn_base_counts_fixed_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, size_t n_reads, size_t read_len, uint64_t *d_counts4, double *d_gc, uint64_t *d_totals);
/* d_out_word_offsets[n_reads+1] is produced by a device scan.  n_bytes = the bytes the reads span
 * (d_offsets[n_reads] - d_offsets[0], or any upper bound such as the size of the byte buffer); d_out_words
 * needs n_bytes/32 + n_reads words and d_scratch bn_encode_batch_scratch_bytes(n_reads, n_bytes) bytes. */
size_t bn_encode_batch_scratch_bytes(size_t n_reads, size_t n_bytes);
int bn_encode_batch_dev(bn_ctx *ctx, void *stream, const uint8_t *d_bytes, const uint64_t *d_offsets, size_t n_reads, size_t n_bytes, uint64_t *d_out_words, uint64_t *d_out_word_offsets, uint32_t *d_read_status, uint64_t *d_status, void *d_scratch);

/* d_status (device uint64_t) receives min(read index << 1 | kind) over failing reads (kind 0: idx > len,
 * kind 1: ebuf too short), or UINT64_MAX; failing reads produce no output words.
 * d_scratch needs bn_split_packed_scratch_bytes(n_reads) bytes. */
size_t bn_split_packed_scratch_bytes(size_t n_reads);
int bn_split_packed_batch_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, const uint64_t *d_word_offsets, const uint64_t *d_lens, const uint64_t *d_idx, size_t n_reads, uint64_t *d_left, uint64_t *d_left_offsets, uint64_t *d_right, uint64_t *d_right_offsets, uint64_t *d_status, void *d_scratch);

/* n_bytes = the bytes the reads span (or an upper bound); d_out needs sum(max(0, len - k + 1)) <= n_bytes words,
 * d_scratch bn_kmers_batch_scratch_bytes(n_reads, n_bytes) bytes; status word as in bn_encode_batch_dev. */
size_t bn_kmers_batch_scratch_bytes(size_t n_reads, size_t n_bytes);
int bn_kmers_batch_dev(bn_ctx *ctx, void *stream, const uint8_t *d_bytes, const uint64_t *d_offsets, size_t n_reads, size_t n_bytes, uint32_t k, uint64_t *d_out, uint64_t *d_out_offsets, uint64_t *d_status, void *d_scratch);

/* d_status (device uint64_t) receives the smallest failing query index, or UINT64_MAX; failing queries produce no
 * bytes (slice) / a 0 byte (get).  d_out needs the sum of the valid range lengths; d_scratch needs
 * bn_slice_batch_scratch_bytes(nq) bytes. */
size_t bn_slice_batch_scratch_bytes(size_t nq);
int bn_slice_batch_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, const uint64_t *d_word_offsets, const uint64_t *d_lens, size_t n_reads, const uint64_t *d_q_read, const uint64_t *d_q_start, const uint64_t *d_q_end, size_t nq, uint8_t *d_out, uint64_t *d_out_offsets, uint64_t *d_status, void *d_scratch);
int bn_get_batch_dev(bn_ctx *ctx, void *stream, const uint64_t *d_words, const uint64_t *d_word_offsets, const uint64_t *d_lens, size_t n_reads, const uint64_t *d_q_read, const uint64_t *d_q_index, size_t nq, uint8_t *d_out, uint64_t *d_status);

/* FASTQ on the device, three enqueue-only steps with the two sizes read back by the caller in between.  d_text must
 * be 16-byte aligned.  d_scratch: bn_fastq_scratch_bytes(n_bytes) bytes, shared by the three calls.
 *   1. bn_fastq_count_dev: *d_n_lines (device uint64_t) = number of lines.  n_reads = n_lines / 4 (a remainder means
 *      the text ends inside a record: bn_fastq_status_fetch reports it).
 *   2. bn_fastq_index_dev: d_seq_offsets[n_reads], d_seq_lens[n_reads], d_word_offsets[n_reads+1] (the last entry is
 *      the number of output words); d_index_scratch: bn_fastq_index_scratch_bytes(n_reads) bytes, 16-byte aligned.
 *      d_status = two device uint64_t, reset here: [0] min(text offset << 8 | byte) over invalid bases (written by
 *      step 3), [1] min(record << 8 | bn_fastq_fault) over malformed records.
 *   3. bn_fastq_encode_dev: d_out_words[d_word_offsets[n_reads]].
 * bn_fastq_status_fetch synchronises and translates d_status (n_lines as read back after step 1): BN_ERR_FASTQ first,
 * then BN_INVALID_BASE with err->record / err->b resolved through d_seq_offsets. */
size_t bn_fastq_scratch_bytes(size_t n_bytes);
size_t bn_fastq_index_scratch_bytes(size_t n_reads);
int bn_fastq_count_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, void *d_scratch, uint64_t *d_n_lines);
int bn_fastq_index_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, size_t n_reads, void *d_scratch, void *d_index_scratch, uint64_t *d_seq_offsets, uint64_t *d_seq_lens, uint64_t *d_word_offsets, uint64_t *d_status);
int bn_fastq_encode_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, size_t n_reads, void *d_scratch, const uint64_t *d_seq_offsets, const uint64_t *d_seq_lens, const uint64_t *d_word_offsets, uint64_t *d_out_words, uint64_t *d_status);
int bn_fastq_status_fetch(bn_ctx *ctx, void *stream, const uint64_t *d_status, uint64_t n_lines, const uint64_t *d_seq_offsets, size_t n_reads, bn_error_t *err);
/* One-sequence-line FASTA on the device: the same three steps (same scratch sizes), n_reads = n_lines / 2. */
int bn_fasta_count_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, void *d_scratch, uint64_t *d_n_lines);
int bn_fasta_index_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, size_t n_reads, void *d_scratch, void *d_index_scratch, uint64_t *d_seq_offsets, uint64_t *d_seq_lens, uint64_t *d_word_offsets, uint64_t *d_status);
int bn_fasta_encode_dev(bn_ctx *ctx, void *stream, const uint8_t *d_text, size_t n_bytes, size_t n_reads, void *d_scratch, const uint64_t *d_seq_offsets, const uint64_t *d_seq_lens, const uint64_t *d_word_offsets, uint64_t *d_out_words, uint64_t *d_status);
int bn_fasta_status_fetch(bn_ctx *ctx, void *stream, const uint64_t *d_status, uint64_t n_lines, const uint64_t *d_seq_offsets, size_t n_reads, bn_error_t *err);

/* d_out needs n - k + 1 words; any alignment of d_seq. */
int bn_kmers_dev(bn_ctx *ctx, void *stream, const uint8_t *d_seq, size_t n, uint32_t k, uint64_t *d_out, uint64_t *d_status);

/* Synchronises `stream`, reads *d_status back and translates it: BN_OK, or BN_INVALID_BASE with
 * err->base / err->offset filled (record/a are filled by the host-pointer wrappers). */
int bn_status_fetch(bn_ctx *ctx, void *stream, const uint64_t *d_status, bn_error_t *err);

/* ------------------------------------------------------------------ multi-GPU (one process, N devices) ------ */
/* The path shards with no exchange step: every 32-base word, record, pair and read is independent
 * (src/utils/packing/avx.rs:138-145 carries nothing between words).  A bn_multi owns one bn_ctx per device (own
 * streams, own pinned staging) and one host worker thread per device; a bn_multi_* host-pointer call cuts its input
 * into contiguous shards -- base ranges on multiples of 64 bases (64 B of ASCII / 16 B packed, so every shard keeps
 * 128-bit accesses), records / pairs / reads by index, variable-length reads by byte volume on read boundaries -- and
 * runs the single-device call of the same name on every shard at once.  Results are those of the single-device call
 * on the whole input: the first failing shard in input order reports the error, with offsets and record indices
 * rebased to the caller's buffers.  The reference has nothing here (it is single-threaded).
 *
 * One collective exists on the path: the sum of the four base counters (and of the hdist total) over the shards.
 *   BN_REDUCE_NCCL  ncclAllReduce(4 x uint64, ncclSum) over per-device communicators from ncclCommInitAll, grouped
 *                   (libnccl.so.2 is resolved with dlopen at bn_multi_create: no link-time dependency);
 *   BN_REDUCE_P2P   our own all-reduce kernel over NVLink peer memory: every device stores its four counters into a
 *                   mailbox on every peer and sums the n mailboxes it received -- one 32-thread launch per device
 *                   instead of NCCL's, for a 32-byte message that is pure latency.  Also the mode for device lists
 *                   that name one device twice (NCCL refuses those), which is how a one-GPU box tests the sharding.
 * After a reduction every device holds the global value. */
typedef struct bn_multi bn_multi;
typedef enum bn_reduce { BN_REDUCE_NCCL = 0, BN_REDUCE_P2P = 1 } bn_reduce;

/* devs = n device ordinals (NULL: devices 0..n-1; n <= 0: every visible device). */
int bn_multi_create(const int *devs, int n, int reduce, bn_multi **out);
void bn_multi_destroy(bn_multi *m);
int bn_multi_size(const bn_multi *m);
bn_ctx *bn_multi_ctx(bn_multi *m, int i);            /* the context of shard i (device-resident calls, memory helpers) */
int bn_multi_reduce(const bn_multi *m);              /* bn_reduce in use */
int bn_multi_nccl_version(const bn_multi *m);        /* ncclGetVersion() of the library in use, 0 in P2P mode */
int bn_multi_set_chunk_bytes(bn_multi *m, size_t bytes);
int bn_multi_synchronize(bn_multi *m);               /* all streams of all devices; reports a collective that timed out */

/* The shard plan, for callers that place device-resident shards themselves: starts[n+1], shard i = units
 * [starts[i], starts[i+1]).  bn_multi_shard_units cuts n_units into near-equal ranges whose interior cuts are multiples
 * of `align` (bases: 64); bn_multi_shard_reads cuts an offset-indexed read batch into ranges of near-equal byte volume
 * on read boundaries (the first read whose start reaches the ideal cut). */
int bn_multi_shard_units(const bn_multi *m, size_t n_units, size_t align, size_t *starts);
int bn_multi_shard_reads(const bn_multi *m, const uint64_t *offsets, size_t n_reads, size_t *starts);

/* Host-pointer calls: the signatures and results of bn_encode ... bn_encode_batch with a bn_multi in place of the
 * bn_ctx.  bn_multi_base_counts[_batch] reduce the four counters with the collective above; bn_multi_hdist sums its
 * per-shard totals on the host (one uint64 per shard is already there). */
int bn_multi_encode(bn_multi *m, const uint8_t *seq, size_t n, uint64_t *out, size_t *n_words, bn_error_t *err);
int bn_multi_decode(bn_multi *m, const uint64_t *words, size_t n_words, size_t n_bases, uint8_t *out, bn_error_t *err);
int bn_multi_as_2bit_batch(bn_multi *m, const uint8_t *recs, size_t n, uint32_t k, size_t stride, uint64_t *out, bn_error_t *err);
int bn_multi_from_2bit_batch(bn_multi *m, const uint64_t *packed, size_t n, uint32_t k, uint8_t *out, size_t stride, bn_error_t *err);
int bn_multi_hdist(bn_multi *m, const uint64_t *a, size_t n_words_a, const uint64_t *b, size_t n_words_b, size_t n_bases, uint64_t *total, bn_error_t *err);
int bn_multi_hdist_pairs(bn_multi *m, const uint64_t *u, const uint64_t *v, size_t n_pairs, uint32_t len, uint32_t *out, bn_error_t *err);
int bn_multi_base_counts(bn_multi *m, const uint64_t *words, size_t n_words, size_t n_bases, uint64_t counts[4], double *gc, bn_error_t *err);
int bn_multi_base_counts_batch(bn_multi *m, const uint64_t *words, size_t n_words, const uint64_t *word_offsets, const uint64_t *lens, size_t n_reads, uint64_t *counts4, double *gc, uint64_t totals[4], bn_error_t *err);
int bn_multi_encode_batch(bn_multi *m, const uint8_t *bytes, const uint64_t *offsets, size_t n_reads, uint64_t *out_words, uint64_t *out_word_offsets, uint32_t *read_status, bn_error_t *err);

/* Device-resident sharded reductions: shard i lives on the device of bn_multi_ctx(m, i); every argument is an array of
 * n per-shard values.  Enqueue-only on each context's own stream (order other work against it with bn_ctx_stream),
 * all devices are launched before anything waits, and the collective follows on the same streams: when the streams
 * drain, d_counts[i] / d_totals[i] / d_total[i] on EVERY device hold the global [A,C,G,T] / total and d_gc[i] (array or
 * entries may be NULL) the global gc_content from those counts, (gc as f64 / len as f64) * 100.0 (analysis.rs:14).
 *   bn_multi_base_counts_dev        one long packed sequence, shard i = n_bases[i] bases at d_words[i]
 *   bn_multi_base_counts_fixed_dev  n_reads[i] reads of read_len bases each at d_words[i]; per-read d_counts4[i] /
 *                                   d_gc_reads[i] (arrays or entries may be NULL) stay local to their shard
 *   bn_multi_hdist_dev              whole-sequence mismatch total of shard pairs (d_a[i], d_b[i]) */
int bn_multi_base_counts_dev(bn_multi *m, const uint64_t *const *d_words, const size_t *n_bases, uint64_t *const *d_counts, double *const *d_gc);
int bn_multi_base_counts_fixed_dev(bn_multi *m, const uint64_t *const *d_words, const size_t *n_reads, size_t read_len, uint64_t *const *d_counts4, double *const *d_gc_reads, uint64_t *const *d_totals, double *const *d_gc);
int bn_multi_hdist_dev(bn_multi *m, const uint64_t *const *d_a, const uint64_t *const *d_b, const size_t *n_bases, uint64_t *const *d_total);
/* The collective alone, in place on n per-device buffers of count <= 4 uint64_t each (sum). */
int bn_multi_allreduce_u64_dev(bn_multi *m, uint64_t *const *d_buf, int count);
/* Device time of the last bn_multi_*_dev call per shard: ms[i] = first launch to end of the collective on device i's
 * stream (CUDA events).  The job's time is the maximum over i.  Synchronises those streams. */
int bn_multi_last_ms(bn_multi *m, float *ms);

/* ------------------------------------------------------------------ synthetic input --------- */
/* Counter-based generator used by the benchmarks and parity tests: word j of
 * stream s is splitmix64((seed ^ s*0x9E3779B97F4A7C15) + j); base 32j+i = "ACGT"[(W >> 2i) & 3].
 * bn_synth_words_dev writes words [first_word, first_word+n_words); bn_synth_ascii_dev writes the
 * ASCII of bases [first_base, first_base+n) with first_base a multiple of 32. */
int bn_synth_words_dev(bn_ctx *ctx, void *stream, uint64_t seed, uint64_t stream_id, uint64_t first_word, size_t n_words, uint64_t *d_out);
int bn_synth_ascii_dev(bn_ctx *ctx, void *stream, uint64_t seed, uint64_t stream_id, uint64_t first_base, size_t n, uint8_t *d_out);

#ifdef __cplusplus
}
#endif
#endif /* BITNUC_CUDA_H */
