/*
 * bitnuc_oracle.h -- CPU restatement of the bitnuc hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle: a plain-C restatement of the reference's algorithms, checked against
 * every known-answer vector in the reference's own unit tests (tests/golden/reference_kats.json).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under bitnuc_b200/ links, imports or calls it; the product path is CUDA-only.
 *
 * The reference is Rust and cannot be compiled in this image (no cargo/rustc), so there is no
 * oracle/_ref build.  Parity is pinned through the reference's own test vectors instead.
 *
 * All citations are relative to /root/reference/.
 */
#ifndef BITNUC_ORACLE_H
#define BITNUC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NucleotideError variants in declaration order (src/error.rs:4-18). 0 = Ok. */
enum {
    ORC_OK = 0,
    ORC_INVALID_BASE = 1,       /* a = offending byte                       */
    ORC_SEQUENCE_TOO_LONG = 2,  /* a = length                               */
    ORC_INVALID_LENGTH = 3,     /* a = length                               */
    ORC_INDEX_OUT_OF_BOUNDS = 4,/* a = index, b = length                    */
    ORC_INVALID_RANGE = 5,      /* a = start, b = end, c = length           */
    ORC_UNSUPPORTED = 6,
    ORC_PANIC = -100            /* the reference would panic on this input  */
};

typedef struct {
    int32_t code;
    uint64_t a, b, c;
} orc_error;

/* which of the reference's two x86 code paths to follow where their edge behaviour differs */
enum { ORC_PATH_NAIVE = 0, ORC_PATH_AVX2 = 1 };

/* src/error.rs:20-45 -- Display strings. Returns bytes written (excluding NUL). */
int orc_error_string(const orc_error *e, char *buf, size_t cap);

/* src/utils/packing/naive.rs:4-20 (semantics) ; avx.rs:76-128 has identical results */
int orc_as_2bit(const uint8_t *seq, size_t len, uint64_t *out, orc_error *err);
/* AVX2 restatement of src/utils/packing/avx.rs:76-128, used as the timed CPU baseline */
int orc_as_2bit_avx2(const uint8_t *seq, size_t len, uint64_t *out, orc_error *err);

/* src/utils/packing/naive.rs:22-43 == avx.rs:130-151. ebuf needs ceil(len/32) words.
 * *n_words = words pushed (on error: words of the chunks before the failing chunk).
 * len == 0 -> ORC_PANIC (0..n_chunks-1 underflow, avx.rs:138). */
int orc_encode(const uint8_t *seq, size_t len, uint64_t *ebuf, size_t *n_words, orc_error *err);
int orc_encode_avx2(const uint8_t *seq, size_t len, uint64_t *ebuf, size_t *n_words, orc_error *err);

/* src/utils/unpacking/naive.rs:3-25 ; avx.rs:50-114 gives the same bytes.
 * Writes expected_size bytes at out (the caller models Vec append). */
int orc_from_2bit(uint64_t packed, size_t expected_size, uint8_t *out, orc_error *err);
int orc_from_2bit_avx2(uint64_t packed, size_t expected_size, uint8_t *out, orc_error *err);

/* decode == from_2bit_multi: src/utils/unpacking/mod.rs:10-48 (naive path) and
 * src/utils/unpacking/avx.rs:117-153 (AVX2 path).  out needs room for
 * 32*max(n_words, ceil(n_bases/32)) bytes; *n_out = bytes appended. */
int orc_decode(const uint64_t *ebuf, size_t n_words, size_t n_bases, uint8_t *out, size_t *n_out,
               int path, orc_error *err);

/* src/utils/functions/hamming/scalar.rs:11-48 */
int orc_hdist_scalar(uint64_t u, uint64_t v, size_t len, uint32_t *out, orc_error *err);
/* src/utils/functions/hamming/multi.rs:122-160; u32 accumulator wraps as in a release build.
 * *total64 (optional) receives the unwrapped count. path selects :12-67 or the scalar loop. */
int orc_hdist(const uint64_t *e1, size_t n1, const uint64_t *e2, size_t n2, size_t n_bases,
              uint32_t *out, uint64_t *total64, int path, orc_error *err);

/* PackedSequence (src/sequence.rs) as (data, n_words, length) */
int orc_seq_get(const uint64_t *data, size_t length, size_t index, uint8_t *out, orc_error *err);   /* :116-135 */
int orc_seq_slice(const uint64_t *data, size_t length, size_t start, size_t end, uint8_t *out,
                  orc_error *err);                                                                   /* :198-212 */
/* src/utils/analysis.rs:19-39 and :3-17 (decode every base with get(), then count bytes) */
void orc_base_counts(const uint64_t *data, size_t length, uint64_t counts[4]);
double orc_gc_content(const uint64_t *data, size_t length);

/* src/utils/functions/split.rs:14-102.  lbuf/rbuf need n_words+1 words each. */
int orc_split_packed(const uint64_t *ebuf, size_t n_words, size_t slen, size_t idx, uint64_t *lbuf,
                     size_t *n_left, uint64_t *rbuf, size_t *n_right, orc_error *err);

/* ---- FASTQ record scanning ---------------------------------------------------
 * The reference has no parser (README.md:160-180 shows the caller's loop over a FASTQ reader); this is the
 * definition the CUDA path is held to: strict four-line records ('@' header, sequence, '+' separator, quality as long
 * as the sequence), "\n" or "\r\n" line ends, the last newline may be missing.  Fills starts[r] / lens[r] (byte offset
 * and length of every sequence line; cap entries available) and *n_reads.  Returns 0, or -5 with *bad_record and
 * *fault (1 header, 2 separator, 3 quality length, 4 text ends inside the record): the first fault in file order. */
int orc_fastq_scan(const uint8_t *text, size_t n, uint64_t *starts, uint64_t *lens, size_t cap, size_t *n_reads,
                   uint64_t *bad_record, int *fault);

/* The same for FASTA text with one sequence line per record ('>' header line, sequence line): faults 1 (header) and 4
 * (text ends inside the record).  A sequence wrapped over several lines is not this format. */
int orc_fasta_scan(const uint8_t *text, size_t n, uint64_t *starts, uint64_t *lens, size_t cap, size_t *n_reads,
                   uint64_t *bad_record, int *fault);

/* The caller's loop over a FASTQ / FASTA reader (README.md:160-180) in one call, for the timed CPU baseline of that row:
 * scan the records, then PackedSequence::new(record.seq()) = the AVX2 (or naive) encode of every sequence line onto
 * fresh words.  words needs sum(ceil(len/32)) entries (words_cap available), word_offsets n_reads + 1.  Returns 0,
 * -5 (format fault, as orc_fastq_scan), 1 (InvalidBase, err filled), or -2 when a buffer is too small. */
int orc_fastx_encode(const uint8_t *text, size_t n, int fasta, int path, uint64_t *words, size_t words_cap,
                     uint64_t *word_offsets, size_t reads_cap, size_t *n_reads, uint64_t *bad_record, int *fault,
                     orc_error *err);

/* ---- synthetic input: counter-based splitmix64 stream --------------------- */
uint64_t orc_splitmix64(uint64_t x);
uint64_t orc_synth_word(uint64_t seed, uint64_t stream, uint64_t j);
void orc_synth_ascii(uint64_t seed, uint64_t stream, uint64_t first_base, size_t n, uint8_t *out);

/* ---- timed CPU baseline (bench.py only) ---------------------------------------------------
 * Runs encode (+ decode when do_decode) over seq[0..n) split on 32-base boundaries across
 * n_threads pthreads, each thread calling the single-threaded reference restatement on its slice
 * (the chunking belongs to the harness; the reference is single-threaded).  Returns wall seconds
 * of the best of `reps` repetitions, or <0 on error.  path = ORC_PATH_AVX2 | ORC_PATH_NAIVE. */
double orc_bench_codec(const uint8_t *seq, size_t n, int n_threads, int reps, int path,
                       int do_encode, int do_decode, uint64_t *ebuf, uint8_t *dbuf);
int orc_have_avx2(void);

/* ---- timed CPU baselines for every row (bench.py only) -----------------------------------------
 * One op over n_units units split across n_threads threads created once (pinned to the CPUs of the
 * process's affinity mask when pin != 0) that meet at a barrier around each of `reps` repetitions;
 * times[r] = wall seconds of repetition r; *check = a checksum of the outputs (the hdist total, the sum of
 * the pair distances, the G+C total ...) so the work cannot be optimised away and can be compared.
 * The chunking is the harness's: the reference is single-threaded (src/lib.rs has no threads).
 *   ORC_OP_ENCODE / DECODE   units = bases (in0 ASCII -> out0 words / in0 words -> out0 ASCII), cut on 128-base bounds
 *   ORC_OP_AS_2BIT           units = records of k bytes at in0 + r*stride -> out0[r]      (packing/avx.rs:76-128 | naive.rs:4-20)
 *   ORC_OP_FROM_2BIT         units = words in0[r] -> k bytes at out0 + r*stride           (unpacking/avx.rs:50-114 | naive.rs:3-25)
 *   ORC_OP_HDIST             units = bases of the packed sequences in0, in1               (hamming/multi.rs:122-160, :12-67)
 *   ORC_OP_HDIST_PAIRS       units = pairs (in0[r], in1[r]), len k -> out0[r] (u32)       (hamming/scalar.rs:11-48)
 *   ORC_OP_BASE_COUNTS_GC    units = reads of k bases, ceil(k/32) words each at in0 -> out0[4r..] (u64), out1[r] (f64):
 *                            base_counts() then gc_content(), each with its own to_vec()  (analysis.rs:3-39, sequence.rs:198-212)
 *   ORC_OP_ENCODE_BATCH      units = reads: bytes in0, byte offsets in1[n+1], word offsets out1[n+1] (given by the caller) ->
 *                            words out0; one PackedSequence::new per read (sequence.rs:40-52), threads cut by byte volume
 * Returns 0, a NucleotideError code from the data, or <0 for a bad request. */
enum { ORC_OP_ENCODE = 0, ORC_OP_DECODE = 1, ORC_OP_AS_2BIT = 2, ORC_OP_FROM_2BIT = 3, ORC_OP_HDIST = 4,
       ORC_OP_HDIST_PAIRS = 5, ORC_OP_BASE_COUNTS_GC = 6, ORC_OP_ENCODE_BATCH = 7 };
typedef struct {
    int32_t op, path;
    uint64_t n_units, k, stride;
    const void *in0, *in1;
    void *out0, *out1;
} orc_bench_desc;
int orc_bench_op(const orc_bench_desc *d, int n_threads, int reps, int pin, double *times, uint64_t *check);

#ifdef __cplusplus
}
#endif
#endif
