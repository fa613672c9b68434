/*
 * bitnuc_oracle.c -- CPU restatement of the bitnuc hot path.  TEST INFRASTRUCTURE ONLY.
 * See bitnuc_oracle.h for the rules on who may load this.  Citations: /root/reference/<path>:<line>.
 *
 * Two layers:
 *   1. scalar restatements that define the semantics (naive.rs / scalar.rs / sequence.rs /
 *      analysis.rs), plus a `path` switch for the places where the reference's x86 AVX2 path and
 *      its naive path disagree on edge behaviour;
 *   2. AVX2-intrinsic restatements of packing/avx.rs, unpacking/avx.rs and hamming/multi.rs that
 *      keep the reference's structure (scalar validation scan, 16-lane steps, scalar bit-pack loop,
 *      32-iteration index extraction, per-word append).  They exist to be TIMED as the CPU
 *      baseline; tests check they agree with layer 1.
 */
#define _GNU_SOURCE
#include "bitnuc_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static int fail(orc_error *err, int code, uint64_t a, uint64_t b, uint64_t c) {
    if (err) {
        err->code = code;
        err->a = a;
        err->b = b;
        err->c = c;
    }
    return code;
}

static int ok(orc_error *err) { return fail(err, ORC_OK, 0, 0, 0); }

/* src/error.rs:20-45 */
int orc_error_string(const orc_error *e, char *buf, size_t cap) {
    switch (e->code) {
    case ORC_OK:
        return snprintf(buf, cap, "Ok");
    case ORC_INVALID_BASE: /* the byte is printed as a decimal integer (u8 Display) */
        return snprintf(buf, cap, "Invalid nucleotide base: %llu", (unsigned long long)e->a);
    case ORC_SEQUENCE_TOO_LONG:
        return snprintf(buf, cap, "Sequence length %llu exceeds maximum", (unsigned long long)e->a);
    case ORC_INVALID_LENGTH:
        return snprintf(buf, cap, "Invalid length: %llu", (unsigned long long)e->a);
    case ORC_INDEX_OUT_OF_BOUNDS:
        return snprintf(buf, cap, "Index %llu out of bounds for sequence of length %llu",
                        (unsigned long long)e->a, (unsigned long long)e->b);
    case ORC_INVALID_RANGE:
        return snprintf(buf, cap, "Invalid range %llu..%llu for sequence of length %llu",
                        (unsigned long long)e->a, (unsigned long long)e->b, (unsigned long long)e->c);
    case ORC_UNSUPPORTED:
        return snprintf(buf, cap, "Unsupported architecture");
    default:
        return snprintf(buf, cap, "panic");
    }
}

/* ------------------------------------------------------------------ packing (scalar) -------- */

/* match arm of src/utils/packing/naive.rs:10-16; returns 0..3 or -1 */
static inline int base_code(uint8_t b) {
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* src/utils/packing/naive.rs:4-20 */
int orc_as_2bit(const uint8_t *seq, size_t len, uint64_t *out, orc_error *err) {
    if (len > 32) return fail(err, ORC_SEQUENCE_TOO_LONG, len, 0, 0);
    uint64_t packed = 0;
    for (size_t i = 0; i < len; ++i) {
        int bits = base_code(seq[i]);
        if (bits < 0) return fail(err, ORC_INVALID_BASE, seq[i], 0, 0);
        packed |= (uint64_t)bits << (i * 2);
    }
    *out = packed;
    return ok(err);
}

/* src/utils/packing/naive.rs:22-43 */
int orc_encode(const uint8_t *seq, size_t len, uint64_t *ebuf, size_t *n_words, orc_error *err) {
    *n_words = 0; /* ebuf.clear() */
    size_t n_chunks = (len + 31) / 32;
    if (n_chunks == 0) return fail(err, ORC_PANIC, 0, 0, 0); /* 0..n_chunks-1 underflows */
    size_t l = 0;
    for (size_t k = 0; k + 1 < n_chunks; ++k) {
        uint64_t bits;
        int rc = orc_as_2bit(seq + l, 32, &bits, err);
        if (rc) return rc;
        ebuf[(*n_words)++] = bits;
        l += 32;
    }
    uint64_t bits;
    int rc = orc_as_2bit(seq + l, len - l, &bits, err);
    if (rc) return rc;
    ebuf[(*n_words)++] = bits;
    return ok(err);
}

/* ------------------------------------------------------------------ unpacking (scalar) ------ */

/* src/utils/unpacking/naive.rs:3-25 */
int orc_from_2bit(uint64_t packed, size_t expected_size, uint8_t *out, orc_error *err) {
    static const uint8_t lut[4] = {'A', 'C', 'G', 'T'};
    if (expected_size > 32) return fail(err, ORC_INVALID_LENGTH, expected_size, 0, 0);
    for (size_t i = 0; i < expected_size; ++i) out[i] = lut[(packed >> (i * 2)) & 3];
    return ok(err);
}

int orc_decode(const uint64_t *ebuf, size_t n_words, size_t n_bases, uint8_t *out, size_t *n_out,
               int path, orc_error *err) {
    *n_out = 0;
    if (path == ORC_PATH_AVX2) {
        /* src/utils/unpacking/avx.rs:117-153 */
        size_t full_chunks = n_bases / 32;
        size_t take = full_chunks < n_words ? full_chunks : n_words; /* .take() never over-runs */
        for (size_t k = 0; k < take; ++k) {
            orc_from_2bit(ebuf[k], 32, out + *n_out, NULL);
            *n_out += 32;
        }
        size_t rem = n_bases % 32;
        if (rem > 0) {
            if (full_chunks >= n_words) return fail(err, ORC_PANIC, 0, 0, 0); /* ebuf[full_chunks] */
            orc_from_2bit(ebuf[full_chunks], rem, out + *n_out, NULL);
            *n_out += rem;
        }
        return ok(err);
    }
    /* src/utils/unpacking/mod.rs:29-47 */
    size_t n_chunks = (n_bases + 31) / 32;
    if (n_chunks == 0) return fail(err, ORC_PANIC, 0, 0, 0); /* n_chunks - 1 underflows (debug) */
    size_t rem = n_bases % 32 == 0 ? 32 : n_bases % 32;
    size_t take = n_chunks - 1 < n_words ? n_chunks - 1 : n_words;
    for (size_t k = 0; k < take; ++k) {
        orc_from_2bit(ebuf[k], 32, out + *n_out, NULL);
        *n_out += 32;
    }
    if (n_chunks - 1 >= n_words) return fail(err, ORC_INVALID_LENGTH, n_bases, 0, 0); /* .get() = None */
    orc_from_2bit(ebuf[n_chunks - 1], rem, out + *n_out, NULL);
    *n_out += rem;
    return ok(err);
}

/* ------------------------------------------------------------------ hamming (scalar) -------- */

#define LOWER_BITS 0x5555555555555555ull
#define UPPER_BITS 0xAAAAAAAAAAAAAAAAull

/* src/utils/functions/hamming/scalar.rs:11-48 */
int orc_hdist_scalar(uint64_t u, uint64_t v, size_t len, uint32_t *out, orc_error *err) {
    if (len > 32) return fail(err, ORC_INVALID_LENGTH, len, 0, 0);
    *out = 0;
    if (len == 0 || u == v) return ok(err);
    size_t valid_bits = len * 2;
    uint64_t mask = valid_bits == 64 ? ~0ull : ((1ull << valid_bits) - 1);
    uint64_t diff = (u ^ v) & mask;
    if (diff == 0) return ok(err);
    uint64_t lower = diff & LOWER_BITS & mask;
    uint64_t upper = (diff & UPPER_BITS & mask) >> 1;
    *out = (uint32_t)__builtin_popcountll(lower | upper);
    return ok(err);
}

static uint64_t hdist_full_words(const uint64_t *e1, const uint64_t *e2, size_t full_chunks) {
    uint64_t t = 0;
    for (size_t k = 0; k < full_chunks; ++k) {
        uint64_t d = e1[k] ^ e2[k];
        t += (uint64_t)__builtin_popcountll((d & LOWER_BITS) | ((d & UPPER_BITS) >> 1));
    }
    return t;
}

__attribute__((target("avx2,popcnt"))) static uint32_t
hdist_multi_avx2(const uint64_t *e1, const uint64_t *e2, size_t full_chunks);

/* src/utils/functions/hamming/multi.rs:122-160 */
int orc_hdist(const uint64_t *e1, size_t n1, const uint64_t *e2, size_t n2, size_t n_bases,
              uint32_t *out, uint64_t *total64, int path, orc_error *err) {
    size_t expected = (n_bases + 31) / 32;
    if (n1 < expected || n2 < expected) return fail(err, ORC_INVALID_LENGTH, n_bases, 0, 0);
    size_t full_chunks = n_bases / 32;
    uint32_t total = 0; /* `let mut total_dist = 0u32` -- wraps in a release build */
    if (path == ORC_PATH_AVX2 && orc_have_avx2() && full_chunks >= 4)
        total = hdist_multi_avx2(e1, e2, full_chunks);
    if (total == 0 && full_chunks > 0) { /* multi.rs:147-151: "SIMD not available" re-scan */
        for (size_t k = 0; k < full_chunks; ++k) {
            uint32_t d;
            orc_hdist_scalar(e1[k], e2[k], 32, &d, NULL);
            total += d;
        }
    }
    uint64_t wide = hdist_full_words(e1, e2, full_chunks);
    size_t rem = n_bases % 32;
    if (rem > 0) {
        uint32_t d;
        orc_hdist_scalar(e1[full_chunks], e2[full_chunks], rem, &d, NULL);
        total += d;
        wide += d;
    }
    *out = total;
    if (total64) *total64 = wide;
    return ok(err);
}

/* ------------------------------------------------------------------ PackedSequence ---------- */

/* src/sequence.rs:116-135 */
int orc_seq_get(const uint64_t *data, size_t length, size_t index, uint8_t *out, orc_error *err) {
    static const uint8_t lut[4] = {'A', 'C', 'G', 'T'};
    if (index >= length) return fail(err, ORC_INDEX_OUT_OF_BOUNDS, index, length, 0);
    size_t chunk_idx = index / 32, bit_idx = (index % 32) * 2;
    *out = lut[(data[chunk_idx] >> bit_idx) & 3];
    return ok(err);
}

/* src/sequence.rs:198-212 */
int orc_seq_slice(const uint64_t *data, size_t length, size_t start, size_t end, uint8_t *out,
                  orc_error *err) {
    if (start > end || end > length) return fail(err, ORC_INVALID_RANGE, start, end, length);
    for (size_t i = start; i < end; ++i) {
        int rc = orc_seq_get(data, length, i, out + (i - start), err);
        if (rc) return rc;
    }
    return ok(err);
}

/* src/utils/analysis.rs:19-39: to_vec() (= slice(0..len), per-base get) then a byte match loop */
void orc_base_counts(const uint64_t *data, size_t length, uint64_t counts[4]) {
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (size_t i = 0; i < length; ++i) {
        uint8_t b;
        orc_seq_get(data, length, i, &b, NULL);
        switch (b) {
        case 'A': counts[0]++; break;
        case 'C': counts[1]++; break;
        case 'G': counts[2]++; break;
        case 'T': counts[3]++; break;
        default: break;
        }
    }
}

/* src/utils/analysis.rs:3-17: (gc_count as f64 / len as f64) * 100.0, in exactly this order */
double orc_gc_content(const uint64_t *data, size_t length) {
    if (length == 0) return 0.0;
    size_t gc = 0;
    for (size_t i = 0; i < length; ++i) {
        uint8_t b;
        orc_seq_get(data, length, i, &b, NULL);
        if (b == 'G' || b == 'C') gc++;
    }
    volatile double q = (double)gc / (double)length; /* no contraction / reassociation */
    return q * 100.0;
}

/* ------------------------------------------------------------------ split_packed ------------ */

/* src/utils/functions/split.rs:14-102 */
int orc_split_packed(const uint64_t *ebuf, size_t n_words, size_t slen, size_t idx, uint64_t *lbuf,
                     size_t *n_left, uint64_t *rbuf, size_t *n_right, orc_error *err) {
    if (idx > slen) return fail(err, ORC_INDEX_OUT_OF_BOUNDS, idx, slen, 0);
    *n_left = *n_right = 0;
    if (idx == 0) {
        memcpy(rbuf, ebuf, n_words * 8);
        *n_right = n_words;
        return ok(err);
    }
    if (idx == slen) {
        memcpy(lbuf, ebuf, n_words * 8);
        *n_left = n_words;
        return ok(err);
    }
    if (n_words == 0) return ok(err);
    size_t right_chunks = (slen - idx + 31) / 32;
    size_t chunk_idx = idx / 32, bit_idx = (idx % 32) * 2;
    if (chunk_idx >= n_words) return fail(err, ORC_PANIC, 0, 0, 0); /* ebuf[chunk_idx] */
    if (chunk_idx > 0) {
        memcpy(lbuf, ebuf, chunk_idx * 8);
        *n_left = chunk_idx;
    }
    uint64_t split_mask = bit_idx == 0 ? 0 : ((1ull << bit_idx) - 1);
    lbuf[(*n_left)++] = ebuf[chunk_idx] & split_mask;
    uint64_t carry = 0;
    for (size_t k = chunk_idx; k < n_words; ++k) {
        uint64_t curr = ebuf[k];
        rbuf[(*n_right)++] = carry | (curr >> bit_idx);
        carry = bit_idx == 0 ? 0 : curr << (64 - bit_idx);
    }
    if (carry != 0 && *n_right < right_chunks) rbuf[(*n_right)++] = carry;
    return ok(err);
}

/* ------------------------------------------------------------------ synthetic input --------- */

/* FASTQ record scanning: one sequential walk over the lines, the way a reader would do it. */
int orc_fastq_scan(const uint8_t *text, size_t n, uint64_t *starts, uint64_t *lens, size_t cap, size_t *n_reads,
                   uint64_t *bad_record, int *fault) {
    size_t pos = 0, r = 0;
    *n_reads = 0;
    while (pos < n) {
        size_t ls[4], ll[4]; /* start and length (without "\n" / "\r\n") of the record's four lines */
        int have = 0;
        for (int k = 0; k < 4 && pos < n; ++k) {
            size_t e = pos;
            while (e < n && text[e] != '\n') ++e;
            size_t len = e - pos;
            if (len && text[e - 1] == '\r') --len;
            ls[k] = pos;
            ll[k] = len;
            pos = e < n ? e + 1 : n;
            have = k + 1;
        }
        /* faults in file order: the header's first byte, the separator's first byte, the quality length, the end */
        int f = 0;
        if (text[ls[0]] != '@') f = 1;
        else if (have >= 3 && (ls[2] >= n || text[ls[2]] != '+')) f = 2;
        else if (have == 4 && ll[3] != ll[1]) f = 3;
        else if (have < 4) f = 4;
        if (f) {
            *bad_record = r;
            *fault = f;
            return -5;
        }
        if (r < cap) {
            starts[r] = ls[1];
            lens[r] = ll[1];
        }
        ++r;
        *n_reads = r;
    }
    return 0;
}

/* FASTA with one sequence line per record: the same walk, two lines per record. */
int orc_fasta_scan(const uint8_t *text, size_t n, uint64_t *starts, uint64_t *lens, size_t cap, size_t *n_reads,
                   uint64_t *bad_record, int *fault) {
    size_t pos = 0, r = 0;
    *n_reads = 0;
    while (pos < n) {
        size_t ls[2], ll[2];
        int have = 0;
        for (int k = 0; k < 2 && pos < n; ++k) {
            size_t e = pos;
            while (e < n && text[e] != '\n') ++e;
            size_t len = e - pos;
            if (len && text[e - 1] == '\r') --len;
            ls[k] = pos;
            ll[k] = len;
            pos = e < n ? e + 1 : n;
            have = k + 1;
        }
        const int f = text[ls[0]] != '>' ? 1 : (have < 2 ? 4 : 0);
        if (f) {
            *bad_record = r;
            *fault = f;
            return -5;
        }
        if (r < cap) {
            starts[r] = ls[1];
            lens[r] = ll[1];
        }
        ++r;
        *n_reads = r;
    }
    return 0;
}

int orc_fastx_encode(const uint8_t *text, size_t n, int fasta, int path, uint64_t *words, size_t words_cap,
                     uint64_t *word_offsets, size_t reads_cap, size_t *n_reads, uint64_t *bad_record, int *fault,
                     orc_error *err) {
    /* the reader first (a streaming reader would interleave; the work is the same) */
    uint64_t *starts = (uint64_t *)malloc((reads_cap ? reads_cap : 1) * 2 * sizeof(uint64_t));
    if (!starts) return -2;
    uint64_t *lens = starts + (reads_cap ? reads_cap : 1);
    int rc = fasta ? orc_fasta_scan(text, n, starts, lens, reads_cap, n_reads, bad_record, fault)
                   : orc_fastq_scan(text, n, starts, lens, reads_cap, n_reads, bad_record, fault);
    if (rc == 0 && *n_reads > reads_cap) rc = -2;
    size_t w = 0;
    for (size_t r = 0; rc == 0 && r < *n_reads; ++r) {
        word_offsets[r] = w;
        const size_t need = (size_t)((lens[r] + 31) / 32);
        if (need == 0) continue; /* PackedSequence::new(b"") is Ok and empty, src/sequence.rs:42-46 */
        if (w + need > words_cap) {
            rc = -2;
            break;
        }
        size_t got = 0;
        rc = path == ORC_PATH_AVX2 && orc_have_avx2() ? orc_encode_avx2(text + starts[r], (size_t)lens[r], words + w, &got, err)
                                                      : orc_encode(text + starts[r], (size_t)lens[r], words + w, &got, err);
        if (rc) *bad_record = r;
        w += need;
    }
    if (rc == 0) word_offsets[*n_reads] = w;
    free(starts);
    return rc;
}

uint64_t orc_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* word j of stream s = splitmix64((seed ^ s*golden) + j) */
uint64_t orc_synth_word(uint64_t seed, uint64_t stream, uint64_t j) {
    return orc_splitmix64((seed ^ (stream * 0x9E3779B97F4A7C15ull)) + j);
}

void orc_synth_ascii(uint64_t seed, uint64_t stream, uint64_t first_base, size_t n, uint8_t *out) {
    static const uint8_t lut[4] = {'A', 'C', 'G', 'T'};
    for (size_t i = 0; i < n; ++i) {
        uint64_t b = first_base + i;
        uint64_t w = orc_synth_word(seed, stream, b / 32);
        out[i] = lut[(w >> (2 * (b % 32))) & 3];
    }
}

/* ================================================================== AVX2 restatements ======= */

int orc_have_avx2(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt"); }

/* src/utils/packing/avx.rs:34-74: three (upper|lower) compare masks, select 1/2/3 */
__attribute__((target("avx2"))) static inline __m256i classify32(__m256i chunk) {
    __m256i c = _mm256_or_si256(_mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('C')),
                                _mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('c')));
    __m256i g = _mm256_or_si256(_mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('G')),
                                _mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('g')));
    __m256i t = _mm256_or_si256(_mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('T')),
                                _mm256_cmpeq_epi8(chunk, _mm256_set1_epi8('t')));
    __m256i r = _mm256_setzero_si256();
    r = _mm256_or_si256(_mm256_and_si256(c, _mm256_set1_epi8(1)), _mm256_andnot_si256(c, r));
    r = _mm256_or_si256(_mm256_and_si256(g, _mm256_set1_epi8(2)), _mm256_andnot_si256(g, r));
    r = _mm256_or_si256(_mm256_and_si256(t, _mm256_set1_epi8(3)), _mm256_andnot_si256(t, r));
    return r;
}

/* src/utils/packing/avx.rs:76-128.  `avail` = bytes readable at seq (the reference's 32-byte load
 * at seq[chunk_idx..] over-reads its slice, avx.rs:103; here the load is bounced through a
 * zero-padded temporary when fewer than 32 bytes are readable, which does not change results). */
__attribute__((target("avx2"))) static int as_2bit_avx2_impl(const uint8_t *seq, size_t len,
                                                             size_t avail, uint64_t *out,
                                                             orc_error *err) {
    if (len > 32) return fail(err, ORC_SEQUENCE_TOO_LONG, len, 0, 0);
    if (len < 16) return orc_as_2bit(seq, len, out, err);
    for (size_t i = 0; i < len; ++i) /* scalar validation scan, avx.rs:86-91 */
        if (base_code(seq[i]) < 0) return fail(err, ORC_INVALID_BASE, seq[i], 0, 0);
    uint64_t packed = 0;
    size_t simd_len = len - (len % 16);
    for (size_t chunk_idx = 0; chunk_idx < simd_len; chunk_idx += 16) {
        __m256i chunk;
        if (chunk_idx + 32 <= avail) {
            chunk = _mm256_loadu_si256((const __m256i *)(seq + chunk_idx));
        } else {
            uint8_t pad[32] = {0};
            memcpy(pad, seq + chunk_idx, avail - chunk_idx);
            chunk = _mm256_loadu_si256((const __m256i *)pad);
        }
        uint8_t temp[32];
        _mm256_storeu_si256((__m256i *)temp, classify32(chunk));
        for (size_t i = 0; i < 16; ++i) /* scalar bit-pack loop, avx.rs:109-111 */
            packed |= (uint64_t)temp[i] << ((chunk_idx + i) * 2);
    }
    for (size_t i = simd_len; i < len; ++i) /* scalar tail, avx.rs:115-124 */
        packed |= (uint64_t)base_code(seq[i]) << (i * 2);
    *out = packed;
    return ok(err);
}

int orc_as_2bit_avx2(const uint8_t *seq, size_t len, uint64_t *out, orc_error *err) {
    return as_2bit_avx2_impl(seq, len, len, out, err);
}

/* src/utils/packing/avx.rs:130-151 */
__attribute__((target("avx2"))) int orc_encode_avx2(const uint8_t *seq, size_t len, uint64_t *ebuf,
                                                    size_t *n_words, orc_error *err) {
    *n_words = 0;
    size_t n_chunks = (len + 31) / 32;
    if (n_chunks == 0) return fail(err, ORC_PANIC, 0, 0, 0);
    size_t l = 0;
    for (size_t k = 0; k + 1 < n_chunks; ++k) {
        uint64_t bits;
        int rc = as_2bit_avx2_impl(seq + l, 32, len - l, &bits, err);
        if (rc) return rc;
        ebuf[(*n_words)++] = bits;
        l += 32;
    }
    uint64_t bits;
    int rc = as_2bit_avx2_impl(seq + l, len - l, len - l, &bits, err);
    if (rc) return rc;
    ebuf[(*n_words)++] = bits;
    return ok(err);
}

/* src/utils/unpacking/avx.rs:26-33: 32 scalar shift/mask steps, one 256-bit load, one vpshufb */
__attribute__((target("avx2"))) static inline __m256i unpack_32(uint64_t packed, __m256i lookup) {
    uint8_t idx[32];
    for (int i = 0; i < 32; ++i) idx[i] = (uint8_t)((packed >> (i * 2)) & 3);
    return _mm256_shuffle_epi8(lookup, _mm256_loadu_si256((const __m256i *)idx));
}

__attribute__((target("avx2"))) static inline __m256i acgt_lut(void) {
    return _mm256_setr_epi8('A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G',
                            'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C',
                            'G', 'T');
}

/* src/utils/unpacking/avx.rs:50-114, collapsed: every branch yields the low expected_size bases */
__attribute__((target("avx2"))) int orc_from_2bit_avx2(uint64_t packed, size_t expected_size,
                                                       uint8_t *out, orc_error *err) {
    if (expected_size > 32) return fail(err, ORC_INVALID_LENGTH, expected_size, 0, 0);
    uint8_t temp[32];
    _mm256_storeu_si256((__m256i *)temp, unpack_32(packed, acgt_lut()));
    memcpy(out, temp, expected_size);
    return ok(err);
}

/* src/utils/unpacking/avx.rs:117-153 (well-formed input only; edge cases live in orc_decode) */
__attribute__((target("avx2"))) static void decode_avx2(const uint64_t *ebuf, size_t n_bases,
                                                        uint8_t *out) {
    __m256i lookup = acgt_lut();
    size_t full_chunks = n_bases / 32;
    uint8_t temp[32];
    for (size_t k = 0; k < full_chunks; ++k) {
        _mm256_storeu_si256((__m256i *)temp, unpack_32(ebuf[k], lookup));
        memcpy(out + 32 * k, temp, 32); /* extend_from_slice(&temp) */
    }
    size_t rem = n_bases % 32;
    if (rem) {
        _mm256_storeu_si256((__m256i *)temp, unpack_32(ebuf[full_chunks], lookup));
        memcpy(out + 32 * full_chunks, temp, rem);
    }
}

/* src/utils/functions/hamming/multi.rs:12-67 */
__attribute__((target("avx2,popcnt"))) static uint32_t
hdist_multi_avx2(const uint64_t *e1, const uint64_t *e2, size_t full_chunks) {
    uint32_t total = 0;
    size_t quad = full_chunks / 4;
    __m256i lower = _mm256_set1_epi64x((long long)LOWER_BITS);
    __m256i upper = _mm256_set1_epi64x((long long)UPPER_BITS);
    for (size_t i = 0; i < quad; ++i) {
        __m256i u = _mm256_loadu_si256((const __m256i *)(e1 + 4 * i));
        __m256i v = _mm256_loadu_si256((const __m256i *)(e2 + 4 * i));
        __m256i diff = _mm256_xor_si256(u, v);
        if (_mm256_testz_si256(diff, diff) == 1) continue;
        __m256i comb = _mm256_or_si256(_mm256_and_si256(diff, lower),
                                       _mm256_srli_epi64(_mm256_and_si256(diff, upper), 1));
        total += (uint32_t)__builtin_popcountll((uint64_t)_mm256_extract_epi64(comb, 0)) +
                 (uint32_t)__builtin_popcountll((uint64_t)_mm256_extract_epi64(comb, 1)) +
                 (uint32_t)__builtin_popcountll((uint64_t)_mm256_extract_epi64(comb, 2)) +
                 (uint32_t)__builtin_popcountll((uint64_t)_mm256_extract_epi64(comb, 3));
    }
    for (size_t k = quad * 4; k < full_chunks; ++k) {
        uint32_t d = 0;
        orc_hdist_scalar(e1[k], e2[k], 32, &d, NULL); /* .unwrap_or(0) */
        total += d;
    }
    return total;
}

/* ================================================================== timed baseline harness == */

typedef struct {
    const uint8_t *seq;
    size_t n;
    uint64_t *ebuf;
    uint8_t *dbuf;
    int path, do_encode, do_decode, rc;
} codec_job;

static void *codec_worker(void *p) {
    codec_job *j = (codec_job *)p;
    j->rc = 0;
    if (j->n == 0) return NULL;
    size_t nw = 0;
    if (j->do_encode) {
        orc_error e;
        j->rc = (j->path == ORC_PATH_AVX2 ? orc_encode_avx2 : orc_encode)(j->seq, j->n, j->ebuf, &nw, &e);
        if (j->rc) return NULL;
    }
    if (j->do_decode) {
        if (j->path == ORC_PATH_AVX2) {
            decode_avx2(j->ebuf, j->n, j->dbuf);
        } else {
            size_t n_out;
            orc_error e;
            j->rc = orc_decode(j->ebuf, (j->n + 31) / 32, j->n, j->dbuf, &n_out, ORC_PATH_NAIVE, &e);
        }
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_bench_codec(const uint8_t *seq, size_t n, int n_threads, int reps, int path,
                       int do_encode, int do_decode, uint64_t *ebuf, uint8_t *dbuf) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    if (path == ORC_PATH_AVX2 && !orc_have_avx2()) return -2.0;
    codec_job jobs[1024];
    pthread_t tids[1024];
    size_t n_words = (n + 31) / 32;
    size_t words_per = (n_words + (size_t)n_threads - 1) / (size_t)n_threads;
    double best = -1.0;
    for (int r = 0; r < reps; ++r) {
        double t0 = now_s();
        for (int t = 0; t < n_threads; ++t) {
            size_t w0 = (size_t)t * words_per;
            if (w0 > n_words) w0 = n_words;
            size_t w1 = w0 + words_per < n_words ? w0 + words_per : n_words;
            size_t b0 = w0 * 32, b1 = w1 * 32 < n ? w1 * 32 : n;
            jobs[t] = (codec_job){seq + b0, b1 - b0, ebuf + w0, dbuf ? dbuf + b0 : NULL,
                                  path, do_encode, do_decode, 0};
            if (n_threads == 1)
                codec_worker(&jobs[t]);
            else
                pthread_create(&tids[t], NULL, codec_worker, &jobs[t]);
        }
        if (n_threads > 1)
            for (int t = 0; t < n_threads; ++t) pthread_join(tids[t], NULL);
        double dt = now_s() - t0;
        for (int t = 0; t < n_threads; ++t)
            if (jobs[t].rc) return -1.0;
        if (best < 0 || dt < best) best = dt;
    }
    return best;
}

/* ================================================================== timed baselines, all ops == */
/* bench.py's cpu_baseline legs for the rows other than the codec:
 * the caller's loop over the reference's per-item functions, chunked across pinned threads by the harness
 * (the reference itself is single-threaded).  Threads are created once and meet at a barrier around
 * every repetition, so thread start-up is never inside a timed repetition. */

/* src/sequence.rs:198-212 + :260-262 as the reference runs it inside analysis.rs: to_vec() allocates a
 * Vec::with_capacity(len) and pushes one checked get() per base. */
static uint8_t *to_vec_ref(const uint64_t *data, size_t length) {
    uint8_t *v = (uint8_t *)malloc(length ? length : 1);
    size_t n = 0;
    for (size_t i = 0; i < length; ++i) {
        uint8_t b;
        if (orc_seq_get(data, length, i, &b, NULL)) break; /* `?` */
        v[n++] = b;
    }
    return v;
}

/* src/utils/analysis.rs:19-39 with its own to_vec() */
static void base_counts_ref(const uint64_t *data, size_t length, uint64_t counts[4]) {
    uint8_t *seq = to_vec_ref(data, length);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (size_t i = 0; i < length; ++i) {
        switch (seq[i]) {
        case 'A': counts[0]++; break;
        case 'C': counts[1]++; break;
        case 'G': counts[2]++; break;
        case 'T': counts[3]++; break;
        default: continue;
        }
    }
    free(seq);
}

/* src/utils/analysis.rs:3-17 with its own (second) to_vec() */
static double gc_content_ref(const uint64_t *data, size_t length) {
    uint8_t *seq = to_vec_ref(data, length);
    double r = 0.0;
    if (length) {
        size_t gc = 0;
        for (size_t i = 0; i < length; ++i) gc += (seq[i] == 'G' || seq[i] == 'C');
        volatile double q = (double)gc / (double)length;
        r = q * 100.0;
    }
    free(seq);
    return r;
}

typedef struct {
    const orc_bench_desc *d;
    pthread_barrier_t *bar;
    int tid, n_threads, reps, cpu, rc;
    uint64_t check;
} op_job;

static void op_range(const orc_bench_desc *d, int tid, int n_threads, size_t align, size_t *u0, size_t *u1) {
    size_t n = d->n_units;
    size_t blocks = (n + align - 1) / align;
    size_t per = (blocks + (size_t)n_threads - 1) / (size_t)n_threads;
    size_t b0 = (size_t)tid * per, b1 = b0 + per;
    *u0 = b0 * align < n ? b0 * align : n;
    *u1 = b1 * align < n ? b1 * align : n;
}

__attribute__((target("avx2,popcnt"))) static int op_run_avx2(const orc_bench_desc *d, size_t u0, size_t u1, uint64_t *check) {
    orc_error e;
    switch (d->op) {
    case ORC_OP_AS_2BIT: {
        const uint8_t *recs = (const uint8_t *)d->in0;
        uint64_t *out = (uint64_t *)d->out0;
        size_t total = (d->n_units - 1) * d->stride + d->k; /* bytes readable in the record buffer */
        for (size_t r = u0; r < u1; ++r) {
            int rc = as_2bit_avx2_impl(recs + r * d->stride, d->k, total - r * d->stride, &out[r], &e);
            if (rc) return rc;
        }
        *check = u1 > u0 ? out[u1 - 1] : 0;
        return 0;
    }
    case ORC_OP_FROM_2BIT: {
        const uint64_t *in = (const uint64_t *)d->in0;
        uint8_t *out = (uint8_t *)d->out0;
        for (size_t r = u0; r < u1; ++r) {
            int rc = orc_from_2bit_avx2(in[r], d->k, out + r * d->stride, &e);
            if (rc) return rc;
        }
        *check = u1 > u0 ? out[(u1 - 1) * d->stride] : 0;
        return 0;
    }
    default:
        return -3;
    }
}

static int op_run(const orc_bench_desc *d, size_t u0, size_t u1, uint64_t *check) {
    orc_error e;
    *check = 0;
    if (u1 <= u0) return 0;
    switch (d->op) {
    case ORC_OP_ENCODE: {
        size_t nw;
        uint64_t *out = (uint64_t *)d->out0 + u0 / 32;
        int rc = (d->path == ORC_PATH_AVX2 ? orc_encode_avx2 : orc_encode)((const uint8_t *)d->in0 + u0, u1 - u0, out, &nw, &e);
        *check = out[nw ? nw - 1 : 0];
        return rc;
    }
    case ORC_OP_DECODE: {
        const uint64_t *in = (const uint64_t *)d->in0 + u0 / 32;
        uint8_t *out = (uint8_t *)d->out0 + u0;
        if (d->path == ORC_PATH_AVX2) {
            decode_avx2(in, u1 - u0, out);
        } else {
            size_t n_out;
            int rc = orc_decode(in, (u1 - u0 + 31) / 32, u1 - u0, out, &n_out, ORC_PATH_NAIVE, &e);
            if (rc) return rc;
        }
        *check = out[u1 - u0 - 1];
        return 0;
    }
    case ORC_OP_AS_2BIT:
        if (d->path == ORC_PATH_AVX2) return op_run_avx2(d, u0, u1, check);
        {
            const uint8_t *recs = (const uint8_t *)d->in0;
            uint64_t *out = (uint64_t *)d->out0;
            for (size_t r = u0; r < u1; ++r) {
                int rc = orc_as_2bit(recs + r * d->stride, d->k, &out[r], &e);
                if (rc) return rc;
            }
            *check = out[u1 - 1];
            return 0;
        }
    case ORC_OP_FROM_2BIT:
        if (d->path == ORC_PATH_AVX2) return op_run_avx2(d, u0, u1, check);
        {
            const uint64_t *in = (const uint64_t *)d->in0;
            uint8_t *out = (uint8_t *)d->out0;
            for (size_t r = u0; r < u1; ++r) {
                int rc = orc_from_2bit(in[r], d->k, out + r * d->stride, &e);
                if (rc) return rc;
            }
            *check = out[(u1 - 1) * d->stride];
            return 0;
        }
    case ORC_OP_HDIST: { /* units = bases; the thread's slice is a whole-sequence hdist call of its own */
        uint32_t t32;
        uint64_t t64;
        size_t nw = (u1 - u0 + 31) / 32;
        int rc = orc_hdist((const uint64_t *)d->in0 + u0 / 32, nw, (const uint64_t *)d->in1 + u0 / 32, nw, u1 - u0, &t32, &t64,
                           d->path, &e);
        *check = t64;
        return rc;
    }
    case ORC_OP_HDIST_PAIRS: {
        const uint64_t *u = (const uint64_t *)d->in0, *v = (const uint64_t *)d->in1;
        uint32_t *out = (uint32_t *)d->out0;
        uint64_t sum = 0;
        for (size_t r = u0; r < u1; ++r) {
            int rc = orc_hdist_scalar(u[r], v[r], d->k, &out[r], &e);
            if (rc) return rc;
            sum += out[r];
        }
        *check = sum;
        return 0;
    }
    case ORC_OP_ENCODE_BATCH: { /* the caller's loop of PackedSequence::new(read) (sequence.rs:40-52): encode onto fresh words */
        const uint8_t *bytes = (const uint8_t *)d->in0;
        const uint64_t *off = (const uint64_t *)d->in1, *woff = (const uint64_t *)d->out1;
        uint64_t *words = (uint64_t *)d->out0;
        uint64_t sum = 0;
        for (size_t r = u0; r < u1; ++r) {
            size_t len = off[r + 1] - off[r], nw = 0;
            if (len == 0) continue; /* PackedSequence::new(b"") -> empty data, no encode call (sequence.rs:42-46) */
            int rc = (d->path == ORC_PATH_AVX2 ? orc_encode_avx2 : orc_encode)(bytes + off[r], len, words + woff[r], &nw, &e);
            if (rc) return rc;
            sum += nw;
        }
        *check = sum;
        return 0;
    }
    case ORC_OP_BASE_COUNTS_GC: { /* fixed-length reads, each its own PackedSequence of ceil(k/32) words */
        const uint64_t *w = (const uint64_t *)d->in0;
        uint64_t *counts4 = (uint64_t *)d->out0;
        double *gc = (double *)d->out1;
        size_t wpr = (d->k + 31) / 32;
        uint64_t sum = 0;
        for (size_t r = u0; r < u1; ++r) {
            base_counts_ref(w + r * wpr, d->k, counts4 + 4 * r);
            gc[r] = gc_content_ref(w + r * wpr, d->k);
            sum += counts4[4 * r + 1] + counts4[4 * r + 2];
        }
        *check = sum;
        return 0;
    }
    default:
        return -3;
    }
}

static void *op_worker(void *p) {
    op_job *j = (op_job *)p;
    if (j->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(j->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    size_t align = (j->d->op == ORC_OP_ENCODE || j->d->op == ORC_OP_DECODE || j->d->op == ORC_OP_HDIST) ? 128 : 1;
    size_t u0, u1;
    if (j->d->op == ORC_OP_ENCODE_BATCH) { /* reads cut by byte volume: the first read whose start reaches the ideal cut */
        const uint64_t *off = (const uint64_t *)j->d->in1;
        size_t n = j->d->n_units, cut[2];
        for (int e = 0; e < 2; ++e) {
            uint64_t target = off[0] + (off[n] - off[0]) / (uint64_t)j->n_threads * (uint64_t)(j->tid + e);
            size_t lo = 0, hi = n;
            while (lo < hi) {
                size_t mid = (lo + hi) / 2;
                if (off[mid] < target) lo = mid + 1; else hi = mid;
            }
            cut[e] = (j->tid + e == j->n_threads) ? n : lo;
        }
        u0 = cut[0];
        u1 = cut[1];
    } else {
        op_range(j->d, j->tid, j->n_threads, align, &u0, &u1);
    }
    for (int r = 0; r < j->reps; ++r) {
        pthread_barrier_wait(j->bar);
        uint64_t c = 0;
        int rc = op_run(j->d, u0, u1, &c);
        if (rc) j->rc = rc;
        j->check = c;
        pthread_barrier_wait(j->bar);
    }
    return NULL;
}

int orc_bench_op(const orc_bench_desc *d, int n_threads, int reps, int pin, double *times, uint64_t *check) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    if (reps < 1) return -2;
    if (d->path == ORC_PATH_AVX2 && !orc_have_avx2()) return -2;
    int cpus[1024], n_cpus = 0;
    cpu_set_t mine;
    if (pin && sched_getaffinity(0, sizeof(mine), &mine) == 0)
        for (int c = 0; c < CPU_SETSIZE && n_cpus < 1024; ++c)
            if (CPU_ISSET(c, &mine)) cpus[n_cpus++] = c;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)n_threads + 1);
    op_job *jobs = (op_job *)calloc((size_t)n_threads, sizeof(op_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (op_job){d, &bar, t, n_threads, reps, n_cpus ? cpus[t % n_cpus] : -1, 0, 0};
        pthread_create(&tids[t], NULL, op_worker, &jobs[t]);
    }
    for (int r = 0; r < reps; ++r) {
        pthread_barrier_wait(&bar);
        double t0 = now_s();
        pthread_barrier_wait(&bar);
        times[r] = now_s() - t0;
    }
    int rc = 0;
    uint64_t sum = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(tids[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        sum += jobs[t].check;
    }
    if (check) *check = sum;
    pthread_barrier_destroy(&bar);
    free(jobs);
    free(tids);
    return rc;
}
