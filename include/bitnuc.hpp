// bitnuc.hpp -- host-side C++ mirror of the reference's public API over the C ABI (bitnuc_cuda.h).
//
// The reference is a compiled (Rust) library whose toolchain is absent from this image, so the host
// side above the C ABI is written in C++ with the reference's names, argument meaning and error
// behaviour (/root/reference/src/lib.rs:214-220):
//
//   Rust                                               here
//   as_2bit(&[u8]) -> Result<u64>                      uint64_t as_2bit(Bytes)                       throws NucleotideError
//   encode(&[u8], &mut Vec<u64>) -> Result<()>         void encode(Bytes, std::vector<uint64_t>&)    clears, then fills
//   encode_alloc(&[u8]) -> Result<Vec<u64>>            std::vector<uint64_t> encode_alloc(Bytes)
//   decode(&[u64], usize, &mut Vec<u8>) -> Result<()>  void decode(Words, size_t, std::vector<uint8_t>&)   appends
//   from_2bit(u64, usize, &mut Vec<u8>) -> Result<()>  void from_2bit(uint64_t, size_t, std::vector<uint8_t>&)  appends
//   from_2bit_alloc(u64, usize) -> Result<Vec<u8>>     std::vector<uint8_t> from_2bit_alloc(uint64_t, size_t)
//   hdist(&[u64], &[u64], usize) -> Result<u32>        uint32_t hdist(Words, Words, size_t)
//   hdist_scalar(u64, u64, usize) -> Result<u32>       uint32_t hdist_scalar(uint64_t, uint64_t, size_t)
//   PackedSequence::{new,len,is_empty,get,slice,to_vec} + BaseCount + GCContent     class PackedSequence
//   NucleotideError (6 variants, Display)              class NucleotideError : std::exception
//
// `Result<_, NucleotideError>` becomes an exception; where the reference panics (encode of an empty
// slice, src/utils/packing/avx.rs:138) std::logic_error is thrown.  Everything computes on the GPU
// through libbitnuc_cuda.so; CUDA failures surface as std::runtime_error (there is no CPU fallback).
// Header-only; link with -lbitnuc_cuda.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "bitnuc_cuda.h"

namespace bitnuc {

// borrowed slices (&[u8], &[u64])
struct Bytes {
    const uint8_t* ptr;
    size_t len;
    Bytes(const uint8_t* p, size_t n) : ptr(p), len(n) {}
    Bytes(const char* s) : ptr(reinterpret_cast<const uint8_t*>(s)), len(std::strlen(s)) {}
    Bytes(const std::string& s) : ptr(reinterpret_cast<const uint8_t*>(s.data())), len(s.size()) {}
    Bytes(const std::vector<uint8_t>& v) : ptr(v.data()), len(v.size()) {}
};
struct Words {
    const uint64_t* ptr;
    size_t len;
    Words(const uint64_t* p, size_t n) : ptr(p), len(n) {}
    Words(const std::vector<uint64_t>& v) : ptr(v.data()), len(v.size()) {}
};

// src/error.rs:4-47
class NucleotideError : public std::exception {
public:
    enum Variant { InvalidBase = 1, SequenceTooLong, InvalidLength, IndexOutOfBounds, InvalidRange, Unsupported };
    Variant variant;
    uint64_t a, b, c;  // payload fields in declaration order
    explicit NucleotideError(const bn_error_t& e) : variant(static_cast<Variant>(e.code)), a(e.a), b(e.b), c(e.c) {
        char buf[160];
        bn_error_string(&e, buf, sizeof buf);
        text_ = buf;
    }
    NucleotideError(Variant v, uint64_t a_ = 0, uint64_t b_ = 0, uint64_t c_ = 0) : variant(v), a(a_), b(b_), c(c_) {
        bn_error_t e{};
        e.code = v;
        e.a = a_;
        e.b = b_;
        e.c = c_;
        e.base = static_cast<uint8_t>(a_);
        char buf[160];
        bn_error_string(&e, buf, sizeof buf);
        text_ = buf;
    }
    const char* what() const noexcept override { return text_.c_str(); }  // the reference's Display string
    bool operator==(const NucleotideError& o) const { return variant == o.variant && a == o.a && b == o.b && c == o.c; }

private:
    std::string text_;
};

namespace detail {

inline void check(int rc, const bn_error_t& e) {
    if (rc == BN_OK) return;
    if (rc >= 1 && rc <= 6) throw NucleotideError(e);
    if (rc == BN_ERR_EMPTY_ENCODE) throw std::logic_error("bitnuc: encode of an empty sequence (the reference panics)");
    char buf[200];
    bn_error_t copy = e;
    if (copy.code != rc) copy.code = rc;
    bn_error_string(&copy, buf, sizeof buf);
    throw std::runtime_error(std::string("bitnuc-cuda: ") + buf);
}

// one lazily created context per host thread (the reference's functions are re-entrant)
inline bn_ctx* ctx() {
    struct Holder {
        bn_ctx* c = nullptr;
        ~Holder() { bn_ctx_destroy(c); }
    };
    thread_local Holder h;
    if (!h.c) {
        bn_error_t e{};
        const char* dev = std::getenv("BITNUC_DEVICE");
        int rc = bn_ctx_create(dev ? std::atoi(dev) : 0, &h.c);
        e.code = rc;
        check(rc, e);
    }
    return h.c;
}

}  // namespace detail

inline uint64_t as_2bit(Bytes seq) {  // src/utils/packing/mod.rs:81
    uint64_t out = 0;
    bn_error_t e{};
    detail::check(bn_as_2bit_batch(detail::ctx(), seq.ptr, 1, static_cast<uint32_t>(seq.len > 0xFFFFFFFFu ? 0xFFFFFFFFu : seq.len),
                                   seq.len ? seq.len : 1, &out, &e), e);
    return out;
}

inline void encode(Bytes sequence, std::vector<uint64_t>& ebuf) {  // src/utils/mod.rs:22
    if (sequence.len == 0) detail::check(BN_ERR_EMPTY_ENCODE, bn_error_t{});
    ebuf.clear();
    ebuf.resize((sequence.len + 31) / 32);
    size_t n_words = 0;
    bn_error_t e{};
    const int rc = bn_encode(detail::ctx(), sequence.ptr, sequence.len, ebuf.data(), &n_words, &e);
    ebuf.resize(n_words);  // on InvalidBase: the words of the chunks before the failing chunk (avx.rs:142-143)
    detail::check(rc, e);
}

inline std::vector<uint64_t> encode_alloc(Bytes sequence) {  // src/utils/mod.rs:38
    std::vector<uint64_t> ebuf;
    encode(sequence, ebuf);
    return ebuf;
}

inline void decode(Words ebuf, size_t n_bases, std::vector<uint8_t>& dbuf) {  // src/utils/mod.rs:60 -- appends
    const size_t old = dbuf.size();
    dbuf.resize(old + n_bases);
    bn_error_t e{};
    const int rc = bn_decode(detail::ctx(), ebuf.ptr, ebuf.len, n_bases, dbuf.data() + old, &e);
    if (rc != BN_OK) dbuf.resize(old);
    detail::check(rc, e);
}

inline void from_2bit(uint64_t packed, size_t expected_size, std::vector<uint8_t>& sequence) {  // unpacking/mod.rs:119 -- appends
    const size_t old = sequence.size();
    const uint32_t k = static_cast<uint32_t>(expected_size > 0xFFFFFFFFu ? 0xFFFFFFFFu : expected_size);
    sequence.resize(old + (k <= 32 ? k : 0));
    bn_error_t e{};
    const int rc = bn_from_2bit_batch(detail::ctx(), &packed, 1, k, sequence.data() + old, k ? k : 1, &e);
    if (rc != BN_OK) sequence.resize(old);
    detail::check(rc, e);
}

inline std::vector<uint8_t> from_2bit_alloc(uint64_t packed, size_t expected_size) {  // unpacking/mod.rs:178
    std::vector<uint8_t> sequence;
    sequence.reserve(expected_size <= 32 ? expected_size : 0);
    from_2bit(packed, expected_size, sequence);
    return sequence;
}

inline uint64_t hdist_total(Words ebuf1, Words ebuf2, size_t n_bases) {
    uint64_t total = 0;
    bn_error_t e{};
    detail::check(bn_hdist(detail::ctx(), ebuf1.ptr, ebuf1.len, ebuf2.ptr, ebuf2.len, n_bases, &total, &e), e);
    return total;
}

// src/utils/functions/hamming/multi.rs:122 -- the reference accumulates in a u32 that wraps in release builds
inline uint32_t hdist(Words ebuf1, Words ebuf2, size_t n_bases) { return static_cast<uint32_t>(hdist_total(ebuf1, ebuf2, n_bases)); }

inline uint32_t hdist_scalar(uint64_t u, uint64_t v, size_t len) {  // hamming/scalar.rs:11
    uint32_t out = 0;
    bn_error_t e{};
    detail::check(bn_hdist_pairs(detail::ctx(), &u, &v, 1, static_cast<uint32_t>(len > 0xFFFFFFFFu ? 0xFFFFFFFFu : len), &out, &e), e);
    return out;
}

// `for w in seq.windows(k) { as_2bit(w)? }` (README.md:160-180) in one call: one packed word per window
inline std::vector<uint64_t> kmers(Bytes seq, size_t k) {
    std::vector<uint64_t> out(seq.len >= k && k ? seq.len - k + 1 : 0);
    size_t n_out = 0;
    bn_error_t e{};
    detail::check(bn_kmers(detail::ctx(), seq.ptr, seq.len, static_cast<uint32_t>(k > 0xFFFFFFFFu ? 0xFFFFFFFFu : k), out.data(), &n_out, &e), e);
    out.resize(n_out);
    return out;
}

// FASTQ text -> one PackedSequence-shaped record per read, parsed and encoded on the device: the caller's loop
// `for record in reader { PackedSequence::new(record.seq())? }` (README.md:160-180, src/sequence.rs:40-52).
struct FastqError : std::runtime_error {   // BN_ERR_FASTQ: the reference has no parser, hence no variant for this
    uint64_t record;
    int fault;                             // bn_fastq_fault
    FastqError(uint64_t r, int f, const std::string& text) : std::runtime_error(text), record(r), fault(f) {}
};
struct FastqBatch {
    std::vector<uint64_t> words, word_offsets;   // read r = words[word_offsets[r] .. word_offsets[r+1])
    std::vector<uint64_t> seq_offsets, seq_lens; // where each sequence line sits in the text
    size_t size() const { return seq_lens.size(); }
};
inline FastqBatch fastq_encode(Bytes text, bool fasta = false) {   // fasta: '>' header line + ONE sequence line per record
    FastqBatch out;
    size_t n_reads = 0, n_words = 0;
    bn_error_t e{};
    int rc = (fasta ? bn_fasta_scan : bn_fastq_scan)(detail::ctx(), text.ptr, text.len, &n_reads, &n_words, &e);
    if (rc == BN_ERR_FASTQ) {
        char buf[160];
        bn_error_string(&e, buf, sizeof buf);
        throw FastqError(e.record, static_cast<int>(e.a), buf);
    }
    detail::check(rc, e);
    out.words.resize(n_words);
    out.word_offsets.resize(n_reads + 1);
    out.seq_offsets.resize(n_reads);
    out.seq_lens.resize(n_reads);
    detail::check((fasta ? bn_fasta_encode : bn_fastq_encode)(detail::ctx(), text.ptr, text.len, n_reads, n_words, out.words.data(),
                                                             out.word_offsets.data(), out.seq_offsets.data(), out.seq_lens.data(), &e), e);
    return out;
}
inline FastqBatch fasta_encode(Bytes text) { return fastq_encode(text, true); }
// wrapped (multi-line) FASTA: a '>' header line, then any number of sequence lines per record (genome files); seq_offsets holds
// the byte offset of every HEADER line (a record's bases are not contiguous in the text)
inline FastqBatch fasta_wrapped_encode(Bytes text) {
    FastqBatch out;
    size_t n_records = 0, n_bases = 0, n_words = 0;
    bn_error_t e{};
    const int rc = bn_fasta_wrapped_scan(detail::ctx(), text.ptr, text.len, &n_records, &n_bases, &n_words, &e);
    if (rc == BN_ERR_FASTQ) {
        char buf[160];
        bn_error_string(&e, buf, sizeof buf);
        throw FastqError(e.record, static_cast<int>(e.a), buf);
    }
    detail::check(rc, e);
    out.words.resize(n_words);
    out.word_offsets.resize(n_records + 1);
    out.seq_offsets.resize(n_records);
    out.seq_lens.resize(n_records);
    detail::check(bn_fasta_wrapped_encode(detail::ctx(), text.ptr, text.len, n_records, n_words, out.words.data(), out.word_offsets.data(),
                                          out.seq_offsets.data(), out.seq_lens.data(), &e), e);
    return out;
}

// src/utils/functions/split.rs:14-20 -- validates idx <= slen, then clears both buffers and fills them
inline void split_packed(Words ebuf, size_t slen, size_t idx, std::vector<uint64_t>& lbuf, std::vector<uint64_t>& rbuf) {
    const uint64_t word_offsets[2] = {0, ebuf.len}, len64 = slen, idx64 = idx;
    uint64_t lo[2] = {0, 0}, ro[2] = {0, 0};
    std::vector<uint64_t> left(ebuf.len + 1), right(ebuf.len + 1);
    bn_error_t e{};
    detail::check(bn_split_packed_batch(detail::ctx(), ebuf.ptr, ebuf.len, word_offsets, &len64, &idx64, 1, left.data(), lo,
                                        right.data(), ro, &e), e);
    left.resize(lo[1]);
    right.resize(ro[1]);
    lbuf.swap(left);
    rbuf.swap(right);
}

// src/sequence.rs:5-9 + src/utils/analysis.rs
class PackedSequence {
public:
    explicit PackedSequence(Bytes seq) : length_(seq.len) {  // PackedSequence::new, sequence.rs:40-52
        if (seq.len != 0) encode(seq, data_);
    }
    size_t len() const { return length_; }
    bool is_empty() const { return length_ == 0; }
    uint8_t get(size_t index) const {  // sequence.rs:116-135 (host-side bit poke)
        if (index >= length_) throw NucleotideError(NucleotideError::IndexOutOfBounds, index, length_);
        return "ACGT"[(data_[index / 32] >> ((index % 32) * 2)) & 3];
    }
    std::vector<uint8_t> slice(size_t start, size_t end) const {  // sequence.rs:198-212
        if (start > end || end > length_) throw NucleotideError(NucleotideError::InvalidRange, start, end, length_);
        std::vector<uint8_t> out;
        out.reserve(end - start);
        for (size_t i = start; i < end; ++i) out.push_back(get(i));
        return out;
    }
    std::vector<uint8_t> to_vec() const {  // sequence.rs:260-262, decoded on the GPU
        std::vector<uint8_t> out;
        if (length_) decode(data_, length_, out);
        return out;
    }
    struct Counts {
        uint64_t v[4];  // [A, C, G, T]
        bool operator==(const Counts& o) const { return std::memcmp(v, o.v, sizeof v) == 0; }
    };
    Counts base_counts() const {  // BaseCount::base_counts, analysis.rs:19-39
        Counts c{};
        bn_error_t e{};
        detail::check(bn_base_counts(detail::ctx(), data_.data(), data_.size(), length_, c.v, nullptr, &e), e);
        return c;
    }
    double gc_content() const {  // GCContent::gc_content, analysis.rs:3-17
        uint64_t counts[4];
        double gc = 0.0;
        bn_error_t e{};
        detail::check(bn_base_counts(detail::ctx(), data_.data(), data_.size(), length_, counts, &gc, &e), e);
        return gc;
    }
    const std::vector<uint64_t>& data() const { return data_; }
    bool operator==(const PackedSequence& o) const { return length_ == o.length_ && data_ == o.data_; }
    bool operator!=(const PackedSequence& o) const { return !(*this == o); }

private:
    std::vector<uint64_t> data_;
    size_t length_;
};

// The hot path sharded over the GPUs of one box from one process (bn_multi_* of bitnuc_cuda.h).  The reference is
// single-threaded and has no analogue (src/lib.rs:214-220 is its whole surface); the methods keep the reference's
// argument order and Vec semantics so that `multi.encode(seq, ebuf)` reads like `bitnuc::encode(seq, ebuf)`.  Shards are
// contiguous (base ranges on 64-base boundaries, pairs / reads by index, variable-length reads by byte volume); the four
// base counters are summed over the shards by ncclAllReduce (Reduce::Nccl) or by the library's own all-reduce kernel over
// NVLink peer memory (Reduce::P2p; also the mode for a device list that names one device twice).
class Multi {
public:
    enum class Reduce { Nccl = BN_REDUCE_NCCL, P2p = BN_REDUCE_P2P };
    struct PackedBatch {
        std::vector<uint64_t> words, word_offsets;   // read r = words[word_offsets[r] .. word_offsets[r+1])
    };
    // devices empty = every visible device
    explicit Multi(const std::vector<int>& devices = {}, Reduce reduce = Reduce::Nccl) {
        bn_error_t e{};
        const int rc = bn_multi_create(devices.empty() ? nullptr : devices.data(), static_cast<int>(devices.size()), static_cast<int>(reduce), &m_);
        e.code = rc;
        detail::check(rc, e);
    }
    ~Multi() { bn_multi_destroy(m_); }
    Multi(const Multi&) = delete;
    Multi& operator=(const Multi&) = delete;
    int size() const { return bn_multi_size(m_); }
    bn_multi* handle() const { return m_; }

    void encode(Bytes sequence, std::vector<uint64_t>& ebuf) const {   // src/utils/mod.rs:22, over all devices
        if (sequence.len == 0) detail::check(BN_ERR_EMPTY_ENCODE, bn_error_t{});
        ebuf.clear();
        ebuf.resize((sequence.len + 31) / 32);
        size_t n_words = 0;
        bn_error_t e{};
        const int rc = bn_multi_encode(m_, sequence.ptr, sequence.len, ebuf.data(), &n_words, &e);
        ebuf.resize(n_words);
        detail::check(rc, e);
    }
    void decode(Words ebuf, size_t n_bases, std::vector<uint8_t>& dbuf) const {   // src/utils/mod.rs:60 -- appends
        const size_t old = dbuf.size();
        dbuf.resize(old + n_bases);
        bn_error_t e{};
        const int rc = bn_multi_decode(m_, ebuf.ptr, ebuf.len, n_bases, dbuf.data() + old, &e);
        if (rc != BN_OK) dbuf.resize(old);
        detail::check(rc, e);
    }
    uint64_t hdist_total(Words ebuf1, Words ebuf2, size_t n_bases) const {
        uint64_t total = 0;
        bn_error_t e{};
        detail::check(bn_multi_hdist(m_, ebuf1.ptr, ebuf1.len, ebuf2.ptr, ebuf2.len, n_bases, &total, &e), e);
        return total;
    }
    uint32_t hdist(Words ebuf1, Words ebuf2, size_t n_bases) const { return static_cast<uint32_t>(hdist_total(ebuf1, ebuf2, n_bases)); }
    std::vector<uint32_t> hdist_pairs(Words u, Words v, size_t len) const {   // hdist_scalar per pair, hamming/scalar.rs:11
        const size_t n = u.len < v.len ? u.len : v.len;
        std::vector<uint32_t> out(n);
        bn_error_t e{};
        detail::check(bn_multi_hdist_pairs(m_, u.ptr, v.ptr, n, static_cast<uint32_t>(len > 0xFFFFFFFFu ? 0xFFFFFFFFu : len), out.data(), &e), e);
        return out;
    }
    // BaseCount::base_counts + GCContent::gc_content of one packed sequence; the counters are all-reduced over the shards
    PackedSequence::Counts base_counts(Words words, size_t n_bases, double* gc = nullptr) const {
        PackedSequence::Counts c{};
        bn_error_t e{};
        detail::check(bn_multi_base_counts(m_, words.ptr, words.len, n_bases, c.v, gc, &e), e);
        return c;
    }
    // the caller's loop `for read in reads { PackedSequence::new(read)? }` over a batch of offset-indexed reads
    PackedBatch encode_batch(Bytes bytes, const std::vector<uint64_t>& offsets) const {
        PackedBatch out;
        const size_t n_reads = offsets.empty() ? 0 : offsets.size() - 1;
        out.word_offsets.assign(n_reads + 1, 0);
        out.words.resize(n_reads ? static_cast<size_t>((offsets[n_reads] - offsets[0]) / 32) + n_reads : 0);
        bn_error_t e{};
        detail::check(bn_multi_encode_batch(m_, bytes.ptr, offsets.data(), n_reads, out.words.data(), out.word_offsets.data(), nullptr, &e), e);
        out.words.resize(out.word_offsets[n_reads]);
        return out;
    }

private:
    bn_multi* m_ = nullptr;
};

}  // namespace bitnuc

namespace std {
template <>
struct hash<bitnuc::PackedSequence> {  // #[derive(Hash)] over (data, length)
    size_t operator()(const bitnuc::PackedSequence& s) const noexcept {
        size_t h = std::hash<size_t>()(s.len());
        for (uint64_t w : s.data()) h = h * 0x9E3779B97F4A7C15ull + std::hash<uint64_t>()(w);
        return h;
    }
};
}  // namespace std
