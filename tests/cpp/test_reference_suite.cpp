// test_reference_suite.cpp -- the reference's own unit tests, restated against the C++ host mirror
// (include/bitnuc.hpp) of the drop-in API.  One function per #[test] of the reference, same inputs,
// same assertions; every call runs on the GPU through libbitnuc_cuda.so.
//   build: g++ -std=c++17 -O1 -I include tests/cpp/test_reference_suite.cpp -L bitnuc_b200 -lbitnuc_cuda -Wl,-rpath,$PWD/bitnuc_b200
#include <cctype>
#include <cstdio>
#include <functional>
#include <random>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "bitnuc.hpp"

using namespace bitnuc;
using E = NucleotideError;

static int g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) {                                                               \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

template <class F>
static E expect_err(F f) {
    try {
        f();
    } catch (const E& e) {
        return e;
    }
    std::printf("FAILED: expected a NucleotideError\n");
    std::exit(1);
}
static std::vector<uint8_t> bytes(const std::string& s) { return std::vector<uint8_t>(s.begin(), s.end()); }
static std::string repeat(const std::string& s, int n) {
    std::string r;
    for (int i = 0; i < n; ++i) r += s;
    return r;
}

// src/utils/packing/mod.rs:145-198
static void test_as_2bit() {
    CHECK(as_2bit("ACGT") == 0b11100100);
    CHECK(as_2bit("AAAA") == 0b00000000);
    CHECK(as_2bit("TTTT") == 0b11111111);
    CHECK(as_2bit("GGGG") == 0b10101010);
    CHECK(as_2bit("CCCC") == 0b01010101);
    CHECK(as_2bit("ACTGACTGACTGACTG") == 0b10110100101101001011010010110100ull);
    CHECK(as_2bit("ACTGGAAAATTTTAAGG") == 0b1010000011111111000000001010110100ull);
    CHECK(as_2bit("acgt") == as_2bit("ACGT"));
    CHECK(expect_err([] { as_2bit("ACGN"); }) == E(E::InvalidBase, 'N'));
    CHECK(expect_err([] { as_2bit(repeat("A", 33)); }) == E(E::SequenceTooLong, 33));
}

// src/utils/unpacking/mod.rs:184-216, src/utils/unpacking/avx.rs:155-196
static void test_from_2bit() {
    std::vector<uint8_t> unpacked;
    from_2bit(0b11100100, 4, unpacked);
    CHECK(unpacked == bytes("ACGT"));
    unpacked.clear();
    from_2bit(0, 4, unpacked);
    CHECK(unpacked == bytes("AAAA"));
    unpacked.clear();
    from_2bit(0xFF, 4, unpacked);
    CHECK(unpacked == bytes("TTTT"));
    CHECK(from_2bit_alloc(71620941647064936ull, 28) == bytes("AGGCTTGAGGCCCATTCTCTGATCGTTT"));
    CHECK(expect_err([] { std::vector<uint8_t> v; from_2bit(0, 33, v); }) == E(E::InvalidLength, 33));
    const std::string input = "ACTGACTGACTGACTGACTGACTGACTGACTG";
    for (size_t len = 1; len <= 32; ++len) {
        std::vector<uint8_t> obs;
        from_2bit(as_2bit(Bytes(reinterpret_cast<const uint8_t*>(input.data()), len)), len, obs);
        CHECK(obs == bytes(input.substr(0, len)));
    }
    std::vector<uint8_t> observed;  // test_append: from_2bit appends, never clears
    const uint64_t packed = as_2bit("ACTGACTGACTGACTGACTG");
    from_2bit(packed, 10, observed);
    from_2bit(packed, 10, observed);
    CHECK(observed == bytes("ACTGACTGACACTGACTGAC"));
}

// src/utils/mod.rs:64-134
static void test_roundtrips() {
    for (const char* s : {"A", "C", "G", "T", "AC", "GT", "ACG", "TGC", "ACGT", "TGCA", "ACGTACGT", "AAAA", "CCCC", "GGGG", "TTTT"}) {
        std::vector<uint8_t> unpacked;
        from_2bit(as_2bit(s), std::strlen(s), unpacked);
        CHECK(unpacked == bytes(s));
    }
    std::vector<uint8_t> unpacked;
    const uint64_t packed = as_2bit("ACGT");
    from_2bit(packed, 2, unpacked);
    CHECK(unpacked == bytes("AC"));
    unpacked.clear();
    from_2bit(packed, 3, unpacked);
    CHECK(unpacked == bytes("ACG"));
    std::mt19937_64 rng(20261018);  // test_large_sequence_round_trip: every length 1..=1000
    for (size_t len = 1; len <= 1000; ++len) {
        std::vector<uint8_t> seq(len);
        for (auto& b : seq) b = "ACGT"[rng() & 3];
        std::vector<uint64_t> ebuf;
        encode(seq, ebuf);
        std::vector<uint8_t> back;
        decode(ebuf, len, back);
        CHECK(seq == back);
    }
}

// src/utils/functions/hamming/scalar.rs:50-116, multi.rs:162-208, benches/hdist_benchmark.rs
static void test_hdist() {
    CHECK(expect_err([] { hdist_scalar(0, 0, 33); }) == E(E::InvalidLength, 33));
    CHECK(hdist_scalar(0, 0, 0) == 0);
    CHECK(hdist_scalar(0, 0, 32) == 0);
    CHECK(hdist_scalar(0, 0, 1) == 0);
    CHECK(hdist_scalar(0xFFFFFFFF, 0xFFFFFFFF, 16) == 0);
    CHECK(hdist_scalar(~0ull, ~0ull, 32) == 0);
    CHECK(hdist_scalar(0b0001, 0b0010, 2) == 1);
    CHECK(hdist_scalar(0b0001, 0b0011, 2) == 1);
    CHECK(hdist_scalar(0b0010, 0b0011, 2) == 1);
    const std::pair<const char*, const char*> cases[] = {{"AAAA", "AAAA"}, {"AAAA", "AAAT"}, {"AAAA", "AATT"},
                                                         {"AAAA", "ATTT"}, {"AAAA", "TTTT"}};
    for (uint32_t i = 0; i < 5; ++i) CHECK(hdist_scalar(as_2bit(cases[i].first), as_2bit(cases[i].second), 4) == i);
    CHECK(hdist_scalar(as_2bit("ACTGACTG"), as_2bit("TGCATGCA"), 8) == 8);
    std::vector<uint64_t> one(1, 0);
    CHECK(expect_err([&] { hdist(one, one, 64); }) == E(E::InvalidLength, 64));
    const std::string seq = repeat("ACTG", 16);
    const auto buf = encode_alloc(seq);
    CHECK(hdist(buf, buf, seq.size()) == 0);
    CHECK(hdist(encode_alloc(repeat("A", 128)), encode_alloc(repeat("T", 128)), 128) == 128);
    for (size_t len = 1; len <= 256; ++len)
        CHECK(hdist(encode_alloc(repeat("A", (int)len)), encode_alloc(repeat("T", (int)len)), len) == len);
    for (auto [l, expected] : {std::pair<size_t, uint32_t>{512, 383}, {32, 23}}) {
        std::string s1, s2;
        for (size_t i = 0; i < l; ++i) {
            s1 += "ACGT"[i % 4];
            s2 += "ACGT"[i % 3];
        }
        CHECK(hdist(encode_alloc(s1), encode_alloc(s2), l) == expected);
    }
}

// src/utils/analysis.rs:41-84, src/sequence.rs:264-339, src/lib.rs:222-265
static void test_packed_sequence() {
    const struct { const char* seq; double gc; uint64_t counts[4]; } tests[] = {
        {"ACGT", 50.0, {1, 1, 1, 1}}, {"AAAA", 0.0, {4, 0, 0, 0}}, {"CCCC", 100.0, {0, 4, 0, 0}},
        {"AACG", 50.0, {2, 1, 1, 0}}, {"ACGTA", 40.0, {2, 1, 1, 1}}, {"", 0.0, {0, 0, 0, 0}}, {"ACGTACGT", 50.0, {2, 2, 2, 2}}};
    for (const auto& t : tests) {
        PackedSequence packed(t.seq);
        CHECK(packed.gc_content() == t.gc);
        const auto c = packed.base_counts();
        for (int i = 0; i < 4; ++i) CHECK(c.v[i] == t.counts[i]);
    }
    PackedSequence seq("ACGT");
    CHECK(seq.len() == 4 && !seq.is_empty());
    CHECK(seq.to_vec() == bytes("ACGT"));
    CHECK(seq.get(0) == 'A' && seq.get(1) == 'C' && seq.get(2) == 'G' && seq.get(3) == 'T');
    CHECK(expect_err([&] { seq.get(4); }) == E(E::IndexOutOfBounds, 4, 4));
    CHECK(expect_err([&] { seq.slice(3, 2); }) == E(E::InvalidRange, 3, 2, 4));
    CHECK(expect_err([&] { seq.slice(2, 5); }) == E(E::InvalidRange, 2, 5, 4));
    CHECK(expect_err([] { PackedSequence bad("ACGN"); }) == E(E::InvalidBase, 'N'));
    PackedSequence s8("ACGTACGT");
    CHECK(s8.slice(1, 5) == bytes("CGTA") && s8.slice(2, 6) == bytes("GTAC") && s8.slice(0, 3) == bytes("ACG"));
    CHECK(s8.get(7) == 'T' && s8.slice(2, 2).empty());
    CHECK(PackedSequence("").is_empty() && PackedSequence("").to_vec().empty());
    PackedSequence seq2("ACGT"), seq3("TGCA");
    CHECK(seq == seq2 && seq != seq3);
    std::unordered_set<PackedSequence> set;
    set.insert(seq);
    CHECK(set.count(seq2) == 1 && set.count(seq3) == 0);
}

// src/error.rs:20-45 and the encode error contract (packing/avx.rs:132,142-143)
static void test_errors() {
    CHECK(std::string(E(E::InvalidBase, 78).what()) == "Invalid nucleotide base: 78");
    CHECK(std::string(E(E::SequenceTooLong, 33).what()) == "Sequence length 33 exceeds maximum");
    CHECK(std::string(E(E::InvalidLength, 64).what()) == "Invalid length: 64");
    CHECK(std::string(E(E::IndexOutOfBounds, 4, 4).what()) == "Index 4 out of bounds for sequence of length 4");
    CHECK(std::string(E(E::InvalidRange, 3, 2, 4).what()) == "Invalid range 3..2 for sequence of length 4");
    CHECK(std::string(E(E::Unsupported).what()) == "Unsupported architecture");
    std::vector<uint64_t> ebuf = {1, 2, 3};
    const std::string bad = repeat("ACGT", 16) + "ACNT";
    CHECK(expect_err([&] { encode(bad, ebuf); }) == E(E::InvalidBase, 'N'));
    CHECK(ebuf == std::vector<uint64_t>(2, 0xE4E4E4E4E4E4E4E4ull));
    bool panicked = false;
    try {
        encode("", ebuf);
    } catch (const std::logic_error&) {
        panicked = true;
    }
    CHECK(panicked);
}

// src/utils/functions/split.rs:110-225
static void test_split_packed() {
    auto dec = [](const std::vector<uint64_t>& w, size_t n) {
        std::vector<uint8_t> out;
        decode(w, n, out);
        return std::string(out.begin(), out.end());
    };
    std::vector<uint64_t> lbuf, rbuf;
    {
        const std::string seq = "ACTGACTG";
        split_packed(encode_alloc(seq), seq.size(), 4, lbuf, rbuf);
        CHECK(lbuf.size() == 1 && rbuf.size() == 1);
        CHECK(dec(lbuf, 4) == "ACTG" && dec(rbuf, 4) == "ACTG");
    }
    {
        const std::string seq = "ACTG";
        const auto ebuf = encode_alloc(seq);
        split_packed(ebuf, seq.size(), 0, lbuf, rbuf);
        CHECK(dec(rbuf, seq.size()) == seq && lbuf.size() == 0 && rbuf.size() == 1);
        split_packed(ebuf, seq.size(), seq.size(), lbuf, rbuf);
        CHECK(dec(lbuf, seq.size()) == seq && lbuf.size() == 1 && rbuf.size() == 0);
        CHECK(expect_err([&] { split_packed(ebuf, seq.size(), seq.size() + 1, lbuf, rbuf); }) == E(E::IndexOutOfBounds, 5, 4));
    }
    {
        const std::string seq = "ACTGACTGAC";
        split_packed(encode_alloc(seq), seq.size(), 7, lbuf, rbuf);
        CHECK(lbuf.size() == 1 && rbuf.size() == 1);
        CHECK(dec(lbuf, 7) == "ACTGACT" && dec(rbuf, 3) == "GAC");
    }
    {
        const std::string seq = repeat("ACTG", 10);
        split_packed(encode_alloc(seq), seq.size(), 32, lbuf, rbuf);
        CHECK(lbuf.size() == 2 && rbuf.size() == 1);
        CHECK(dec(lbuf, 32) == seq.substr(0, 32) && dec(rbuf, 8) == seq.substr(32));
    }
}

// README.md:160-180 ("Efficient k-mer counting") with the whole window loop in one call
static void test_readme_kmer_counting() {
    const std::string sequence = "ACGTACGT";
    std::unordered_map<uint64_t, int> kmer_counts;
    for (uint64_t packed : kmers(sequence, 4)) ++kmer_counts[packed];
    CHECK(kmer_counts[as_2bit("ACGT")] == 2);
    CHECK(kmers(sequence, 4).size() == 5);
    for (size_t i = 0; i + 4 <= sequence.size(); ++i) CHECK(kmers(sequence, 4)[i] == as_2bit(sequence.substr(i, 4)));
    CHECK(expect_err([] { kmers("ACGTNACGT", 3); }) == E(E::InvalidBase, 'N'));
    CHECK(kmers("ACG", 4).empty());
}

// README.md:160-180: the caller's loop over a FASTQ reader, here one call on the raw text
static void test_fastq_records() {
    const std::string text = "@r1\nACGTACGT\n+\nIIIIIIII\n@r2 second\r\nacgtn\r\n+r2\r\n!!!!!\r\n";
    const E bad = expect_err([&] { fastq_encode(text); });   // b carries the position inside the read
    CHECK(bad.variant == E::InvalidBase && bad.a == 'n' && bad.b == 4);
    const std::string ok = "@r1\nACGTACGT\n+\nIIIIIIII\n@r2 second\r\nacgtt\r\n+r2\r\n!!!!!\r\n@empty\n\n+\n\n";
    const FastqBatch b = fastq_encode(ok);
    CHECK(b.size() == 3);
    CHECK(b.seq_lens[0] == 8 && b.seq_lens[1] == 5 && b.seq_lens[2] == 0);
    CHECK(b.word_offsets == (std::vector<uint64_t>{0, 1, 2, 2}));
    CHECK(b.words[0] == as_2bit("ACGTACGT") && b.words[1] == as_2bit("ACGTT"));
    CHECK(ok.substr(b.seq_offsets[1], b.seq_lens[1]) == "acgtt");
    bool threw = false;
    try {
        fastq_encode(std::string("@r1\nACGT\n-\nIIII\n"));
    } catch (const FastqError& f) {
        threw = f.record == 0 && f.fault == BN_FASTQ_BAD_SEPARATOR;
    }
    CHECK(threw);
    CHECK(fastq_encode(std::string()).size() == 0);
    const std::string fa = ">a\nACGTACGT\n>b second\nacgtt\n";
    const FastqBatch f = fasta_encode(fa);
    CHECK(f.size() == 2 && f.seq_lens[1] == 5 && f.words[1] == as_2bit("ACGTT") && fa.substr(f.seq_offsets[0], 8) == "ACGTACGT");
    const std::string wrapped = ">chr1 x\nACGTAC\nGT\n>chr2\n>chr3\r\nacg\r\ntt\r\n";   // lines of one record are joined
    const FastqBatch w = fasta_wrapped_encode(wrapped);
    CHECK(w.size() == 3 && w.seq_lens[0] == 8 && w.seq_lens[1] == 0 && w.seq_lens[2] == 5);
    CHECK(w.word_offsets == (std::vector<uint64_t>{0, 1, 1, 2}) && w.words[0] == as_2bit("ACGTACGT") && w.words[1] == as_2bit("ACGTT"));
    CHECK(wrapped.substr(w.seq_offsets[2], 5) == ">chr3");
}

// bitnuc::Multi: the reference's functions with the work cut over two shards (device 0 named twice -- the way a one-GPU
// box exercises the sharding; that needs the NVLink-mailbox reduce, NCCL refuses a repeated device).  Results must equal
// the single-context calls.
static void test_multi() {
    Multi m({0, 0}, Multi::Reduce::P2p);
    CHECK(m.size() == 2);
    std::string seq;
    for (int i = 0; i < 5000; ++i) seq += "ACGTTGCAacgt"[(i * 7 + i / 13) % 12];
    std::vector<uint64_t> a, b;
    encode(seq, a);
    m.encode(seq, b);
    CHECK(a == b);
    std::vector<uint8_t> back{'x'};
    m.decode(b, seq.size(), back);                       // appends
    CHECK(back.size() == seq.size() + 1 && back[0] == 'x');
    for (size_t i = 0; i < seq.size(); ++i) CHECK(back[i + 1] == (uint8_t)std::toupper(seq[i]));
    std::string other = seq;
    for (size_t i = 0; i < other.size(); i += 17) other[i] = other[i] == 'A' ? 'C' : 'A';
    std::vector<uint64_t> c;
    encode(other, c);
    CHECK(m.hdist(a, c, seq.size()) == hdist(a, c, seq.size()));
    PackedSequence ps(seq);
    double gc = -1.0;
    CHECK(m.base_counts(a, seq.size(), &gc) == ps.base_counts() && gc == ps.gc_content());
    bool threw = false;
    try {
        std::string bad = seq;
        bad[4321] = 'N';
        m.encode(bad, b);
    } catch (const NucleotideError& e) {
        threw = e.variant == NucleotideError::InvalidBase && e.a == 'N';
    }
    CHECK(threw && b.size() == 4321 / 32);               // the words of the chunks before the failing chunk (avx.rs:142-143)
    const std::vector<uint64_t> offsets{0, 10, 10, 75, 5000};
    const Multi::PackedBatch pb = m.encode_batch(seq, offsets);
    CHECK(pb.word_offsets == (std::vector<uint64_t>{0, 1, 1, 4, 4 + (4925 + 31) / 32}));
    CHECK(pb.words[0] == as_2bit(seq.substr(0, 10)) && pb.words[1] == as_2bit(seq.substr(10, 32)));
}

int main() {
    const std::pair<const char*, std::function<void()>> tests[] = {
        {"as_2bit", test_as_2bit}, {"from_2bit", test_from_2bit}, {"roundtrips", test_roundtrips},
        {"hdist", test_hdist},     {"packed_sequence", test_packed_sequence}, {"errors", test_errors},
        {"split_packed", test_split_packed}, {"readme_kmer_counting", test_readme_kmer_counting},
        {"fastq_records", test_fastq_records}, {"multi", test_multi}};
    for (const auto& t : tests) {
        t.second();
        std::printf("ok %s\n", t.first);
        std::fflush(stdout);
    }
    std::printf("reference suite passed: %d checks\n", g_checks);
    return 0;
}
