"""The oracle (C restatement, its AVX2 baseline variants and the numpy restatement) against every
known-answer vector in the reference's own tests.  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import OracleError, OraclePanic, PATH_AVX2, PATH_NAIVE
from oracle import oracle_np as onp

AVX2 = [False] + ([True] if oracle.have_avx2() else [])


def expect(exc_info, key):
    assert exc_info.value.key() == tuple(key)


@pytest.mark.parametrize("avx2", AVX2)
def test_as_2bit_kats(kats, to_int, avx2):
    for k in kats["as_2bit"]:
        assert oracle.as_2bit(k["seq"].encode(), avx2=avx2) == to_int(k["packed"]), k["src"]
        assert onp.as_2bit(k["seq"].encode()) == to_int(k["packed"]), k["src"]
    for k in kats["as_2bit_equal"]:
        assert oracle.as_2bit(k["a"].encode(), avx2=avx2) == oracle.as_2bit(k["b"].encode(), avx2=avx2)


@pytest.mark.parametrize("avx2", AVX2)
def test_as_2bit_errors(kats, avx2):
    for k in kats["as_2bit_errors"]:
        seq = k["seq"].encode() if "seq" in k else k["seq_repeat"][0].encode() * k["seq_repeat"][1]
        with pytest.raises(OracleError) as ei:
            oracle.as_2bit(seq, avx2=avx2)
        expect(ei, k["error"])
        with pytest.raises(onp.NpError) as ei:
            onp.as_2bit(seq)
        expect(ei, k["error"])
    # SequenceTooLong outranks InvalidBase (naive.rs:5-7 / avx.rs:77-79)
    with pytest.raises(OracleError) as ei:
        oracle.as_2bit(b"N" * 40, avx2=avx2)
    expect(ei, ["SequenceTooLong", 40])
    assert oracle.as_2bit(b"", avx2=avx2) == 0


@pytest.mark.parametrize("avx2", AVX2)
def test_from_2bit_kats(kats, to_int, avx2):
    for k in kats["from_2bit"]:
        assert bytes(oracle.from_2bit_alloc(to_int(k["packed"]), k["n"], avx2=avx2)) == k["seq"].encode(), k["src"]
        assert onp.from_2bit(to_int(k["packed"]), k["n"]) == k["seq"].encode()
    for k in kats["from_2bit_errors"]:
        with pytest.raises(OracleError) as ei:
            oracle.from_2bit_alloc(to_int(k["packed"]), k["n"], avx2=avx2)
        expect(ei, k["error"])
    k = kats["from_2bit_append"]
    packed, buf = oracle.as_2bit(k["seq"].encode()), bytearray()
    for _ in range(k["calls"]):
        oracle.from_2bit(packed, k["n"], buf, avx2=avx2)
    assert bytes(buf) == k["expected"].encode()
    k = kats["from_2bit_prefixes"]
    for n in range(k["lengths"][0], k["lengths"][1] + 1):
        s = k["seq"].encode()[:n]
        assert bytes(oracle.from_2bit_alloc(oracle.as_2bit(s, avx2=avx2), n, avx2=avx2)) == s


def test_roundtrips(kats):
    for s in kats["roundtrip_short"]["seqs"]:
        b = s.encode()
        assert bytes(oracle.from_2bit_alloc(oracle.as_2bit(b), len(b))) == b
    rng = np.random.default_rng(20261018)
    lo, hi = kats["roundtrip_lengths"]["lengths"]
    for n in range(lo, hi + 1):
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
        for avx2 in AVX2:
            words = oracle.encode_np(seq, avx2=avx2)
            assert np.array_equal(words, onp.encode(seq))
            for path in (PATH_AVX2, PATH_NAIVE):
                assert np.array_equal(oracle.decode_np(words, n, path), seq)
        assert np.array_equal(onp.decode(words, n), seq)


def test_encode_semantics():
    # clear-then-fill; on error keeps the words before the failing chunk (avx.rs:132,142-143)
    ebuf = [123, 456]
    oracle.encode(b"ACGT", ebuf)
    assert ebuf == [0xE4]
    seq = b"ACGT" * 8 + b"ACGT" * 8 + b"ACNT"
    with pytest.raises(OracleError) as ei:
        oracle.encode(seq, ebuf)
    expect(ei, ["InvalidBase", ord("N")])
    assert ebuf == [0xE4E4E4E4E4E4E4E4] * 2
    # first invalid byte in sequence order wins
    with pytest.raises(OracleError) as ei:
        oracle.encode(b"ACGT" * 9 + b"X" + b"ACGTN", ebuf)
    expect(ei, ["InvalidBase", ord("X")])
    assert ebuf == [0xE4E4E4E4E4E4E4E4]
    with pytest.raises(OraclePanic):
        oracle.encode(b"", ebuf)  # avx.rs:138 underflow
    assert oracle.encode_alloc(b"acgtACGT") == [0xE4E4]  # lower case accepted


def test_decode_edge_semantics():
    w = oracle.encode_alloc(b"ACGT" * 16)  # 2 words
    dbuf = bytearray(b"xx")
    oracle.decode(w, 40, dbuf)
    assert bytes(dbuf) == b"xx" + b"ACGT" * 10  # append, never clear
    # short ebuf: the naive path reports InvalidLength (unpacking/mod.rs:42-45) ...
    with pytest.raises(OracleError) as ei:
        oracle.decode_np(w[:1], 64, PATH_NAIVE)
    expect(ei, ["InvalidLength", 64])
    # ... the AVX2 path silently yields fewer bases (avx.rs:137) or panics (avx.rs:146)
    assert oracle.decode_np(w[:1], 64, PATH_AVX2).size == 32
    with pytest.raises(OraclePanic):
        oracle.decode_np(w[:1], 40, PATH_AVX2)
    assert oracle.decode_np(w, 0, PATH_AVX2).size == 0
    with pytest.raises(OraclePanic):
        oracle.decode_np(w, 0, PATH_NAIVE)  # n_chunks - 1 underflow


def test_hdist_scalar_kats(kats, to_int):
    for k in kats["hdist_scalar"]:
        assert oracle.hdist_scalar(to_int(k["u"]), to_int(k["v"]), k["len"]) == k["dist"], k["src"]
    for k in kats["hdist_scalar_seqs"]:
        u, v = oracle.as_2bit(k["a"].encode()), oracle.as_2bit(k["b"].encode())
        assert oracle.hdist_scalar(u, v, len(k["a"])) == k["dist"], k["src"]
        assert int(onp.hdist_pairs([u], [v], len(k["a"]))[0]) == k["dist"]
    for k in kats["hdist_scalar_errors"]:
        with pytest.raises(OracleError) as ei:
            oracle.hdist_scalar(to_int(k["u"]), to_int(k["v"]), k["len"])
        expect(ei, k["error"])
    # bits above 2*len are ignored
    assert oracle.hdist_scalar(0xFF00, 0x0000, 4) == 0


@pytest.mark.parametrize("path", [PATH_NAIVE, PATH_AVX2])
def test_hdist_kats(kats, path):
    h = kats["hdist"]
    with pytest.raises(OracleError) as ei:
        oracle.hdist([0] * h["too_small"]["n_words"], [0] * h["too_small"]["n_words"], h["too_small"]["n_bases"], path)
    expect(ei, h["too_small"]["error"])
    s = h["identical"]["seq_repeat"][0].encode() * h["identical"]["seq_repeat"][1]
    buf = oracle.encode_alloc(s)
    assert oracle.hdist(buf, buf, len(s), path) == 0
    lo, hi = h["a_vs_t_lengths"]["lengths"]
    for n in list(range(lo, hi + 1)) + [h["a_vs_t_128"]["n"]]:
        a, t = oracle.encode_alloc(b"A" * n), oracle.encode_alloc(b"T" * n)
        assert oracle.hdist(a, t, n, path) == n
        assert onp.hdist(a, t, n) == n
    for k in h["mod4_vs_mod3"]:
        s1 = bytes(b"ACGT"[i % 4] for i in range(k["n"]))
        s2 = bytes(b"ACGT"[i % 3] for i in range(k["n"]))
        assert oracle.hdist(oracle.encode_alloc(s1), oracle.encode_alloc(s2), k["n"], path) == k["dist"]
    # extra words are ignored, the tail is masked
    a, t = oracle.encode_alloc(b"A" * 70), oracle.encode_alloc(b"T" * 70)
    assert oracle.hdist(a + [7], t + [9, 9], 33, path) == 33


def test_analysis_kats(kats):
    for k in kats["analysis"]:
        ps = oracle.PackedSequence(k["seq"].encode())
        assert ps.base_counts() == k["counts"], k["src"]
        assert ps.gc_content() == k["gc"], k["src"]
        assert onp.base_counts(ps.data, len(ps)) == k["counts"]
        assert onp.gc_content(ps.data, len(ps)) == k["gc"]
    # operation order matters for the last ulp: (1/3)*100 != 100*1/3
    ps = oracle.PackedSequence(b"CAA")
    assert ps.gc_content() == (1.0 / 3.0) * 100.0 != 100.0 * 1.0 / 3.0
    # zero padding in the tail word must not be counted as 'A'
    assert oracle.PackedSequence(b"T" * 33).base_counts() == [0, 0, 0, 33]


def test_packed_sequence_kats(kats):
    p = kats["packed_sequence"]
    with pytest.raises(OracleError) as ei:
        oracle.PackedSequence(p["new_error"]["seq"].encode())
    expect(ei, p["new_error"]["error"])
    ps = oracle.PackedSequence(p["get"]["seq"].encode())
    assert bytes(ps.get(i) for i in range(len(ps))) == p["get"]["values"].encode()
    with pytest.raises(OracleError) as ei:
        oracle.PackedSequence(p["get_oob"]["seq"].encode()).get(p["get_oob"]["index"])
    expect(ei, p["get_oob"]["error"])
    for k in p["slice"]:
        assert oracle.PackedSequence(k["seq"].encode()).slice(k["start"], k["end"]) == k["out"].encode()
    for k in p["slice_errors"]:
        with pytest.raises(OracleError) as ei:
            oracle.PackedSequence(k["seq"].encode()).slice(k["start"], k["end"])
        expect(ei, k["error"])
    for s in p["to_vec"]:
        ps = oracle.PackedSequence(s.encode())
        assert ps.to_vec() == s.encode() and len(ps) == len(s) and ps.is_empty() == (len(s) == 0)
    a, b = (oracle.PackedSequence(s.encode()) for s in p["eq_hash"]["same"])
    c = oracle.PackedSequence(p["eq_hash"]["different"][1].encode())
    assert a == b and a != c and hash(a) == hash(b) and c not in {a}


def test_error_display(kats):
    codes = {v: k for k, v in oracle.VARIANTS.items()}
    for k in kats["error_display"]:
        payload = (k["error"][1:] + [0, 0, 0])[:3]
        e = oracle._Err(codes[k["error"][0]], *payload)
        buf = oracle.C.create_string_buffer(160)
        oracle.lib().orc_error_string(oracle.C.byref(e), buf, 160)
        assert buf.value.decode() == k["text"], k["src"]


def test_split_packed_kats(kats):
    for k in kats["split_packed"]:
        s = k["seq"].encode()
        left, right = oracle.split_packed(oracle.encode_alloc(s), len(s), k["idx"])
        assert (len(left), len(right)) == (k["n_left"], k["n_right"]), k["src"]
        assert oracle.decode_np(left, k["idx"]).tobytes() == k["left"].encode()
        assert oracle.decode_np(right, len(s) - k["idx"]).tobytes() == k["right"].encode()
    for k in kats["split_packed_errors"]:
        s = k["seq"].encode()
        with pytest.raises(OracleError) as ei:
            oracle.split_packed(oracle.encode_alloc(s), len(s), k["idx"])
        expect(ei, k["error"])


def test_synthetic_generator_consistency():
    seed = oracle.DEFAULT_SEED
    for stream in (0, 1, 5):
        w = onp.synth_words(seed, stream, 0, 40)
        assert [oracle.synth_word(seed, stream, j) for j in range(40)] == [int(x) for x in w]
        asc = oracle.synth_ascii(seed, stream, 0, 40 * 32 - 7)
        assert np.array_equal(asc, onp.synth_ascii(seed, stream, 40 * 32 - 7))
        enc = oracle.encode_np(asc)
        w_tail = w.copy()
        w_tail[-1] &= np.uint64((1 << (2 * 25)) - 1)
        assert np.array_equal(enc, w_tail)  # encode(generated stream) == the word stream itself
    assert np.array_equal(oracle.synth_ascii(seed, 2, 100, 77), onp.synth_ascii(seed, 2, 177)[100:])


def test_avx2_restatement_matches_scalar_on_random_input():
    if not oracle.have_avx2():
        pytest.skip("host has no AVX2")
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for n in [1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 1000, 4099]:
        seq = alphabet[rng.integers(0, 8, n)]
        assert np.array_equal(oracle.encode_np(seq, avx2=True), oracle.encode_np(seq))
        for pos in {0, n // 2, n - 1}:
            bad = seq.copy()
            bad[pos] = ord("N")
            keys = []
            for avx2 in (False, True):
                ebuf = []
                with pytest.raises(OracleError) as ei:
                    oracle.encode(bad, ebuf, avx2=avx2)
                keys.append((ei.value.key(), tuple(ebuf)))
            assert keys[0] == keys[1] and len(keys[0][1]) == pos // 32
    t = oracle.bench_codec(alphabet[rng.integers(0, 4, 100000)], threads=2, reps=1)
    assert t > 0


def test_baseline_config0_round_trip_on_cpu():
    """BASELINE.json configs[0]: encode + decode round trip of one 1,000,000-base random sequence on the CPU -- both
    restatements (naive and AVX2 paths) return the input, and the packed words are the generator's own words."""
    import numpy as np
    import oracle
    from oracle import oracle_np as onp
    n = 1_000_000
    seq = onp.synth_ascii(oracle.DEFAULT_SEED, 0, n)
    words = oracle.encode_np(seq, avx2=oracle.have_avx2())
    assert np.array_equal(words, oracle.encode_np(seq)) and words.size == 31_250
    assert [int(w) for w in words[:4]] == [oracle.synth_word(oracle.DEFAULT_SEED, 0, j) for j in range(4)]
    for path in (oracle.PATH_NAIVE, oracle.PATH_AVX2):
        assert np.array_equal(oracle.decode_np(words, n, path=path), seq)
    assert np.array_equal(onp.encode(seq), words)
