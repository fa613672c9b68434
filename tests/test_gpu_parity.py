"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's known-answer
vectors.  Every test here needs a B200: run with ``pytest -m gpu``.

The bar is bit-exact for every integer/byte result and exact ``==`` for gc_content (f64 computed in
the reference's operation order from exact integer counts).
"""
import numpy as np
import pytest

import oracle
from oracle import OracleError, PATH_AVX2
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
ACGT_MIXED = np.frombuffer(b"ACGTacgt", dtype=np.uint8)


@pytest.fixture(scope="module")
def bn():
    import bitnuc_b200
    assert bitnuc_b200._lib.load().bn_device_count() >= 1
    return bitnuc_b200


def rand_seq(rng, n, mixed=False):
    a = ACGT_MIXED if mixed else ACGT
    return a[rng.integers(0, a.size, n)]


def gpu_error(bn, fn, *args):
    with pytest.raises(bn.NucleotideError) as ei:
        fn(*args)
    return ei.value


# ------------------------------------------------------------------ known-answer vectors -----

def test_as_2bit_kats(bn, kats, to_int):
    for k in kats["as_2bit"]:
        assert bn.as_2bit(k["seq"].encode()) == to_int(k["packed"]), k["src"]
    for k in kats["as_2bit_equal"]:
        assert bn.as_2bit(k["a"].encode()) == bn.as_2bit(k["b"].encode())
    for k in kats["as_2bit_errors"]:
        seq = k["seq"].encode() if "seq" in k else k["seq_repeat"][0].encode() * k["seq_repeat"][1]
        assert gpu_error(bn, bn.as_2bit, seq).key() == tuple(k["error"]), k["src"]
    assert gpu_error(bn, bn.as_2bit, b"N" * 40).key() == ("SequenceTooLong", 40)  # length outranks content
    assert bn.as_2bit(b"") == 0


def test_from_2bit_kats(bn, kats, to_int):
    for k in kats["from_2bit"]:
        assert bytes(bn.from_2bit_alloc(to_int(k["packed"]), k["n"])) == k["seq"].encode(), k["src"]
    for k in kats["from_2bit_errors"]:
        assert gpu_error(bn, bn.from_2bit_alloc, to_int(k["packed"]), k["n"]).key() == tuple(k["error"])
    k = kats["from_2bit_append"]
    packed, buf = bn.as_2bit(k["seq"].encode()), bytearray()
    for _ in range(k["calls"]):
        bn.from_2bit(packed, k["n"], buf)
    assert bytes(buf) == k["expected"].encode()
    k = kats["from_2bit_prefixes"]
    for n in range(k["lengths"][0], k["lengths"][1] + 1):
        s = k["seq"].encode()[:n]
        assert bytes(bn.from_2bit_alloc(bn.as_2bit(s), n)) == s
    assert bytes(bn.from_2bit_alloc(0xFFFF, 0)) == b""


def test_roundtrip_kats(bn, kats):
    for s in kats["roundtrip_short"]["seqs"]:
        b = s.encode()
        assert bytes(bn.from_2bit_alloc(bn.as_2bit(b), len(b))) == b
    rng = np.random.default_rng(1)
    lo, hi = kats["roundtrip_lengths"]["lengths"]
    for n in range(lo, hi + 1):  # every length 1..=1000, as src/utils/mod.rs:114-133
        seq = rand_seq(rng, n)
        ebuf, dbuf = [], bytearray()
        bn.encode(seq, ebuf)
        assert ebuf == oracle.encode_alloc(seq)
        bn.decode(ebuf, n, dbuf)
        assert bytes(dbuf) == seq.tobytes()


def test_hdist_kats(bn, kats, to_int):
    for k in kats["hdist_scalar"]:
        assert bn.hdist_scalar(to_int(k["u"]), to_int(k["v"]), k["len"]) == k["dist"], k["src"]
    for k in kats["hdist_scalar_seqs"]:
        assert bn.hdist_scalar(bn.as_2bit(k["a"].encode()), bn.as_2bit(k["b"].encode()), len(k["a"])) == k["dist"]
    for k in kats["hdist_scalar_errors"]:
        assert gpu_error(bn, bn.hdist_scalar, to_int(k["u"]), to_int(k["v"]), k["len"]).key() == tuple(k["error"])
    h = kats["hdist"]
    z = [0] * h["too_small"]["n_words"]
    assert gpu_error(bn, bn.hdist, z, z, h["too_small"]["n_bases"]).key() == tuple(h["too_small"]["error"])
    s = h["identical"]["seq_repeat"][0].encode() * h["identical"]["seq_repeat"][1]
    buf = bn.encode_alloc(s)
    assert bn.hdist(buf, buf, len(s)) == 0
    lo, hi = h["a_vs_t_lengths"]["lengths"]
    for n in range(lo, hi + 1):
        assert bn.hdist(bn.encode_alloc(b"A" * n), bn.encode_alloc(b"T" * n), n) == n
    for k in h["mod4_vs_mod3"]:
        s1 = bytes(b"ACGT"[i % 4] for i in range(k["n"]))
        s2 = bytes(b"ACGT"[i % 3] for i in range(k["n"]))
        assert bn.hdist(bn.encode_alloc(s1), bn.encode_alloc(s2), k["n"]) == k["dist"], k["src"]
    a, t = bn.encode_alloc(b"A" * 70), bn.encode_alloc(b"T" * 70)
    assert bn.hdist(a + [7], t + [9, 9], 33) == 33  # extra words ignored, tail masked


def test_analysis_and_packed_sequence_kats(bn, kats):
    for k in kats["analysis"]:
        ps = bn.PackedSequence(k["seq"].encode())
        assert ps.base_counts() == k["counts"], k["src"]
        assert ps.gc_content() == k["gc"], k["src"]
    ps = bn.PackedSequence(b"CAA")
    assert ps.gc_content() == (1.0 / 3.0) * 100.0 != 100.0 * 1.0 / 3.0  # operation order is the reference's
    assert bn.PackedSequence(b"T" * 33).base_counts() == [0, 0, 0, 33]    # padding is not 'A'
    p = kats["packed_sequence"]
    assert gpu_error(bn, bn.PackedSequence, p["new_error"]["seq"].encode()).key() == tuple(p["new_error"]["error"])
    ps = bn.PackedSequence(p["get"]["seq"].encode())
    assert bytes(ps.get(i) for i in range(len(ps))) == p["get"]["values"].encode()
    ps = bn.PackedSequence(p["get_oob"]["seq"].encode())
    assert gpu_error(bn, ps.get, p["get_oob"]["index"]).key() == tuple(p["get_oob"]["error"])
    for k in p["slice"]:
        assert bn.PackedSequence(k["seq"].encode()).slice(k["start"], k["end"]) == k["out"].encode()
    for k in p["slice_errors"]:
        ps = bn.PackedSequence(k["seq"].encode())
        assert gpu_error(bn, ps.slice, k["start"], k["end"]).key() == tuple(k["error"])
    for s in p["to_vec"]:
        ps = bn.PackedSequence(s.encode())
        assert ps.to_vec() == s.encode() and len(ps) == len(s) and ps.is_empty() == (len(s) == 0)
    a, b = (bn.PackedSequence(s.encode()) for s in p["eq_hash"]["same"])
    c = bn.PackedSequence(p["eq_hash"]["different"][1].encode())
    assert a == b and a != c and hash(a) == hash(b) and c not in {a}
    for k in kats["error_display"]:
        assert str(bn.NucleotideError(*k["error"])) == k["text"], k["src"]


# ------------------------------------------------------------------ differential vs oracle ---

@pytest.mark.parametrize("n", [1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 2047, 2048, 2049, 4096 + 7,
                               100_003, 1_000_000, (1 << 22) + 17])
def test_encode_decode_matches_oracle(bn, n):
    rng = np.random.default_rng(n)
    seq = rand_seq(rng, n, mixed=True)
    words = bn.encode_np(seq)
    assert np.array_equal(words, oracle.encode_np(seq))
    assert np.array_equal(words, onp.encode(seq))
    dec = bn.decode_np(words, n)
    assert np.array_equal(dec, oracle.decode_np(words, n, PATH_AVX2))
    # extra words are ignored; partial decodes are prefixes
    for m in {0, 1, n // 2, n - 1}:
        assert np.array_equal(bn.decode_np(np.concatenate([words, words[:1]]), m), dec[:m])


def test_encode_chunked_pipeline_matches_oracle(bn):
    ctx = bn.Context(0)
    ctx.set_chunk_bytes(8192)  # many chunks, ragged last one, all three stages in flight
    rng = np.random.default_rng(5)
    for n in [8192 * 7 + 13, 8192 * 3, 8191, 100_000]:
        seq = rand_seq(rng, n, mixed=True)
        words = bn.encode_np(seq, ctx)
        assert np.array_equal(words, oracle.encode_np(seq))
        assert np.array_equal(bn.decode_np(words, n, ctx), oracle.decode_np(words, n))
        bad = seq.copy()
        pos = [n - 1, 8192 * 2 + 5, 8192 * 2 + 4000]
        for p in pos:
            if p < n:
                bad[p] = ord("N") if p != pos[1] else ord("x")
        ebuf = []
        err = gpu_error(bn, bn.encode, bad, ebuf, ctx)
        first = min(p for p in pos if p < n)
        assert err.key() == ("InvalidBase", int(bad[first]))
        assert ebuf == [int(w) for w in oracle.encode_np(seq)[: first // 32]]
    ctx.close()


def test_encode_error_parity(bn):
    rng = np.random.default_rng(11)
    for n in [1, 5, 31, 32, 33, 64, 100, 1000, 70_001]:
        seq = rand_seq(rng, n)
        positions = sorted({0, n // 2, n - 1, max(0, n - 17), min(n - 1, 31), min(n - 1, 32)})
        for pos in positions:
            for byte in (ord("N"), ord("n"), 0, 255, ord("B"), ord("U"), ord("@"), ord("d"), ord(" ")):
                bad = seq.copy()
                bad[pos] = byte
                ebuf, obuf = [1, 2, 3], []
                err = gpu_error(bn, bn.encode, bad, ebuf)
                with pytest.raises(OracleError) as oe:
                    oracle.encode(bad, obuf)
                assert err.key() == oe.value.key() == ("InvalidBase", byte)
                assert ebuf == obuf and len(ebuf) == pos // 32
    # several invalid bytes: the first in sequence order wins
    bad = rand_seq(rng, 5000)
    bad[[4999, 1234, 1235, 3000]] = [ord("X"), ord("Y"), ord("Z"), ord("N")]
    assert gpu_error(bn, bn.encode_np, bad).key() == ("InvalidBase", ord("Y"))
    with pytest.raises(bn.ReferencePanic):
        bn.encode(b"", [])
    # every byte value, alone in an otherwise valid sequence
    base = rand_seq(rng, 100)
    valid = set(b"ACGTacgt")
    for b in range(256):
        s = base.copy()
        s[37] = b
        if b in valid:
            assert np.array_equal(bn.encode_np(s), oracle.encode_np(s))
        else:
            assert gpu_error(bn, bn.encode_np, s).key() == ("InvalidBase", b)


def test_decode_length_contract(bn):
    w = bn.encode_alloc(b"ACGT" * 16)
    dbuf = bytearray(b"xx")
    bn.decode(w, 40, dbuf)
    assert bytes(dbuf) == b"xx" + b"ACGT" * 10  # appends, never clears
    # short ebuf -> InvalidLength(n_bases), the reference's naive-path contract (unpacking/mod.rs:42-45)
    assert gpu_error(bn, bn.decode, w[:1], 64, bytearray()).key() == ("InvalidLength", 64)
    assert gpu_error(bn, bn.decode, w[:1], 40, bytearray()).key() == ("InvalidLength", 40)
    dbuf = bytearray()
    bn.decode(w, 0, dbuf)
    assert bytes(dbuf) == b""


@pytest.mark.parametrize("k", [1, 2, 5, 8, 15, 16, 17, 21, 31, 32])
@pytest.mark.parametrize("layout", ["tight", "pad32", "pad40", "pad100"])
def test_kmer_batches_match_oracle(bn, k, layout):
    stride = {"tight": k, "pad32": 32, "pad40": 40, "pad100": 100}[layout]
    rng = np.random.default_rng(1000 * k + stride)
    for n in [1, 2, 255, 256, 257, 5000]:
        buf = rng.integers(0, 256, (n - 1) * stride + k).astype(np.uint8)  # garbage between records
        recs = rand_seq(rng, n * k, mixed=True).reshape(n, k)
        for r in range(n):
            buf[r * stride : r * stride + k] = recs[r]
        expect = np.array([oracle.as_2bit(recs[r]) for r in range(n)], dtype=np.uint64)
        got = bn.as_2bit_batch(buf, n, k, stride)
        assert np.array_equal(got, expect)
        # from_2bit: exactly k bytes per record, bytes between records untouched
        noise = rng.integers(0, 2**63, n).astype(np.uint64) << np.uint64(1)  # bits above 2k are ignored
        packed = expect | (noise << np.uint64(2 * k) if k < 32 else np.uint64(0))
        out = np.full((n - 1) * stride + k, ord("#"), dtype=np.uint8)
        res = bn.from_2bit_batch(packed, k, stride, out=out)
        want = np.full((n - 1) * stride + k, ord("#"), dtype=np.uint8)
        for r in range(n):
            want[r * stride : r * stride + k] = np.frombuffer(bytes(oracle.from_2bit_alloc(int(expect[r]), k)), dtype=np.uint8)
        assert np.array_equal(res, want)
        # error parity: first failing record in index order, first bad byte inside it
        if n >= 2:
            bad = buf.copy()
            r_bad = [n - 1, n // 2]
            for r in r_bad:
                bad[r * stride + (k - 1)] = ord("N")
                bad[r * stride + k // 2] = ord("Z") if k // 2 != k - 1 else ord("N")
            err = gpu_error(bn, bn.as_2bit_batch, bad, n, k, stride)
            first_r = min(r_bad)
            with pytest.raises(OracleError) as oe:
                oracle.as_2bit(bad[first_r * stride : first_r * stride + k])
            assert err.key() == oe.value.key()
            assert err.record == first_r and err.offset == first_r * stride + k // 2


def test_kmer_batch_argument_errors(bn):
    assert gpu_error(bn, bn.as_2bit_batch, np.zeros(400, np.uint8), 10, 33, 40).key() == ("SequenceTooLong", 33)
    assert gpu_error(bn, bn.from_2bit_batch, np.zeros(4, np.uint64), 33, 40).key() == ("InvalidLength", 33)
    assert np.array_equal(bn.as_2bit_batch(np.zeros(0, np.uint8), 5, 0, 0), np.zeros(5, np.uint64))


@pytest.mark.parametrize("length", [0, 1, 2, 15, 16, 17, 31, 32])
def test_hdist_pairs_match_oracle(bn, length):
    rng = np.random.default_rng(length)
    for n in [1, 2, 3, 255, 256, 1001, 100_000]:
        u = rng.integers(0, 2**64, n, dtype=np.uint64)
        v = rng.integers(0, 2**64, n, dtype=np.uint64)
        v[::3] = u[::3]
        got = bn.hdist_pairs(u, v, length)
        assert np.array_equal(got, onp.hdist_pairs(u, v, length))
        for i in (0, n // 2, n - 1):
            assert int(got[i]) == oracle.hdist_scalar(int(u[i]), int(v[i]), length)
    assert gpu_error(bn, bn.hdist_pairs, [0], [0], 33).key() == ("InvalidLength", 33)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 4096, 4097, 1_000_003])
def test_hdist_and_counts_match_oracle(bn, n):
    rng = np.random.default_rng(n + 1)
    nw = (n + 31) // 32
    a = rng.integers(0, 2**64, nw + 2, dtype=np.uint64)  # two extra garbage words, unmasked tail bits
    b = a.copy()
    flip = rng.random(nw + 2) < 0.5
    b[flip] = rng.integers(0, 2**64, int(flip.sum()), dtype=np.uint64)
    assert bn.hdist_total(a, b, n) == oracle.hdist(a, b, n, wide=True) == onp.hdist(a, b, n)
    assert bn.hdist(a, b, n) == oracle.hdist(a, b, n)
    counts, gc = bn.base_counts_gc(a, n)
    if n:
        mask = np.uint64((1 << (2 * (n % 32))) - 1) if n % 32 else np.uint64(2**64 - 1)
        clean = a[:nw].copy()
        clean[-1] &= mask
    else:
        clean = a[:0]
    assert counts == oracle.base_counts(clean, n) == onp.base_counts(clean, n)
    assert gc == oracle.gc_content(clean, n)
    assert sum(counts) == n
    if n:
        assert gpu_error(bn, bn.base_counts_gc, a[: nw - 1], n).key() == ("InvalidLength", n)
        assert gpu_error(bn, bn.hdist, a[: nw - 1], b, n).key() == ("InvalidLength", n)


def test_gc_content_is_bit_exact_for_awkward_ratios(bn):
    rng = np.random.default_rng(3)
    for n in [3, 7, 9, 11, 13, 33, 49, 97, 101, 150, 151, 1021]:
        for _ in range(20):
            seq = rand_seq(rng, n)
            ps, ref = bn.PackedSequence(seq), oracle.PackedSequence(seq.tobytes())
            assert ps.gc_content() == ref.gc_content()
            assert ps.base_counts() == ref.base_counts()


@pytest.mark.parametrize("read_len", [150, "mixed", "long"])
def test_base_counts_batch_matches_oracle(bn, read_len):
    rng = np.random.default_rng(17)
    n_reads = 3000 if read_len != "long" else 40
    if read_len == 150:
        lens = np.full(n_reads, 150, dtype=np.uint64)
    elif read_len == "mixed":
        lens = rng.integers(0, 400, n_reads).astype(np.uint64)
        lens[[0, 5, n_reads - 1]] = 0
    else:
        lens = rng.integers(2000, 9000, n_reads).astype(np.uint64)
    nws = (lens + np.uint64(31)) // np.uint64(32)
    wo = np.concatenate([[0], np.cumsum(nws)]).astype(np.uint64)
    words = np.zeros(int(wo[-1]), dtype=np.uint64)
    exp_counts, exp_gc = [], []
    for r in range(n_reads):
        seq = rand_seq(rng, int(lens[r]))
        ps = oracle.PackedSequence(seq.tobytes())
        words[int(wo[r]) : int(wo[r + 1])] = np.array(ps.data, dtype=np.uint64)
        exp_counts.append(ps.base_counts())
        exp_gc.append(ps.gc_content())
    # unmasked garbage above the tail of every read must not matter
    dirty = words.copy()
    for r in range(n_reads):
        rem = int(lens[r]) % 32
        if rem:
            dirty[int(wo[r + 1]) - 1] |= np.uint64(((1 << 64) - 1) ^ ((1 << (2 * rem)) - 1))
    for w in (words, dirty):
        counts, gc, totals = bn.base_counts_batch(w, wo[:-1], lens)
        assert np.array_equal(counts, np.array(exp_counts, dtype=np.uint64).reshape(n_reads, 4))
        assert np.array_equal(gc, np.array(exp_gc))  # exact f64 equality
        assert totals == [int(x) for x in np.array(exp_counts, dtype=np.uint64).sum(axis=0)]


def make_reads(rng, lens, mixed=True):
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rand_seq(rng, int(offsets[-1]), mixed=mixed)
    return data, offsets


@pytest.mark.parametrize("profile", ["short", "cfg5", "with_empties", "single", "mostly_empty", "illumina", "mixed", "short_with_long",
                                     "strip_edge", "many_tiles"])
def test_encode_batch_matches_oracle(bn, profile):
    rng = np.random.default_rng(23)
    lens = {
        "short": rng.integers(1, 70, 4000),
        "cfg5": 50 + rng.integers(0, 9951, 300),   # 50 bp .. 10 kbp, SURVEY.md 8(d) cfg 5
        "with_empties": np.where(rng.random(3000) < 0.3, 0, rng.integers(1, 200, 3000)),
        "single": np.array([100_001]),
        "mostly_empty": np.where(rng.random(20000) < 0.97, 0, rng.integers(1, 40, 20000)),   # > 32 reads start in one group
        "illumina": rng.integers(100, 152, 6000),
        "mixed": np.where(rng.random(2500) < 0.9, rng.integers(20, 300, 2500), rng.integers(300, 20000, 2500)),
        # the one-pass short-read kernel (average <= 192 bytes per read) with rows that do not fit a warp's strip: a long read among
        # short ones, and rows of 32 reads right at the strip's capacity (32 x 256 bytes / 288 words) between rows of small reads
        "short_with_long": np.where(rng.random(6000) < 0.99, rng.integers(50, 151, 6000), rng.integers(2000, 5000, 6000)),
        "strip_edge": np.concatenate([np.concatenate([rng.integers(250, 262, 32), rng.integers(0, 120, 96)]) for _ in range(30)]),
        "many_tiles": rng.integers(0, 70, 150_000),   # 74 tiles of 2048 reads: the look-back chain
    }[profile]
    data, offsets = make_reads(rng, lens)
    words, wo = bn.encode_batch(data, offsets)
    exp_words, exp_wo = [], [0]
    for r in range(len(lens)):
        ps = oracle.PackedSequence(data[int(offsets[r]) : int(offsets[r + 1])].tobytes())
        exp_words.extend(ps.data)
        exp_wo.append(exp_wo[-1] + len(ps.data))
    assert np.array_equal(wo, np.array(exp_wo, dtype=np.uint64))
    assert np.array_equal(words, np.array(exp_words, dtype=np.uint64))
    # the batch may start anywhere in the byte buffer (offsets[0] != 0)
    shifted = np.concatenate([np.full(13, ord("#"), dtype=np.uint8), data])
    w2, wo2 = bn.encode_batch(shifted, offsets + np.uint64(13))
    assert np.array_equal(w2, words) and np.array_equal(wo2, wo)
    # chunked 3-stage pipeline of the host-pointer call: many small chunks of whole reads
    small = bn.Context(0)
    small.set_chunk_bytes(4096)
    w3, wo3 = bn.encode_batch(shifted, offsets + np.uint64(13), ctx=small)
    assert np.array_equal(w3, words) and np.array_equal(wo3, wo)
    # injected N: first invalid base in input order, with read index and position (cfg 5 error parity)
    nonempty = [r for r in range(len(lens)) if lens[r] > 0]
    victims = sorted(set(rng.choice(nonempty, size=min(5, len(nonempty)), replace=False).tolist()))
    bad = data.copy()
    where = {}
    for r in victims:
        pos = int(rng.integers(0, lens[r]))
        bad[int(offsets[r]) + pos] = ord("N")
        where[r] = pos
    err = gpu_error(bn, bn.encode_batch, bad, offsets)
    r0 = victims[0]
    assert err.key() == ("InvalidBase", ord("N"))
    assert (err.record, err.position, err.offset) == (r0, where[r0], int(offsets[r0]) + where[r0])
    _, _, status = bn.encode_batch(bad, offsets, per_read_status=True)
    expect = np.full(len(lens), 0xFFFFFFFF, dtype=np.uint32)
    for r, pos in where.items():
        expect[r] = pos
    assert np.array_equal(status, expect)
    err3 = gpu_error(bn, bn.encode_batch, bad, offsets, small)
    assert (err3.key(), err3.record, err3.position, err3.offset) == (err.key(), r0, where[r0], int(offsets[r0]) + where[r0])
    _, wo4, status4 = bn.encode_batch(bad, offsets, ctx=small, per_read_status=True)
    assert np.array_equal(status4, expect) and np.array_equal(wo4, wo)


# ------------------------------------------------------------------ device-resident path -----

def test_device_resident_codec_properties_at_scale(bn):
    """Size-independent properties on 2^28+17 bases (ragged tail), all on the device:
    encode(synth_ascii) == synth_words with the tail masked; decode(encode(x)) == x."""
    import torch
    from bitnuc_b200 import device as dv
    seed, n = oracle.DEFAULT_SEED, (1 << 28) + 17
    asc = dv.synth_ascii(seed, 0, 0, n)
    words, status = dv.encode(asc)
    status.check()
    expect = dv.synth_words(seed, 0, 0, dv.words_for(n))
    expect[-1] &= (1 << (2 * (n % 32))) - 1
    assert torch.equal(words, expect)
    back = dv.decode(words, n)
    assert torch.equal(back, asc)
    # the generator itself against the oracle on a window
    assert np.array_equal(asc[: 10_000].cpu().numpy(), oracle.synth_ascii(seed, 0, 0, 10_000))
    assert np.array_equal(asc[-4113:].cpu().numpy(), oracle.synth_ascii(seed, 0, n - 4113, 4113))
    # whole-sequence reductions against independent torch arithmetic on the same device data
    counts, gc = dv.base_counts(words, n)
    hist = torch.bincount(asc.to(torch.int64), minlength=256)
    want = [int(hist[c]) for c in b"ACGT"]
    assert counts.tolist() == want
    assert gc.item() == (float(want[1] + want[2]) / float(n)) * 100.0
    other = dv.synth_words(seed, 3, 0, dv.words_for(n))
    other_ascii = dv.decode(other, n)
    assert dv.hdist(words, other, n).item() == int((asc != other_ascii).sum().item())
    # an injected invalid base far into the buffer is located exactly
    asc[(1 << 27) + 12345] = ord("N")
    asc[(1 << 27) + 99999] = ord("Q")
    _, st = dv.encode(asc)
    with pytest.raises(bn.NucleotideError) as ei:
        st.check()
    assert ei.value.key() == ("InvalidBase", ord("N")) and ei.value.offset == (1 << 27) + 12345


def test_device_resident_misaligned_pointers(bn):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(9)
    n = 100_000 + 5
    seq = rand_seq(rng, n + 3, mixed=True)
    t = torch.from_numpy(seq).cuda()
    for shift in (1, 3):
        words, st = dv.encode(t[shift : shift + n])
        st.check()
        assert np.array_equal(words.cpu().numpy().view(np.uint64), oracle.encode_np(seq[shift : shift + n]))
        out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        dv.decode(words, n, out=out[shift : shift + n])
        assert np.array_equal(out[shift : shift + n].cpu().numpy(), oracle.decode_np(words.cpu().numpy().view(np.uint64), n))
    # packed buffers that are only 8-byte aligned
    w = torch.from_numpy(rng.integers(0, 2**63, 5001, dtype=np.int64)).cuda()
    a, b = w[1:2500], w[2501:5000]
    nb = 2499 * 32 - 9
    an, bnp = a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64)
    assert dv.hdist(a, b, nb).item() == onp.hdist(an, bnp, nb)
    assert np.array_equal(dv.hdist_pairs(a, b, 31).cpu().numpy().view(np.uint32), onp.hdist_pairs(an, bnp, 31))
    clean = an.copy()
    clean[-1] &= np.uint64((1 << (2 * (nb % 32))) - 1)
    assert dv.base_counts(a, nb)[0].tolist() == onp.base_counts(clean, nb)


def test_device_resident_kmers_cfg3_shape(bn):
    """cfg 3 at reduced count: 31-mers, record r = first 31 bases of word r of stream 1."""
    import torch
    from bitnuc_b200 import device as dv
    seed, n = oracle.DEFAULT_SEED, (1 << 20) + 3
    words = dv.synth_words(seed, 1, 0, n)
    expect = words & ((1 << 62) - 1)
    recs = dv.from_2bit_batch(words, 31)                   # tight 31-byte records
    assert recs.numel() == 31 * n
    packed, st = dv.as_2bit_batch(recs, n, 31)
    st.check()
    assert torch.equal(packed, expect)
    host = recs[: 31 * 100].cpu().numpy().reshape(100, 31)
    for r in range(100):
        assert host[r].tobytes() == bytes(oracle.from_2bit_alloc(int(expect[r].item()) & (2**64 - 1), 31))
    padded = torch.zeros(32 * n, dtype=torch.uint8, device="cuda")
    dv.from_2bit_batch(words, 32, 32, out=padded)
    p2, st = dv.as_2bit_batch(padded, n, 31, 32)          # padded records, k < stride
    st.check()
    assert torch.equal(p2, expect)


@pytest.mark.parametrize("read_len", [1, 31, 32, 33, 150, 1000])
def test_device_fixed_length_counts_match_oracle(bn, read_len):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(read_len)
    n_reads = 777
    wpr = (read_len + 31) // 32
    words = np.zeros(n_reads * wpr, dtype=np.uint64)
    exp_counts, exp_gc = [], []
    for r in range(n_reads):
        ps = oracle.PackedSequence(rand_seq(rng, read_len).tobytes())
        words[r * wpr : (r + 1) * wpr] = np.array(ps.data, dtype=np.uint64)
        exp_counts.append(ps.base_counts())
        exp_gc.append(ps.gc_content())
    if read_len % 32:  # garbage above the tail must not matter
        words[wpr - 1 :: wpr] |= np.uint64(((1 << 64) - 1) ^ ((1 << (2 * (read_len % 32))) - 1))
    d = torch.from_numpy(words.view(np.int64)).cuda()
    counts, gc, totals = dv.base_counts_fixed(d, n_reads, read_len)
    assert np.array_equal(counts.cpu().numpy(), np.array(exp_counts, dtype=np.int64))
    assert np.array_equal(gc.cpu().numpy(), np.array(exp_gc))
    assert totals.tolist() == np.array(exp_counts).sum(axis=0).tolist()


# ------------------------------------------------------------------ split_packed ----

def test_split_packed_kats(bn, kats):
    for k in kats["split_packed"]:
        s = k["seq"].encode()
        left, right = [123], [456]           # cleared, then filled (split.rs:30-31)
        bn.split_packed(bn.encode_alloc(s), len(s), k["idx"], left, right)
        assert (len(left), len(right)) == (k["n_left"], k["n_right"]), k["src"]
        assert bytes(bn.decode_np(left, k["idx"])) == k["left"].encode()
        assert bytes(bn.decode_np(right, len(s) - k["idx"])) == k["right"].encode()
    for k in kats["split_packed_errors"]:
        s = k["seq"].encode()
        left, right = [123], [456]
        err = gpu_error(bn, bn.split_packed, bn.encode_alloc(s), len(s), k["idx"], left, right)
        assert err.key() == tuple(k["error"])
        assert (left, right) == ([123], [456])  # validation comes before the clears (split.rs:22-31)


def _split_case(rng, n_reads, max_len, extra_words=False):
    lens = rng.integers(0, max_len + 1, n_reads)
    lens[rng.random(n_reads) < 0.05] = 0
    idx = (rng.random(n_reads) * (lens + 1)).astype(np.int64)
    pick = rng.random(n_reads)
    idx = np.where(pick < 0.1, 0, np.where(pick < 0.2, lens, np.where(pick < 0.35, (idx // 32) * 32, idx)))
    idx = np.minimum(idx, lens)
    words, wo = [], [0]
    for r in range(n_reads):
        w = oracle.PackedSequence(rand_seq(rng, int(lens[r])).tobytes()).data
        if extra_words and r % 3 == 0:
            w = list(w) + [int(x) for x in rng.integers(0, 2**63, int(rng.integers(1, 3)))]  # ebuf longer than needed
        words.extend(w)
        wo.append(wo[-1] + len(w))
    return (np.array(words, dtype=np.uint64), np.array(wo, dtype=np.uint64), lens.astype(np.uint64), idx.astype(np.uint64))


def _split_expect(words, wo, lens, idx):
    L, R, lo, ro = [], [], [0], [0]
    for r in range(lens.size):
        left, right = oracle.split_packed(words[int(wo[r]) : int(wo[r + 1])], int(lens[r]), int(idx[r]))
        L.extend(left)
        R.extend(right)
        lo.append(len(L))
        ro.append(len(R))
    return (np.array(L, dtype=np.uint64), np.array(lo, dtype=np.uint64), np.array(R, dtype=np.uint64), np.array(ro, dtype=np.uint64))


@pytest.mark.parametrize("profile", ["barcode", "long", "extra_words", "one"])
def test_split_packed_batch_matches_oracle(bn, profile):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(41)
    n_reads, max_len = {"barcode": (5000, 280), "long": (300, 5000), "extra_words": (2000, 200), "one": (1, 100)}[profile]
    words, wo, lens, idx = _split_case(rng, n_reads, max_len, extra_words=profile == "extra_words")
    exp = _split_expect(words, wo, lens, idx)
    got = bn.split_packed_batch(words, wo, lens, idx)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)
    # device-resident form
    t = [torch.from_numpy(a.view(np.int64)).cuda() for a in (words, wo, lens, idx)]
    left, lo, right, ro, st = dv.split_packed_batch(*t)
    st.check(t[2], t[3])
    assert np.array_equal(lo.cpu().numpy().view(np.uint64), exp[1]) and np.array_equal(ro.cpu().numpy().view(np.uint64), exp[3])
    assert np.array_equal(left[: int(lo[-1])].cpu().numpy().view(np.uint64), exp[0])
    assert np.array_equal(right[: int(ro[-1])].cpu().numpy().view(np.uint64), exp[2])
    # first failing read in index order
    if n_reads > 10:
        bad = idx.copy()
        for r in (n_reads // 2, n_reads // 3):
            bad[r] = lens[r] + np.uint64(1 + r % 5)
        err = gpu_error(bn, bn.split_packed_batch, words, wo, lens, bad)
        r0 = n_reads // 3
        assert err.key() == ("IndexOutOfBounds", int(bad[r0]), int(lens[r0])) and err.record == r0
        tb = torch.from_numpy(bad.view(np.int64)).cuda()
        *_, st = dv.split_packed_batch(t[0], t[1], t[2], tb)
        with pytest.raises(bn.NucleotideError) as ei:
            st.check(t[2], tb)
        assert ei.value.key() == err.key() and ei.value.record == r0


def test_split_packed_empty_batch_and_short_buffer(bn):
    got = bn.split_packed_batch(np.zeros(0, np.uint64), np.zeros(1, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    assert [g.size for g in got] == [0, 1, 0, 1]
    # empty ebuf in the general case: both halves empty (split.rs:45-47)
    left, right = [1], [2]
    bn.split_packed([], 10, 3, left, right)
    assert (left, right) == ([], [])
    # a non-empty ebuf that cannot hold slen bases: the reference panics or truncates; rejected here
    err = gpu_error(bn, bn.split_packed, [0x1B], 100, 40, [], [])
    assert err.key() == ("InvalidLength", 100)


# ------------------------------------------------------------------ get / slice gathers ----

def _packed_batch(rng, lens):
    seqs = [rand_seq(rng, int(n)).tobytes() for n in lens]
    words, wo = [], []
    for s in seqs:
        wo.append(len(words))
        words.extend(oracle.PackedSequence(s).data)
    return seqs, np.array(words, dtype=np.uint64), np.array(wo, dtype=np.uint64), np.asarray(lens, dtype=np.uint64)


@pytest.mark.parametrize("profile", ["windows", "whole_and_empty", "long"])
def test_slice_and_get_batches_match_oracle(bn, profile):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(77)
    lens = {"windows": rng.integers(1, 400, 300), "whole_and_empty": rng.integers(0, 100, 200), "long": rng.integers(5000, 20000, 12)}[profile]
    seqs, words, wo, ln = _packed_batch(rng, lens)
    nq = 2000
    qr = rng.integers(0, len(seqs), nq)
    a, b = (rng.random(nq) * (lens[qr] + 1)).astype(np.int64), (rng.random(nq) * (lens[qr] + 1)).astype(np.int64)
    qs, qe = np.minimum(a, b), np.maximum(a, b)
    if profile == "whole_and_empty":
        qs[::3], qe[::3] = 0, lens[qr[::3]]      # to_vec
        qe[1::3] = qs[1::3]                       # empty ranges
    exp = [seqs[r][s:e] for r, s, e in zip(qr, qs, qe)]
    data, oo = bn.slice_batch(words, wo, ln, qr, qs, qe)
    assert np.array_equal(oo, np.concatenate([[0], np.cumsum([len(x) for x in exp])]).astype(np.uint64))
    assert data.tobytes() == b"".join(exp)
    for q in range(0, nq, 97):  # the reference's own slice on the same PackedSequence
        assert oracle.PackedSequence(seqs[qr[q]]).slice(int(qs[q]), int(qe[q])) == exp[q]
    # device-resident form, arbitrary output alignment comes from the prefix sums
    t = [torch.from_numpy(np.ascontiguousarray(x).astype(np.uint64).view(np.int64)).cuda() for x in (words, wo, ln, qr, qs, qe)]
    d_out, d_oo, st = dv.slice_batch(*t, out_bytes=int(oo[-1]))
    assert st.first_failing() is None
    assert d_out.cpu().numpy().tobytes() == b"".join(exp) and np.array_equal(d_oo.cpu().numpy().view(np.uint64), oo)
    # get
    nz = np.flatnonzero(lens[qr] > 0)
    gi = (rng.random(nz.size) * lens[qr[nz]]).astype(np.int64)
    got = bn.get_batch(words, wo, ln, qr[nz], gi)
    assert got.tobytes() == bytes(seqs[r][i] for r, i in zip(qr[nz], gi))
    assert all(oracle.PackedSequence(seqs[qr[nz][k]]).get(int(gi[k])) == got[k] for k in range(0, nz.size, 131))
    d_get, st = dv.get_batch(t[0], t[1], t[2], t[3][torch.from_numpy(nz).cuda()], torch.from_numpy(gi).cuda())
    assert st.first_failing() is None and np.array_equal(d_get.cpu().numpy(), got)
    # errors: the first failing query in index order, with the reference's payloads
    bad_s, bad_e = qs.copy(), qe.copy()
    q1, q2 = nq // 3, nq // 2
    bad_e[q1] = lens[qr[q1]] + 3                      # end > len
    bad_s[q2], bad_e[q2] = 5, 2                       # start > end
    err = gpu_error(bn, bn.slice_batch, words, wo, ln, qr, bad_s, bad_e)
    assert err.key() == ("InvalidRange", int(bad_s[q1]), int(bad_e[q1]), int(lens[qr[q1]])) and err.record == q1
    with pytest.raises(OracleError) as ei:
        oracle.PackedSequence(seqs[qr[q1]]).slice(int(bad_s[q1]), int(bad_e[q1]))
    assert ei.value.key() == err.key()
    *_, st = dv.slice_batch(t[0], t[1], t[2], t[3], torch.from_numpy(bad_s).cuda(), torch.from_numpy(bad_e).cuda(), out_bytes=int(oo[-1]) + 64)
    assert st.first_failing() == q1
    bad_i = gi.copy()
    bad_i[7] = lens[qr[nz[7]]]
    err = gpu_error(bn, bn.get_batch, words, wo, ln, qr[nz], bad_i)
    assert err.key() == ("IndexOutOfBounds", int(bad_i[7]), int(lens[qr[nz[7]]])) and err.record == 7


def test_slice_kats(bn, kats):
    """PackedSequence doctests / tests of the reference (src/sequence.rs) through the batched gathers."""
    ps = oracle.PackedSequence(b"ACGTACGT")
    w, wo, ln = np.array(ps.data, dtype=np.uint64), np.array([0], dtype=np.uint64), np.array([8], dtype=np.uint64)
    data, _ = bn.slice_batch(w, wo, ln, [0, 0, 0], [1, 0, 8], [4, 8, 8])
    assert data.tobytes() == b"CGT" + b"ACGTACGT"
    assert bn.get_batch(w, wo, ln, [0, 0], [0, 3]).tobytes() == b"AT"
    assert gpu_error(bn, bn.slice_batch, w, wo, ln, [0], [2], [9]).key() == ("InvalidRange", 2, 9, 8)
    assert gpu_error(bn, bn.get_batch, w, wo, ln, [0], [8]).key() == ("IndexOutOfBounds", 8, 8)


# ------------------------------------------------------------------ k-mer windows ----

@pytest.mark.parametrize("k", [1, 4, 15, 16, 17, 21, 31, 32])
def test_kmer_windows_match_oracle(bn, k):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(k)
    for n in [k, k + 1, 100, 2047 + k, 2048 + k, 5000, 70001]:
        seq = rand_seq(rng, n, mixed=True)
        got = bn.kmers(seq, k)
        assert got.size == n - k + 1
        idx = sorted({0, min(1, n - k), n - k, (n - k) // 2, *rng.integers(0, n - k + 1, 40).tolist()})
        assert [int(got[i]) for i in idx] == [oracle.as_2bit(seq[i : i + k]) for i in idx]   # the caller's loop
        enc = oracle.encode_np(seq)                                                        # and all of them, via the packed stream
        j = np.arange(n - k + 1, dtype=np.uint64)
        lo = enc[(j >> np.uint64(5)).astype(np.int64)] >> (np.uint64(2) * (j & np.uint64(31)))
        nxt = np.concatenate([enc[1:], [np.uint64(0)]])[(j >> np.uint64(5)).astype(np.int64)]
        sh = np.uint64(64) - np.uint64(2) * (j & np.uint64(31))
        hi = np.where(sh == np.uint64(64), np.uint64(0), nxt << (sh & np.uint64(63)))
        mask = np.uint64((1 << (2 * k)) - 1)
        assert np.array_equal(got, (lo | hi) & mask)
        for shift in (0, 3):  # device-resident, any alignment
            t = torch.from_numpy(np.concatenate([np.zeros(shift, np.uint8), seq])).cuda()[shift:]
            d, st = dv.kmers(t, k)
            st.check()
            assert np.array_equal(d.cpu().numpy().view(np.uint64), got)


def test_kmer_windows_errors(bn):
    assert bn.kmers(b"ACG", 4).size == 0                       # no window: nothing is looked at, not even ...
    assert bn.kmers(b"NNN", 4).size == 0                       # ... invalid bytes
    assert gpu_error(bn, bn.kmers, b"A" * 40, 33).key() == ("SequenceTooLong", 33)
    assert bn.kmers(b"A" * 10, 33).size == 0
    seq = bytearray(b"ACGT" * 1000)
    seq[2500] = ord("N")
    seq[3000] = ord("x")
    err = gpu_error(bn, bn.kmers, bytes(seq), 21)
    assert err.key() == ("InvalidBase", ord("N")) and err.offset == 2500 and err.record == 2480
    assert [int(x) for x in err.partial] == [oracle.as_2bit(bytes(seq[i : i + 21])) for i in range(2480)]
    with pytest.raises(OracleError) as ei:                      # the caller's loop stops in window 2480 with the same error
        [oracle.as_2bit(bytes(seq[i : i + 21])) for i in range(2480, 2482)]
    assert ei.value.key() == err.key()
    readme = bn.kmers(b"ACGTACGT", 4)                          # README.md:170-180: "ACGT" occurs twice
    assert int((readme == np.uint64(bn.as_2bit(b"ACGT"))).sum()) == 2


# ------------------------------------------------------------------ chunked host-pointer pipelines ----

def test_host_pointer_calls_are_chunk_invariant(bn):
    """Every pipelined host-pointer call gives the same answer with tiny chunks (many pipeline rounds, every stage
    reused) as the oracle: as_2bit / from_2bit batches, hdist, hdist_pairs, base_counts, kmers, incl. error offsets."""
    rng = np.random.default_rng(5)
    ctx = bn.Context(0)
    ctx.set_chunk_bytes(4096)
    n = 50_000
    for k, stride in [(31, 31), (31, 32), (17, 40), (32, 32)]:
        recs = rand_seq(rng, (n - 1) * stride + k, mixed=True)
        got = bn.as_2bit_batch(recs, n, k, stride, ctx=ctx)
        view = np.lib.stride_tricks.as_strided(recs, shape=(n, k), strides=(stride, 1))
        sample = rng.integers(0, n, 200)
        assert [int(got[r]) for r in sample] == [oracle.as_2bit(view[r]) for r in sample]
        assert np.array_equal(got, bn.as_2bit_batch(recs, n, k, stride))          # default (single chunk) context
        out = np.full((n - 1) * stride + k, ord("#"), dtype=np.uint8)
        back = bn.from_2bit_batch(got, k, stride, ctx=ctx, out=out)
        v2 = np.lib.stride_tricks.as_strided(back, shape=(n, k), strides=(stride, 1))
        assert np.array_equal(v2, np.where(view >= 97, view - 32, view))
        if stride > k:
            gaps = np.lib.stride_tricks.as_strided(back[k:], shape=(n - 1, stride - k), strides=(stride, 1))
            assert (gaps == ord("#")).all()                                       # bytes between records untouched
        bad = recs.copy()
        r_bad = 33_333
        bad[r_bad * stride + 5] = ord("N")
        bad[(r_bad + 7000) * stride] = ord("x")
        err = gpu_error(bn, bn.as_2bit_batch, bad, n, k, stride, ctx)
        assert err.key() == ("InvalidBase", ord("N")) and err.record == r_bad and err.offset == r_bad * stride + 5
    nb = 1_000_003
    a, b = rand_seq(rng, nb), rand_seq(rng, nb)
    wa, wb = oracle.encode_np(a), oracle.encode_np(b)
    assert bn.hdist_total(wa, wb, nb, ctx=ctx) == oracle.hdist(wa, wb, nb, wide=True)
    assert np.array_equal(bn.hdist_pairs(wa, wb, 29, ctx=ctx), onp.hdist_pairs(wa, wb, 29))
    counts, gc = bn.base_counts_gc(wa, nb, ctx=ctx)
    assert counts == oracle.base_counts(wa, nb) and gc == oracle.gc_content(wa, nb)
    seq = rand_seq(rng, 100_000, mixed=True)
    assert np.array_equal(bn.kmers(seq, 27, ctx=ctx), bn.kmers(seq, 27))
    seq[77_777] = ord("N")
    err = gpu_error(bn, bn.kmers, seq, 27, ctx)
    assert err.offset == 77_777 and err.record == 77_777 - 26 and err.partial.size == 77_777 - 26
    assert [int(x) for x in err.partial[-50:]] == [oracle.as_2bit(seq[i : i + 27]) for i in range(77_777 - 26 - 50, 77_777 - 26)]


def test_base_counts_batch_layouts(bn):
    """bn_base_counts_batch cuts read tables that are in order into chunks of whole reads (3-stage pipeline); overlapping
    reads stay in order, out-of-order tables take the staged path.  Whatever the context's chunk size, every read equals
    the oracle's PackedSequence."""
    rng = np.random.default_rng(17)
    lens = np.concatenate([rng.integers(0, 400, 3000), [5000, 0, 0, 33, 9000], rng.integers(1, 64, 500)]).astype(np.uint64)
    seqs = [ACGT[rng.integers(0, 4, int(n))].tobytes() for n in lens]
    words, wo = [], [0]
    for s in seqs:
        words += oracle.encode_alloc(s) if s else []
        wo.append(len(words))
    w, wo = np.array(words, dtype=np.uint64), np.array(wo, dtype=np.uint64)
    exp_counts = np.array([oracle.base_counts(oracle.encode_alloc(s), len(s)) if s else [0, 0, 0, 0] for s in seqs], dtype=np.uint64)
    exp_gc = np.array([oracle.gc_content(oracle.encode_alloc(s), len(s)) if s else 0.0 for s in seqs])
    for chunk in (0, 4096, 65536):
        ctx = bn.Context(0)
        if chunk:
            ctx.set_chunk_bytes(chunk)
        counts, gc, totals = bn.base_counts_batch(w, wo, lens, ctx=ctx)
        assert np.array_equal(counts, exp_counts) and np.array_equal(gc, exp_gc)
        assert totals == [int(x) for x in exp_counts.sum(axis=0)]
        # overlapping but ordered reads: read r+1 starts inside read r (every read is a window of one long sequence)
        starts = np.sort(rng.integers(0, max(1, w.size - 40), 2000)).astype(np.uint64)
        olens = np.minimum(rng.integers(1, 1200, 2000), (w.size - starts) * 32).astype(np.uint64)
        c2, g2, t2 = bn.base_counts_batch(w, np.concatenate([starts, [w.size]]).astype(np.uint64), olens, ctx=ctx)
        for r in (0, 1, 999, 1999):
            sl = w[int(starts[r]) : int(starts[r]) + (int(olens[r]) + 31) // 32]
            assert [int(x) for x in c2[r]] == oracle.base_counts(sl, int(olens[r])) and g2[r] == oracle.gc_content(sl, int(olens[r]))
        assert sum(t2) == int(olens.sum())
        # out of order: the staged path
        perm = rng.permutation(len(lens))
        c3, g3, t3 = bn.base_counts_batch(w, np.concatenate([wo[:-1][perm], [w.size]]).astype(np.uint64), lens[perm], ctx=ctx)
        assert np.array_equal(c3, exp_counts[perm]) and np.array_equal(g3, exp_gc[perm]) and t3 == totals
        ctx.close()


def test_base_counts_batch_offsets_table_has_n_reads_entries(bn):
    """Regression: the staged call copied n_reads + 1 offsets into a device buffer sized for n_reads -- with n_reads a
    multiple of 32 (8 n_reads a multiple of the 256-byte allocation granule) the copy failed with a CUDA error, and it
    read one entry past a caller's table of exactly n_reads offsets."""
    import ctypes as C
    from bitnuc_b200._lib import BnError
    ctx = bn.Context(0)
    for n in (32, 64, 4096):
        seqs = [ACGT[np.random.default_rng(r).integers(0, 4, 40 + r % 7)].tobytes() for r in range(n)]
        words, wo = [], []
        for s in seqs:
            wo.append(len(words))
            words += oracle.encode_alloc(s)
        w, wo = np.array(words, dtype=np.uint64), np.array(wo, dtype=np.uint64)       # exactly n offsets
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        counts, gc = np.zeros((n, 4), dtype=np.uint64), np.zeros(n)
        totals, err = (C.c_uint64 * 4)(), BnError()
        rc = ctx.lib.bn_base_counts_batch(ctx.handle, w.ctypes.data, w.size, wo.ctypes.data, lens.ctypes.data, n, counts.ctypes.data,
                                          gc.ctypes.data, totals, C.byref(err))
        assert rc == 0
        assert [int(x) for x in counts[n - 1]] == oracle.base_counts(oracle.encode_alloc(seqs[-1]), len(seqs[-1]))
        assert sum(totals) == int(lens.sum())
    ctx.close()


def test_base_counts_batch_many_chunks(bn):
    """More reads than one 65536-read block, a small chunk size: several pipeline chunks, each a whole number of blocks;
    totals and sampled reads against the oracle, and the first read that does not fit `words` is reported by index."""
    rng = np.random.default_rng(23)
    n = 200_000
    lens = rng.integers(0, 61, n).astype(np.uint64)
    nw = (lens + np.uint64(31)) // np.uint64(32)
    wo = np.concatenate([[0], np.cumsum(nw)]).astype(np.uint64)
    w = rng.integers(0, 1 << 63, int(wo[-1]), dtype=np.int64).view(np.uint64)
    ctx = bn.Context(0)
    ctx.set_chunk_bytes(65536)
    counts, gc, totals = bn.base_counts_batch(w, wo, lens, ctx=ctx)
    assert sum(totals) == int(lens.sum())
    assert np.array_equal(counts.sum(axis=0), np.array(totals, dtype=np.uint64))
    for r in (0, 65535, 65536, 131071, 131072, n - 1):
        sl = w[int(wo[r]) : int(wo[r + 1])]
        assert [int(x) for x in counts[r]] == oracle.base_counts(sl, int(lens[r])) and gc[r] == oracle.gc_content(sl, int(lens[r]))
    big = bn.Context(0)
    c2, g2, t2 = bn.base_counts_batch(w, wo, lens, ctx=big)          # default chunk: the staged path (small input)
    assert np.array_equal(c2, counts) and np.array_equal(g2, gc) and t2 == totals
    bad = lens.copy()
    bad[150_000] = np.uint64(64 * (w.size + 5))
    bad[170_000] = np.uint64(64 * (w.size + 5))
    with pytest.raises(bn.NucleotideError) as ei:
        bn.base_counts_batch(w, wo, bad, ctx=ctx)
    assert ei.value.key() == ("InvalidLength", int(bad[150_000]))
    ctx.close()
    big.close()


# ------------------------------------------------------------------ look-back chains longer than the resident grid ----

def test_split_and_slice_lookback_over_many_tiles(bn):
    """split_packed / slice place their outputs with a single-pass look-back across 256-item tiles (lookback.cuh): far more
    tiles than CTAs fit on the device at once, offsets against numpy prefix sums, spot reads against the oracle."""
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(5)
    n = 1_500_000                                        # ~5900 tiles
    lens = rng.integers(0, 97, n).astype(np.uint64)      # 0..3 words
    nw = (lens + np.uint64(31)) // np.uint64(32)
    wo = np.concatenate([[0], np.cumsum(nw)]).astype(np.uint64)
    words = rng.integers(0, 2**63, int(wo[-1]), dtype=np.int64).view(np.uint64)
    last = wo[1:][nw > 0] - np.uint64(1)                 # zero padding of every read's last word
    rem = (lens[nw > 0] % np.uint64(32)).astype(np.uint64)
    mask = np.where(rem == 0, np.uint64(2**64 - 1), (np.uint64(1) << (np.uint64(2) * rem)) - np.uint64(1))
    words[last.astype(np.int64)] &= mask
    idx = (rng.random(n) * (lens + np.uint64(1)).astype(np.float64)).astype(np.uint64)
    idx = np.minimum(idx, lens)
    t = [torch.from_numpy(a.view(np.int64)).cuda() for a in (words, wo, lens, idx)]
    left, lo, right, ro, st = dv.split_packed_batch(*t)
    st.check(t[2], t[3])
    inner = (idx > 0) & (idx < lens)
    nl = np.where(idx == 0, 0, np.where(idx == lens, nw, np.where(inner & (nw > 0), idx // np.uint64(32) + np.uint64(1), 0))).astype(np.uint64)
    nr = np.where(idx == 0, nw, np.where(idx == lens, 0, np.where(inner & (nw > 0), nw - idx // np.uint64(32), 0))).astype(np.uint64)
    assert np.array_equal(lo.cpu().numpy().view(np.uint64), np.concatenate([[0], np.cumsum(nl)]).astype(np.uint64))
    assert np.array_equal(ro.cpu().numpy().view(np.uint64), np.concatenate([[0], np.cumsum(nr)]).astype(np.uint64))
    h_left, h_right = left.cpu().numpy().view(np.uint64), right.cpu().numpy().view(np.uint64)
    h_lo, h_ro = lo.cpu().numpy().view(np.uint64), ro.cpu().numpy().view(np.uint64)
    for r in list(range(0, n, 7919)) + [n - 1]:
        el, er = oracle.split_packed(words[int(wo[r]): int(wo[r + 1])], int(lens[r]), int(idx[r]))
        assert [int(x) for x in h_left[int(h_lo[r]): int(h_lo[r + 1])]] == list(el), r
        assert [int(x) for x in h_right[int(h_ro[r]): int(h_ro[r + 1])]] == list(er), r
    # slice: one window per read
    qr = np.arange(n, dtype=np.uint64)
    qs = idx // np.uint64(2)
    qe = idx
    tq = [torch.from_numpy(a.view(np.int64)).cuda() for a in (qr, qs, qe)]
    total = int((qe - qs).sum())
    d_out, d_oo, qst = dv.slice_batch(t[0], t[1][:-1].contiguous(), t[2], *tq, out_bytes=total)
    assert qst.first_failing() is None
    oo = d_oo.cpu().numpy().view(np.uint64)
    assert np.array_equal(oo, np.concatenate([[0], np.cumsum(qe - qs)]).astype(np.uint64))
    h_out = d_out.cpu().numpy()
    for r in list(range(0, n, 7919)) + [n - 1]:
        if lens[r]:
            exp = oracle.decode_np(words[int(wo[r]): int(wo[r + 1])], int(lens[r]))[int(qs[r]): int(qe[r])]
            assert np.array_equal(h_out[int(oo[r]): int(oo[r + 1])], exp), r
