"""Property tests (hypothesis) that tie the two CPU restatements of the reference together: the C oracle
(oracle/bitnuc_oracle.c, line by line after the Rust sources) against the independent numpy restatement
(oracle/oracle_np.py) and against the algebraic properties the reference's own tests rely on
(src/utils/mod.rs:64-123: encode -> decode round trips on random lengths; src/utils/functions/hamming: hdist of a
sequence with itself is 0 and is symmetric; split.rs: the two halves decode back to the input).
CPU only; the GPU parity tests (-m gpu) then compare the CUDA path with this oracle."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle
from oracle import OracleError
from oracle import oracle_np as onp

BASES = st.sampled_from(b"ACGTacgt")
seqs = st.lists(BASES, min_size=1, max_size=400).map(bytes)
any_bytes = st.binary(min_size=1, max_size=200)


@settings(max_examples=150, deadline=None)
@given(seqs)
def test_encode_decode_roundtrip_and_restatements_agree(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    words = oracle.encode_np(a)
    assert np.array_equal(words, onp.encode(a))
    assert words.size == (len(seq) + 31) // 32
    if len(seq) % 32:                                   # zero-padded tail (src/lib.rs:96-98)
        assert int(words[-1]) >> (2 * (len(seq) % 32)) == 0
    back = oracle.decode_np(words, len(seq))
    assert back.tobytes() == seq.upper() and np.array_equal(back, onp.decode(words, len(seq)))
    if oracle.have_avx2():
        assert np.array_equal(oracle.encode_np(a, avx2=True), words)


@settings(max_examples=150, deadline=None)
@given(any_bytes)
def test_first_invalid_byte_in_sequence_order(raw):
    a = np.frombuffer(raw, dtype=np.uint8)
    bad = onp.first_invalid(a)
    ebuf = [123]
    if bad is None:
        oracle.encode(raw, ebuf)
        assert len(ebuf) == (len(raw) + 31) // 32
    else:
        pos, byte = bad
        with pytest.raises(OracleError) as ei:
            oracle.encode(raw, ebuf)
        assert ei.value.key() == ("InvalidBase", byte)
        assert len(ebuf) == pos // 32                   # the chunks before the failing chunk (packing/avx.rs:142-143)
        assert ebuf == [int(x) for x in onp.encode(a[: 32 * (pos // 32)])] if pos >= 32 else ebuf == []


@settings(max_examples=100, deadline=None)
@given(seqs, st.data())
def test_hdist_properties(seq, data):
    other = bytes(data.draw(st.lists(BASES, min_size=len(seq), max_size=len(seq))))
    a, b = oracle.encode_np(np.frombuffer(seq, np.uint8)), oracle.encode_np(np.frombuffer(other, np.uint8))
    n = len(seq)
    expect = sum(x != y for x, y in zip(seq.upper(), other.upper()))
    assert oracle.hdist(a, b, n) == oracle.hdist(b, a, n) == expect == onp.hdist(a, b, n)
    assert oracle.hdist(a, a, n) == 0
    per_word = onp.hdist_pairs(a, b, 32) if n % 32 == 0 else None
    if per_word is not None:
        assert int(per_word.sum()) == expect           # checksum of checksums, the full-size GPU property


@settings(max_examples=100, deadline=None)
@given(seqs)
def test_counts_and_gc(seq):
    words = oracle.encode_np(np.frombuffer(seq, np.uint8))
    up = seq.upper()
    counts = [up.count(c) for c in b"ACGT"]
    assert oracle.base_counts(words, len(seq)) == counts == onp.base_counts(words, len(seq))
    gc = (float(counts[1] + counts[2]) / float(len(seq))) * 100.0      # analysis.rs:14, this operation order
    assert oracle.gc_content(words, len(seq)) == gc == onp.gc_content(words, len(seq))


@settings(max_examples=150, deadline=None)
@given(seqs, st.data())
def test_split_packed_halves(seq, data):
    idx = data.draw(st.integers(min_value=0, max_value=len(seq)))
    words = oracle.encode_alloc(seq)
    left, right = oracle.split_packed(words, len(seq), idx)
    assert oracle.decode_np(left, idx).tobytes() == seq.upper()[:idx]
    if idx in (0, len(seq)):
        assert (left, right) == (([], words) if idx == 0 else (words, []))
    else:
        assert len(left) == idx // 32 + 1 and len(right) == len(words) - idx // 32
        first = min(32 - idx % 32, len(seq) - idx)     # the right half's first word always holds the bases after the split
        assert oracle.decode_np(right[:1], first).tobytes() == seq.upper()[idx : idx + first]
    with pytest.raises(OracleError) as ei:
        oracle.split_packed(words, len(seq), len(seq) + 1)
    assert ei.value.key() == ("IndexOutOfBounds", len(seq) + 1, len(seq))


@settings(max_examples=100, deadline=None)
@given(st.integers(min_value=0, max_value=2**64 - 1), st.integers(min_value=0, max_value=32))
def test_from_2bit_as_2bit_inverse(word, k):
    seq = bytes(oracle.from_2bit_alloc(word, k))
    assert len(seq) == k and seq == onp.from_2bit(word, k)
    assert oracle.as_2bit(seq) == word & ((1 << (2 * k)) - 1) == onp.as_2bit(np.frombuffer(seq, np.uint8))
