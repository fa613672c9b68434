"""the aarch64 paths of the reference where they differ from x86-64 (block-first error byte, push /
overwrite instead of clear / append).  CPU: the plain-Python restatement against cases read off the source; GPU: the
product in aarch64 compatibility mode against that restatement."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp


def test_restatement_on_cases_read_off_the_source():
    # packing/aarch64.rs:223-227: under 32 bases one word is pushed, nothing is cleared
    ebuf = [7]
    onp.encode_aarch64(b"ACGT", ebuf)
    assert ebuf == [7, 0b11100100]
    onp.encode_aarch64(b"", ebuf)
    assert ebuf == [7, 0b11100100, 0]
    with pytest.raises(onp.Aarch64Error) as ei:
        onp.encode_aarch64(b"ACNT", ebuf)
    assert ei.value.key() == ("InvalidBase", ord("N")) and len(ebuf) == 3
    # :235 resize + :185 fill(0) + block loop: the Vec ends up exactly ceil(n/32) long whatever it held
    seq = b"ACGT" * 8 + b"TTTTT"
    ebuf = [1, 2, 3, 4, 5]
    onp.encode_aarch64(seq, ebuf)
    assert ebuf == oracle.encode_alloc(seq)
    # :194-196: a bad byte inside a whole block reports the block's first byte; the words before it stay, the rest are 0
    bad = bytearray(b"C" * 32 + b"G" * 32 + b"T" * 10)
    bad[32 + 17] = ord("N")
    ebuf = [9]
    with pytest.raises(onp.Aarch64Error) as ei:
        onp.encode_aarch64(bytes(bad), ebuf)
    assert ei.value.key() == ("InvalidBase", ord("G")) and ebuf == [oracle.as_2bit(b"C" * 32), 0, 0]
    # :208-214: in the ragged tail the byte itself
    bad = bytearray(b"C" * 32 + b"T" * 10)
    bad[35] = ord("x")
    with pytest.raises(onp.Aarch64Error) as ei:
        onp.encode_aarch64(bytes(bad), [])
    assert ei.value.key() == ("InvalidBase", ord("x"))
    # unpacking/aarch64.rs:127-130: decode overwrites; :113 missing whole words read as zero; :121 the tail panics
    dbuf = bytearray(b"zzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzz")
    onp.decode_aarch64(oracle.encode_alloc(seq), len(seq), dbuf)
    assert bytes(dbuf) == seq
    onp.decode_aarch64([], 64, dbuf)
    assert bytes(dbuf) == b"A" * 64
    with pytest.raises(onp.Aarch64Panic):
        onp.decode_aarch64([0], 40, dbuf)


@pytest.mark.gpu
def test_product_in_aarch64_mode_matches_the_restatement():
    import bitnuc_b200 as bn
    ctx = bn.Context(0)
    assert ctx.compat == "x86_64"
    ctx.set_compat("aarch64")
    assert ctx.compat == "aarch64"
    rng = np.random.default_rng(4)
    al = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for n in [0, 1, 7, 8, 31, 32, 33, 63, 64, 65, 100, 1000, 4097]:
        for trial in range(6):
            seq = bytearray(al[rng.integers(0, 8, n)].tobytes())
            if trial and n:
                for _ in range(int(rng.integers(1, 3))):
                    seq[int(rng.integers(0, n))] = int(rng.choice([ord("N"), 0, 255, ord("@")]))
            pre = [int(x) for x in rng.integers(0, 100, int(rng.integers(0, 5)))]
            exp, got = list(pre), list(pre)
            try:
                onp.encode_aarch64(bytes(seq), exp)
                e_exp = None
            except onp.Aarch64Error as e:
                e_exp = e.key()
            try:
                bn.encode(bytes(seq), got, ctx=ctx)
                e_got = None
            except bn.NucleotideError as e:
                e_got = e.key()
            assert e_got == e_exp and got == exp, (n, trial, e_got, e_exp)
            if e_exp is None and n >= 32:
                dbuf_e, dbuf_g = bytearray(b"previous"), bytearray(b"previous")
                onp.decode_aarch64(exp, n, dbuf_e)
                bn.decode(got, n, dbuf_g, ctx=ctx)
                assert dbuf_g == dbuf_e == bytes(seq).upper()
    # missing words: zeros in whole chunks, a panic in the tail
    d = bytearray()
    bn.decode([], 64, d, ctx=ctx)
    assert bytes(d) == b"A" * 64
    with pytest.raises(bn.ReferencePanic):
        bn.decode([0], 40, d, ctx=ctx)
    # the default mode is untouched on another context
    other = bn.Context(0)
    bad = bytearray(b"G" * 64)
    bad[40] = ord("N")
    with pytest.raises(bn.NucleotideError) as ei:
        bn.encode_alloc(bytes(bad), ctx=other)
    assert ei.value.key() == ("InvalidBase", ord("N"))
    ctx.close()
    other.close()
