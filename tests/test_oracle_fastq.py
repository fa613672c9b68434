"""CPU checks of the FASTQ oracle: the C restatement and the independent Python one against the
hand-written cases of tests/golden/fastq_cases.json, and against each other on random / mutated texts."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

CASES = json.loads((Path(__file__).parent / "golden" / "fastq_cases.json").read_text())


def both(text: bytes, fasta: bool = False):
    try:
        s, l = oracle.fastq_scan(text, fasta)
        a = [[int(x), int(y)] for x, y in zip(s, l)]
    except oracle.FastqFault as e:
        a = ("fault", e.record, e.fault)
    b = onp.fasta_scan(text) if fasta else onp.fastq_scan(text)
    if not isinstance(b, tuple):
        b = [[int(x), int(y)] for x, y in b]
    return a, b


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_cases(case):
    a, b = both(case["text"].encode(), case.get("fasta", False))
    exp = ("fault", *case["fault"]) if "fault" in case else case["reads"]
    assert a == exp
    assert b == exp


def make_fastq(rng, lens, crlf=False, final_newline=True, alphabet=b"ACGT", fasta=False):
    eol = b"\r\n" if crlf else b"\n"
    al = np.frombuffer(alphabet, dtype=np.uint8)
    parts = []
    for r, n in enumerate(lens):
        name = (b">r%d" if fasta else b"@r%d") % r + b" x" * int(rng.integers(0, 4))
        seq = al[rng.integers(0, al.size, int(n))].tobytes()
        qual = bytes(rng.integers(33, 74, int(n)).astype(np.uint8))
        parts += [name, eol, seq, eol] if fasta else [name, eol, seq, eol, b"+", eol, qual, eol]
    text = b"".join(parts)
    if not final_newline and text and int(lens[-1]) > 0:   # an empty last quality line needs its newline to exist at all
        text = text[: -len(eol)]
    return text


@pytest.mark.parametrize("fasta", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_restatements_agree_on_random_and_mutated_texts(seed, fasta):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        lens = rng.integers(0, 90, int(rng.integers(0, 12)))
        text = bytearray(make_fastq(rng, lens, crlf=bool(rng.integers(0, 2)), final_newline=bool(rng.integers(0, 2)), fasta=fasta))
        a, b = both(bytes(text), fasta)
        assert a == b and not isinstance(a, tuple)
        assert [x[1] for x in a] == [int(n) for n in lens]
        for _ in range(4):   # mutations: flip a byte to a newline / '@' / '+' / delete a byte / cut the text
            if not text:
                break
            t = bytearray(text)
            i = int(rng.integers(0, len(t)))
            k = int(rng.integers(0, 5))
            if k == 0:
                t[i] = 10
            elif k == 1:
                t[i] = ord("@")
            elif k == 2:
                del t[i]
            elif k == 3:
                t = t[:i]
            else:
                t[i] = ord("x")
            a, b = both(bytes(t), fasta)
            assert a == b


def test_fastq_encode_is_the_callers_loop():
    rng = np.random.default_rng(3)
    lens = [0, 1, 31, 32, 33, 64, 150]
    text = make_fastq(rng, lens, alphabet=b"ACGTacgt")
    words, wo, so, sl = oracle.fastq_encode(text)
    assert [int(x) for x in sl] == lens
    for r, n in enumerate(lens):
        seq = text[int(so[r]) : int(so[r]) + n]
        exp = oracle.encode_alloc(seq) if n else []
        assert [int(x) for x in words[int(wo[r]) : int(wo[r + 1])]] == exp
    bad = bytearray(text)
    bad[int(so[4]) + 7] = ord("N")
    with pytest.raises(oracle.OracleError) as ei:
        oracle.fastq_encode(bytes(bad))
    assert ei.value.key() == ("InvalidBase", ord("N"))


@pytest.mark.parametrize("fasta", [False, True])
def test_one_call_cpu_form_equals_the_callers_loop(fasta):
    """orc_fastx_encode (the timed CPU baseline of the FASTQ / FASTA row) against fastq_scan + per-record encode."""
    rng = np.random.default_rng(31 + fasta)
    text = make_fastq(rng, rng.integers(0, 300, 500), crlf=bool(fasta), alphabet=b"ACGTacgt", fasta=fasta)
    _, w, wo = oracle.fastx_encode_timed(text, fasta, reps=1)
    ew, ewo, _, _ = oracle.fastq_encode(text, fasta)
    assert np.array_equal(w, ew) and np.array_equal(wo, ewo)
    _, w2, _ = oracle.fastx_encode_timed(text, fasta, path=oracle.PATH_NAIVE, reps=1)
    assert np.array_equal(w2, ew)
    bad = bytearray(text)
    bad[bad.find(b"\n") + 2] = ord("N")
    with pytest.raises(oracle.OracleError) as ei:
        oracle.fastx_encode_timed(bytes(bad), fasta, reps=1)
    assert ei.value.key() == ("InvalidBase", ord("N"))
    with pytest.raises(oracle.FastqFault):
        oracle.fastx_encode_timed(text[1:], fasta, reps=1)


def test_wrapped_fasta_definition():
    """oracle.fasta_wrapped_encode on hand-written texts: joined lines, empty records, CRLF, a missing final newline."""
    w, wo, ho, sl = oracle.fasta_wrapped_encode(b">chr1 test\nACGTAC\nGTTT\n\nAC\n>chr2\n>chr3\r\nacgt\r\nTTGG")
    assert wo.tolist() == [0, 1, 1, 2] and ho.tolist() == [0, 27, 33] and sl.tolist() == [12, 0, 8]
    assert [int(x) for x in w] == [oracle.as_2bit(b"ACGTACGTTTAC"), oracle.as_2bit(b"ACGTTTGG")]
    assert oracle.fasta_wrapped_encode(b"")[1].tolist() == [0]
    w, wo, ho, sl = oracle.fasta_wrapped_encode(b">a\n" + b"A" * 31 + b"\n" + b"C" * 33 + b"\nG\r")   # a trailing '\r' is a line end's
    assert sl.tolist() == [65] and wo.tolist() == [0, 3]
    assert oracle.decode_np(w, 65).tobytes() == b"A" * 31 + b"C" * 33 + b"G"
    import pytest
    with pytest.raises(oracle.FastqFault) as ei:
        oracle.fasta_wrapped_encode(b"\n>a\nAC\n")          # the text must open with a header
    assert (ei.value.record, ei.value.fault) == (0, 1)
    with pytest.raises(oracle.OracleError) as eo:
        oracle.fasta_wrapped_encode(b">a\nAC\n>b\nAC\nGN\n")
    assert eo.value.key() == ("InvalidBase", ord("N")) and (eo.value.record, eo.value.position) == (1, 3)
