"""Wrapped (multi-line) FASTA through bn_fasta_wrapped_scan / bn_fasta_wrapped_encode against the oracle's definition
(oracle.fasta_wrapped_encode): hand-written cases, random genome-style texts (wrapped at 60 / 70 / 80 columns, LF and CRLF,
empty records, empty lines, missing final newline, lower case), faults and invalid bases."""
import numpy as np
import pytest

import oracle
from oracle import OracleError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    import bitnuc_b200 as b
    return b


def _same(bn, text: bytes):
    exp = oracle.fasta_wrapped_encode(text)
    got = bn.fasta_wrapped_encode(np.frombuffer(text, dtype=np.uint8))
    for g, e, what in zip(got, exp, ("words", "word_offsets", "header_offsets", "seq_lens")):
        assert np.array_equal(g, e), what
    return exp


CASES = [
    b">chr1 test\nACGTAC\nGTTT\n\nAC\n>chr2\n>chr3\r\nacgt\r\nTTGG",          # empty line, empty record, CRLF, no final newline
    b">a\nA\n",
    b">a\n",
    b">only header",
    b">a\n" + b"ACGT" * 8 + b"\n" + b"ACGT" * 8 + b"\n",                          # two lines that fill one word each
    b">a\n" + b"A" * 31 + b"\n" + b"C" * 33 + b"\n" + b"G\n",                     # lines that straddle word boundaries
    b">x\r\n\r\nAC\r\n\r\nGT\r\n",
    b"",
]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_hand_written_cases(bn, i):
    _same(bn, CASES[i])


def _genome(rng, n_records, width, crlf, final_newline=True, lower=0.0):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for r in range(n_records):
        out.append(b">seq%d some description" % r + nl)
        n = int(rng.integers(0, 3000)) if rng.random() > 0.1 else 0
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        if lower:
            m = rng.random(n) < lower
            seq[m] |= 0x20
        seq = seq.tobytes()
        for o in range(0, n, width):
            out.append(seq[o:o + width] + nl)
        if rng.random() < 0.05:
            out.append(nl)                                   # a blank line inside / after a record
    text = b"".join(out)
    return text if final_newline else text.rstrip(b"\r\n")


@pytest.mark.parametrize("width,crlf,final_newline", [(60, False, True), (70, True, True), (80, False, False), (61, True, False), (7, False, True)])
def test_random_genome_style_texts(bn, width, crlf, final_newline):
    rng = np.random.default_rng(width * 7 + crlf)
    text = _genome(rng, 400, width, crlf, final_newline, lower=0.2)
    words, wo, ho, sl = _same(bn, text)
    assert wo[-1] == words.size and sl.size == 400
    # the headers are where the oracle says, and the joined sequences decode back
    for r in (0, 7, 399):
        assert text[int(ho[r]):int(ho[r]) + 4] == b">seq"


def test_large_text_many_tiles(bn):
    rng = np.random.default_rng(3)
    text = _genome(rng, 6000, 60, False)                     # ~9 MB: hundreds of 16 KiB tiles, lines crossing tile boundaries
    _same(bn, text)


def test_faults_and_invalid_bases(bn):
    import bitnuc_b200 as b
    with pytest.raises(b.FastqError) as ei:
        bn.fasta_wrapped_encode(np.frombuffer(b"ACGT\n>x\nAC\n", dtype=np.uint8))
    assert ei.value.key() == (0, 1)
    with pytest.raises(oracle.FastqFault):
        oracle.fasta_wrapped_encode(b"ACGT\n>x\nAC\n")
    rng = np.random.default_rng(11)
    text = bytearray(_genome(rng, 50, 60, False))
    exp = oracle.fasta_wrapped_encode(bytes(text))
    # an N in the third line of some record: first in file order wins, with the record and the position in the joined sequence
    ho, sl = exp[2], exp[3]
    victims = [r for r in range(50) if sl[r] > 200][:3]
    where = {}
    for r in victims[::-1]:
        hdr_end = text.index(b"\n", int(ho[r])) + 1
        p = hdr_end + 2 * 61 + 5                              # line 3, column 6
        text[p] = ord("N")
        where[r] = 2 * 60 + 5
    r0 = victims[0]
    with pytest.raises(b.NucleotideError) as ei:
        bn.fasta_wrapped_encode(np.frombuffer(bytes(text), dtype=np.uint8))
    assert ei.value.key() == ("InvalidBase", ord("N")) and (ei.value.record, ei.value.position) == (r0, where[r0])
    with pytest.raises(OracleError) as eo:
        oracle.fasta_wrapped_encode(bytes(text))
    assert eo.value.key() == ei.value.key() and (eo.value.record, eo.value.position) == (r0, where[r0])


@pytest.mark.parametrize("kind", ["dense_lines", "tile_edges", "one_long_line", "headers_only", "crlf_dense", "mixed"])
def test_adversarial_shapes(bn, kind):
    """Shapes chosen against the implementation: more than 2048 lines in a 16 KiB tile (the slot rows overflow: the dense index must
    carry the result), newlines and headers on tile boundaries, one line of several tiles, records without sequence."""
    rng = np.random.default_rng(sum(kind.encode()))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def seq(n):
        return acgt[rng.integers(0, 4, n)].tobytes()

    if kind == "dense_lines":                       # 1-3 byte lines: ~8000 lines per tile
        text = b">d\n" + b"".join(seq(int(rng.integers(0, 3))) + b"\n" for _ in range(60000)) + b">e\n" + seq(100) + b"\n"
    elif kind == "crlf_dense":
        text = b">d\r\n" + b"".join(seq(int(rng.integers(0, 2))) + b"\r\n" for _ in range(40000))
    elif kind == "tile_edges":                      # every record is exactly one 16 KiB tile: the header starts on the boundary
        recs = []
        for r in range(12):
            hdr = b">r%02d\n" % r
            body = 16384 - len(hdr)
            lines = [seq(62) + b"\n" for _ in range(body // 63)]
            last = body - 63 * (body // 63)
            recs.append(hdr + b"".join(lines) + (seq(last - 1) + b"\n" if last else b""))
        text = b"".join(recs)
        assert len(text) == 12 * 16384
    elif kind == "one_long_line":                   # a sequence line spanning ~20 tiles, then short ones
        text = b">long\n" + seq(333_333) + b"\n" + seq(10) + b"\n>short\n" + seq(5)
    elif kind == "headers_only":
        text = b"".join(b">h%d\n" % i for i in range(5000))
    else:
        parts = []
        for r in range(300):
            parts.append(b">m%d\n" % r)
            for _ in range(int(rng.integers(0, 6))):
                parts.append(seq(int(rng.choice([0, 1, 15, 16, 17, 31, 32, 33, 60, 500, 20000]))) + (b"\r\n" if rng.random() < 0.3 else b"\n"))
        text = b"".join(parts)
    _same(bn, text)
