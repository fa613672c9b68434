"""Independent numpy restatement of the bitnuc hot path.  TEST INFRASTRUCTURE ONLY.

A second, array-at-a-time statement of the same semantics as ``bitnuc_oracle.c``; the two are
written differently on purpose (lookup tables + reshapes here, per-base loops there) and tests
require them to agree with each other and with the reference's known-answer vectors.  Citations are
relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)  # src/utils/packing/naive.rs:10-16
for _ch, _v in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"Tt", 3)):
    for _b in _ch:
        _CODE[_b] = _v
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)  # src/utils/unpacking/naive.rs:15-21
_SHIFTS = (2 * np.arange(32, dtype=np.uint64)).reshape(1, 32)
M64 = (1 << 64) - 1


class NpError(Exception):
    def __init__(self, variant: str, *payload: int):
        super().__init__(f"{variant}{payload}")
        self.variant, self.payload = variant, payload

    def key(self):
        return (self.variant,) + tuple(self.payload)


def first_invalid(seq: np.ndarray):
    """(offset, byte) of the first non-ACGTacgt byte, or None."""
    bad = np.flatnonzero(_CODE[seq] == 255)
    return (int(bad[0]), int(seq[bad[0]])) if bad.size else None


def encode(seq) -> np.ndarray:
    """src/utils/packing/naive.rs:22-43 -- ceil(n/32) LSB-first words, zero-padded tail."""
    seq = np.frombuffer(bytes(seq), dtype=np.uint8) if not isinstance(seq, np.ndarray) else seq
    if seq.size == 0:
        raise ZeroDivisionError("reference panics on encode(b'')")  # avx.rs:138
    inv = first_invalid(seq)
    if inv is not None:
        raise NpError("InvalidBase", inv[1])
    n_words = (seq.size + 31) // 32
    codes = np.zeros(n_words * 32, dtype=np.uint64)
    codes[: seq.size] = _CODE[seq]
    return np.bitwise_or.reduce(codes.reshape(n_words, 32) << _SHIFTS, axis=1)


def as_2bit(seq) -> int:
    seq = np.frombuffer(bytes(seq), dtype=np.uint8)
    if seq.size > 32:
        raise NpError("SequenceTooLong", seq.size)  # checked before content, naive.rs:5-7
    if seq.size == 0:
        return 0
    return int(encode(seq)[0])


def decode(words, n_bases: int) -> np.ndarray:
    """src/utils/unpacking/avx.rs:117-153 for well-formed input (enough words)."""
    words = np.asarray(words, dtype=np.uint64)
    if words.size < (n_bases + 31) // 32:
        raise NpError("InvalidLength", n_bases)  # src/utils/unpacking/mod.rs:42-45
    w = words[: (n_bases + 31) // 32].reshape(-1, 1)
    return _ASCII[((w >> _SHIFTS) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n_bases]]


def from_2bit(packed: int, expected_size: int) -> bytes:
    if expected_size > 32:
        raise NpError("InvalidLength", expected_size)
    return decode(np.array([packed & M64], dtype=np.uint64), expected_size).tobytes()


def _popcount64(x: np.ndarray) -> np.ndarray:
    return np.unpackbits(x.view(np.uint8).reshape(-1, 8), axis=1).sum(axis=1, dtype=np.uint64)


def hdist_pairs(u, v, length: int) -> np.ndarray:
    """src/utils/functions/hamming/scalar.rs:11-48, one result per (u[i], v[i])."""
    if length > 32:
        raise NpError("InvalidLength", length)
    u, v = np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)
    mask = np.uint64(M64 if length == 32 else (1 << (2 * length)) - 1)
    d = (u ^ v) & mask
    d = (d | (d >> np.uint64(1))) & np.uint64(0x5555555555555555)
    return _popcount64(np.ascontiguousarray(d)).astype(np.uint32)


def hdist(e1, e2, n_bases: int) -> int:
    """src/utils/functions/hamming/multi.rs:122-160; exact (unwrapped) total."""
    e1, e2 = np.asarray(e1, dtype=np.uint64), np.asarray(e2, dtype=np.uint64)
    need = (n_bases + 31) // 32
    if e1.size < need or e2.size < need:
        raise NpError("InvalidLength", n_bases)
    full, rem = divmod(n_bases, 32)
    total = int(hdist_pairs(e1[:full], e2[:full], 32).sum(dtype=np.uint64)) if full else 0
    if rem:
        total += int(hdist_pairs(e1[full : full + 1], e2[full : full + 1], rem)[0])
    return total


def base_counts(words, length: int) -> list:
    """src/utils/analysis.rs:19-39 -- counts over the `length` decoded bases only."""
    if length == 0:
        return [0, 0, 0, 0]
    asc = decode(words, length)
    return [int((asc == ch).sum()) for ch in b"ACGT"]


def gc_content(words, length: int) -> float:
    """src/utils/analysis.rs:3-17 -- (gc as f64 / len as f64) * 100.0 in that order."""
    if length == 0:
        return 0.0
    c = base_counts(words, length)
    return float((np.float64(c[1] + c[2]) / np.float64(length)) * np.float64(100.0))


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_words(seed: int, stream: int, first_word: int, n_words: int) -> np.ndarray:
    """word j of stream s = splitmix64((seed ^ s*golden) + j)."""
    base = (seed ^ ((stream * 0x9E3779B97F4A7C15) & M64)) & M64
    with np.errstate(over="ignore"):
        j = np.arange(first_word, first_word + n_words, dtype=np.uint64) + np.uint64(base)
    return splitmix64(j)


def synth_ascii(seed: int, stream: int, n_bases: int) -> np.ndarray:
    """ASCII bases 0..n of a stream; by construction encode(synth_ascii) == synth_words (tail masked)."""
    w = synth_words(seed, stream, 0, (n_bases + 31) // 32)
    return decode(w, n_bases)


def fasta_scan(text: bytes):
    """Independent restatement of orc_fasta_scan (one sequence line per record)."""
    return fastq_scan(text, lines_per_record=2)


def fastq_scan(text: bytes, lines_per_record: int = 4):
    """Independent restatement of orc_fastq_scan with Python's own line splitting: [(start, length)] of the sequence
    lines, or ("fault", record, kind)."""
    text = bytes(text)
    lpr, hdr = lines_per_record, (b"@" if lines_per_record == 4 else b">")
    lines, pos = [], 0
    while pos < len(text):
        e = text.find(b"\n", pos)
        if e < 0:
            e = len(text)
        ln = e - pos
        if ln and text[e - 1 : e] == b"\r":
            ln -= 1
        lines.append((pos, ln))
        pos = e + 1
    out = []
    for r in range((len(lines) + lpr - 1) // lpr):
        rec = lines[lpr * r : lpr * r + lpr]
        if text[rec[0][0] : rec[0][0] + 1] != hdr:
            return ("fault", r, 1)
        if lpr == 4 and len(rec) >= 3 and text[rec[2][0] : rec[2][0] + 1] != b"+":
            return ("fault", r, 2)
        if lpr == 4 and len(rec) == 4 and rec[3][1] != rec[1][1]:
            return ("fault", r, 3)
        if len(rec) < lpr:
            return ("fault", r, 4)
        out.append(rec[1])
    return out


# ---- the aarch64 paths where they differ from x86-64: plain-Python restatements ----------------

_A64_VALID = b"ACGTacgt"
_A64_CODE = {65: 0, 97: 0, 67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}


class Aarch64Error(Exception):
    def __init__(self, variant, value):
        super().__init__(f"{variant}({value})")
        self.variant, self.value = variant, value

    def key(self):
        return (self.variant, self.value)


class Aarch64Panic(Exception):
    pass


def _as_2bit_aarch64(seq: bytes) -> int:
    """src/utils/packing/aarch64.rs:76-128: length first, then the first invalid byte in order."""
    if len(seq) > 32:
        raise Aarch64Error("SequenceTooLong", len(seq))
    for b in seq:
        if b not in _A64_VALID:
            raise Aarch64Error("InvalidBase", b)
    return sum(_A64_CODE[b] << (2 * i) for i, b in enumerate(seq))


def encode_aarch64(seq: bytes, ebuf: list) -> None:
    """src/utils/mod.rs:22-25 -> src/utils/packing/aarch64.rs:222-244 (encode_internal) and :173-219
    (encode_nucleotides_simd).  Nothing is cleared: a short sequence PUSHES one word; a long one resizes the Vec to
    ceil(n/32), zero-fills it and overwrites block by block; a bad byte inside a whole block reports the block's
    FIRST byte (:194-196), one in the tail reports itself (:208-214)."""
    seq = bytes(seq)
    if len(seq) < 32:
        ebuf.append(_as_2bit_aarch64(seq))        # :223-227 (`?` leaves ebuf untouched on error)
        return
    n_chunks = (len(seq) + 31) // 32
    del ebuf[n_chunks:]                           # ebuf.resize(n_chunks, 0), :235
    ebuf.extend([0] * (n_chunks - len(ebuf)))
    ebuf[:] = [0] * n_chunks                      # output.fill(0), :185
    full = len(seq) // 32
    for blk in range(full):
        chunk = seq[32 * blk : 32 * blk + 32]
        if any(b not in _A64_VALID for b in chunk):
            raise Aarch64Error("InvalidBase", chunk[0])      # InvalidBase(*ip), :194-196
        ebuf[blk] = sum(_A64_CODE[b] << (2 * i) for i, b in enumerate(chunk))
    tail = seq[32 * full :]
    if tail:
        word = 0
        for i, b in enumerate(tail):
            if (b | 0x20) not in b"acgt":
                raise Aarch64Error("InvalidBase", b)         # :213
            word |= _A64_CODE[b] << (2 * i)
        ebuf[full] = word


def decode_aarch64(ebuf, n_bases: int, dbuf: bytearray) -> None:
    """src/utils/unpacking/aarch64.rs:127-130 (fast_decode: out.resize(len, 0), then overwritten) and :100-125
    (decode_nucleotides_simd: whole chunks read `input.get(i).copied().unwrap_or(0)`, the tail indexes input[j / 32])."""
    ebuf = [int(w) for w in ebuf]
    out = bytearray(n_bases)
    chunks = n_bases // 32
    for i in range(chunks):
        w = ebuf[i] if i < len(ebuf) else 0
        for j in range(32):
            out[32 * i + j] = b"ACGT"[(w >> (2 * j)) & 3]
    for j in range(32 * chunks, n_bases):
        if j // 32 >= len(ebuf):
            raise Aarch64Panic("index out of bounds")
        out[j] = b"ACGT"[(ebuf[j // 32] >> (2 * (j % 32))) & 3]
    dbuf[:] = out
