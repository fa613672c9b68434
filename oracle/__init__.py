"""ctypes front-end of the CPU oracle (oracle/bitnuc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  ``bitnuc_b200`` never does: the product path is CUDA-only.

The wrappers mirror the reference's Rust signatures (``Vec`` arguments become Python lists /
``bytearray`` that are cleared or appended to exactly as the reference does) so that tests can be
written the way the reference's own tests are.  Errors surface as :class:`OracleError` carrying the
``NucleotideError`` variant; inputs on which the reference panics raise :class:`OraclePanic`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libbitnuc_oracle.so"

PATH_NAIVE, PATH_AVX2 = 0, 1

VARIANTS = {
    1: "InvalidBase",
    2: "SequenceTooLong",
    3: "InvalidLength",
    4: "IndexOutOfBounds",
    5: "InvalidRange",
    6: "Unsupported",
}


class OracleError(Exception):
    def __init__(self, code: int, a: int, b: int, c: int, text: str):
        super().__init__(text)
        self.code, self.a, self.b, self.c = code, a, b, c
        self.variant = VARIANTS.get(code, "?")

    def key(self):
        n = {1: 1, 2: 1, 3: 1, 4: 2, 5: 3, 6: 0}[self.code]
        return (self.variant,) + (self.a, self.b, self.c)[:n]


class OraclePanic(Exception):
    """The reference would panic (index out of bounds / arithmetic underflow) on this input."""


class _Err(C.Structure):
    _fields_ = [("code", C.c_int32), ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64)]


class _BenchDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("path", C.c_int32), ("n_units", C.c_uint64), ("k", C.c_uint64), ("stride", C.c_uint64),
                ("in0", C.c_void_p), ("in1", C.c_void_p), ("out0", C.c_void_p), ("out1", C.c_void_p)]


OP_ENCODE, OP_DECODE, OP_AS_2BIT, OP_FROM_2BIT, OP_HDIST, OP_HDIST_PAIRS, OP_BASE_COUNTS_GC, OP_ENCODE_BATCH = range(8)


def build(native: bool = False, force: bool = False) -> Path:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = _HERE / "bitnuc_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "native" if native else "all"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u8p, u64p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
        ep = C.POINTER(_Err)
        sz, u64 = C.c_size_t, C.c_uint64
        sig = {
            "orc_error_string": (C.c_int, [ep, C.c_char_p, sz]),
            "orc_as_2bit": (C.c_int, [C.c_void_p, sz, u64p, ep]),
            "orc_as_2bit_avx2": (C.c_int, [C.c_void_p, sz, u64p, ep]),
            "orc_encode": (C.c_int, [C.c_void_p, sz, C.c_void_p, szp, ep]),
            "orc_encode_avx2": (C.c_int, [C.c_void_p, sz, C.c_void_p, szp, ep]),
            "orc_from_2bit": (C.c_int, [u64, sz, C.c_void_p, ep]),
            "orc_from_2bit_avx2": (C.c_int, [u64, sz, C.c_void_p, ep]),
            "orc_decode": (C.c_int, [C.c_void_p, sz, sz, C.c_void_p, szp, C.c_int, ep]),
            "orc_hdist_scalar": (C.c_int, [u64, u64, sz, C.POINTER(C.c_uint32), ep]),
            "orc_hdist": (C.c_int, [C.c_void_p, sz, C.c_void_p, sz, sz, C.POINTER(C.c_uint32), u64p,
                                    C.c_int, ep]),
            "orc_seq_get": (C.c_int, [C.c_void_p, sz, sz, u8p, ep]),
            "orc_seq_slice": (C.c_int, [C.c_void_p, sz, sz, sz, C.c_void_p, ep]),
            "orc_base_counts": (None, [C.c_void_p, sz, u64p]),
            "orc_gc_content": (C.c_double, [C.c_void_p, sz]),
            "orc_split_packed": (C.c_int, [C.c_void_p, sz, sz, sz, C.c_void_p, szp, C.c_void_p, szp, ep]),
            "orc_fastq_scan": (C.c_int, [C.c_void_p, sz, C.c_void_p, C.c_void_p, sz, szp, u64p, C.POINTER(C.c_int)]),
            "orc_fasta_scan": (C.c_int, [C.c_void_p, sz, C.c_void_p, C.c_void_p, sz, szp, u64p, C.POINTER(C.c_int)]),
            "orc_fastx_encode": (C.c_int, [C.c_void_p, sz, C.c_int, C.c_int, C.c_void_p, sz, C.c_void_p, sz, szp, u64p, C.POINTER(C.c_int), ep]),
            "orc_splitmix64": (u64, [u64]),
            "orc_synth_word": (u64, [u64, u64, u64]),
            "orc_synth_ascii": (None, [u64, u64, u64, sz, C.c_void_p]),
            "orc_bench_codec": (C.c_double, [C.c_void_p, sz, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p]),
            "orc_have_avx2": (C.c_int, []),
            "orc_bench_op": (C.c_int, [C.POINTER(_BenchDesc), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), u64p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _raise(rc: int, e: _Err):
    if rc == 0:
        return
    if rc == -100:
        raise OraclePanic()
    buf = C.create_string_buffer(160)
    lib().orc_error_string(C.byref(e), buf, 160)
    raise OracleError(e.code, e.a, e.b, e.c, buf.value.decode())


def _bytes_arr(seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    return np.frombuffer(bytes(seq), dtype=np.uint8)


def _words_arr(words) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(words, dtype=np.uint64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


def have_avx2() -> bool:
    return bool(lib().orc_have_avx2())


# ------------------------------------------------------------------ reference-shaped API ------

def as_2bit(seq, avx2: bool = False) -> int:
    a = _bytes_arr(seq)
    out, e = C.c_uint64(0), _Err()
    fn = lib().orc_as_2bit_avx2 if avx2 else lib().orc_as_2bit
    _raise(fn(_ptr(a), a.size, C.byref(out), C.byref(e)), e)
    return out.value


def encode(seq, ebuf: list, avx2: bool = False) -> None:
    """``encode(&[u8], &mut Vec<u64>)``: clears ``ebuf`` first; on error it keeps the words of the
    chunks before the failing one (/root/reference/src/utils/packing/avx.rs:132,142-143)."""
    a = _bytes_arr(seq)
    words = np.zeros(max(1, (a.size + 31) // 32), dtype=np.uint64)
    n, e = C.c_size_t(0), _Err()
    fn = lib().orc_encode_avx2 if avx2 else lib().orc_encode
    rc = fn(_ptr(a), a.size, _ptr(words), C.byref(n), C.byref(e))
    if rc != -100:
        ebuf.clear()
        ebuf.extend(int(x) for x in words[: n.value])
    _raise(rc, e)


def encode_alloc(seq, avx2: bool = False) -> list:
    ebuf: list = []
    encode(seq, ebuf, avx2=avx2)
    return ebuf


def encode_np(seq, avx2: bool = False) -> np.ndarray:
    """Array form for large inputs: returns the packed words or raises."""
    a = _bytes_arr(seq)
    words = np.zeros(max(1, (a.size + 31) // 32), dtype=np.uint64)
    n, e = C.c_size_t(0), _Err()
    fn = lib().orc_encode_avx2 if avx2 else lib().orc_encode
    _raise(fn(_ptr(a), a.size, _ptr(words), C.byref(n), C.byref(e)), e)
    return words[: n.value]


def from_2bit(packed: int, expected_size: int, seq: bytearray, avx2: bool = False) -> None:
    """Appends to ``seq`` (never clears), as the reference does."""
    out = np.zeros(max(1, min(expected_size, 32)), dtype=np.uint8)
    e = _Err()
    fn = lib().orc_from_2bit_avx2 if avx2 else lib().orc_from_2bit
    _raise(fn(C.c_uint64(packed & (2**64 - 1)), expected_size, _ptr(out), C.byref(e)), e)
    seq.extend(out[:expected_size].tobytes())


def from_2bit_alloc(packed: int, expected_size: int, avx2: bool = False) -> bytearray:
    seq = bytearray()
    from_2bit(packed, expected_size, seq, avx2=avx2)
    return seq


def decode_np(ebuf, n_bases: int, path: int = PATH_AVX2) -> np.ndarray:
    w = _words_arr(ebuf)
    out = np.zeros(32 * max(1, w.size, (n_bases + 31) // 32), dtype=np.uint8)
    n, e = C.c_size_t(0), _Err()
    rc = lib().orc_decode(_ptr(w), w.size, n_bases, _ptr(out), C.byref(n), path, C.byref(e))
    res = out[: n.value]
    if rc:
        try:
            _raise(rc, e)
        except (OracleError, OraclePanic) as ex:
            ex.partial = res  # what the reference had appended before failing
            raise
    return res


def decode(ebuf, n_bases: int, dbuf: bytearray, path: int = PATH_AVX2) -> None:
    """``decode(&[u64], usize, &mut Vec<u8>)``: appends to ``dbuf``."""
    try:
        dbuf.extend(decode_np(ebuf, n_bases, path).tobytes())
    except (OracleError, OraclePanic) as ex:
        dbuf.extend(ex.partial.tobytes())
        raise


def hdist_scalar(u: int, v: int, length: int) -> int:
    out, e = C.c_uint32(0), _Err()
    _raise(lib().orc_hdist_scalar(C.c_uint64(u), C.c_uint64(v), length, C.byref(out), C.byref(e)), e)
    return out.value


def hdist(ebuf1, ebuf2, n_bases: int, path: int = PATH_AVX2, wide: bool = False) -> int:
    """u32 result with release-build wrap-around; ``wide=True`` returns the unwrapped u64."""
    a, b = _words_arr(ebuf1), _words_arr(ebuf2)
    out, tot, e = C.c_uint32(0), C.c_uint64(0), _Err()
    _raise(lib().orc_hdist(_ptr(a), a.size, _ptr(b), b.size, n_bases, C.byref(out), C.byref(tot),
                           path, C.byref(e)), e)
    return tot.value if wide else out.value


def base_counts(data, length: int) -> list:
    w = _words_arr(data)
    counts = (C.c_uint64 * 4)()
    lib().orc_base_counts(_ptr(w), length, counts)
    return list(counts)


def gc_content(data, length: int) -> float:
    w = _words_arr(data)
    return float(lib().orc_gc_content(_ptr(w), length))


def split_packed(ebuf, slen: int, idx: int):
    w = _words_arr(ebuf)
    lb = np.zeros(w.size + 1, dtype=np.uint64)
    rb = np.zeros(w.size + 1, dtype=np.uint64)
    nl, nr, e = C.c_size_t(0), C.c_size_t(0), _Err()
    _raise(lib().orc_split_packed(_ptr(w), w.size, slen, idx, _ptr(lb), C.byref(nl), _ptr(rb),
                                  C.byref(nr), C.byref(e)), e)
    return [int(x) for x in lb[: nl.value]], [int(x) for x in rb[: nr.value]]


class PackedSequence:
    """/root/reference/src/sequence.rs:5-9 restated over the oracle."""

    def __init__(self, seq):
        seq = bytes(seq)
        self.data = [] if len(seq) == 0 else encode_alloc(seq)  # sequence.rs:42-46
        self.length = len(seq)

    def __len__(self):
        return self.length

    def is_empty(self):
        return self.length == 0

    def get(self, index: int) -> int:
        w = _words_arr(self.data)
        out, e = C.c_uint8(0), _Err()
        _raise(lib().orc_seq_get(_ptr(w), self.length, index, C.byref(out), C.byref(e)), e)
        return out.value

    def slice(self, start: int, end: int) -> bytes:
        w = _words_arr(self.data)
        out = np.zeros(max(1, end - start if end > start else 1), dtype=np.uint8)
        e = _Err()
        _raise(lib().orc_seq_slice(_ptr(w), self.length, start, end, _ptr(out), C.byref(e)), e)
        return out[: end - start].tobytes()

    def to_vec(self) -> bytes:
        return self.slice(0, self.length)

    def base_counts(self):
        return base_counts(self.data, self.length)

    def gc_content(self):
        return gc_content(self.data, self.length)

    def __eq__(self, other):
        return (self.data, self.length) == (other.data, other.length)

    def __hash__(self):
        return hash((tuple(self.data), self.length))


# ------------------------------------------------------------------ synthetic input -----------

DEFAULT_SEED = 0x5EEDB17C0DE5


class FastqFault(Exception):
    """Malformed FASTQ: (record, fault) with fault 1 header, 2 separator, 3 quality length, 4 truncated."""

    def __init__(self, record: int, fault: int):
        super().__init__(f"FASTQ record {record}: fault {fault}")
        self.record, self.fault = record, fault


def fastq_scan(text, fasta: bool = False):
    """(starts, lens) of every sequence line of a FASTQ text (or a one-sequence-line FASTA text), or FastqFault."""
    t = _bytes_arr(text)
    cap = t.size // 2 + 1
    starts, lens = np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64)
    n_reads, bad, fault = C.c_size_t(0), C.c_uint64(0), C.c_int(0)
    scan = lib().orc_fasta_scan if fasta else lib().orc_fastq_scan
    rc = scan(_ptr(t), t.size, _ptr(starts), _ptr(lens), cap, C.byref(n_reads), C.byref(bad), C.byref(fault))
    if rc != 0:
        raise FastqFault(bad.value, fault.value)
    return starts[: n_reads.value].copy(), lens[: n_reads.value].copy()


def fasta_scan(text):
    return fastq_scan(text, fasta=True)


def fastq_encode(text, fasta: bool = False):
    """The caller's loop of README.md:160-180 over a FASTQ text: (words, word_offsets, starts, lens); every read is
    PackedSequence::new(record.seq()) on fresh words.  FastqFault first, then the first InvalidBase in file order."""
    t = _bytes_arr(text)
    starts, lens = fastq_scan(t, fasta)
    words, offs = [], [0]
    for s, l in zip(starts.tolist(), lens.tolist()):
        if l:
            words.append(encode_np(t[s : s + l]))
        offs.append(offs[-1] + (l + 31) // 32)
    w = np.concatenate(words) if words else np.zeros(0, dtype=np.uint64)
    return w, np.asarray(offs, dtype=np.uint64), starts, lens


def fasta_wrapped_encode(text):
    """Wrapped (multi-line) FASTA: the definition the CUDA path is held to (the reference has no parser; README.md:160-180
    shows the caller's loop over a FASTA reader, which joins a record's sequence lines before `PackedSequence::new`).
    Lines end in "\n" or "\r\n", the last newline may be missing; a line that opens with '>' starts a record; every other
    line is sequence of the current record (an empty line adds nothing).  A non-empty text whose first line is not a
    header -> FastqFault(0, 1).  Returns (words, word_offsets, header_offsets, seq_lens); the first byte outside ACGTacgt in
    file order raises OracleError(InvalidBase) carrying .record and .position (inside the record's joined sequence)."""
    t = _bytes_arr(text).tobytes()
    if not t:
        z = np.zeros(0, dtype=np.uint64)
        return z, np.zeros(1, dtype=np.uint64), z.copy(), z.copy()
    lines, pos = [], 0
    while pos < len(t):                       # (start, end without "\n" / "\r\n") of every line
        nl = t.find(b"\n", pos)
        end = len(t) if nl < 0 else nl
        stop = end - 1 if end > pos and t[end - 1:end] == b"\r" else end   # as orc_fastq_scan: a '\r' at the end of a line is stripped
        lines.append((pos, stop))
        pos = end + 1
    if t[lines[0][0]:lines[0][0] + 1] != b">":
        raise FastqFault(0, 1)
    hdr, seqs = [], []
    for s, e in lines:
        if t[s:s + 1] == b">":
            hdr.append(s)
            seqs.append([])
        else:
            seqs[-1].append(t[s:e])
    words, offs, lens = [], [0], []
    for r, parts in enumerate(seqs):
        seq = b"".join(parts)
        lens.append(len(seq))
        if seq:
            try:
                words.append(encode_np(np.frombuffer(seq, dtype=np.uint8)))
            except OracleError as e:
                bad = next(i for i, b in enumerate(seq) if b not in b"ACGTacgt")
                e.record, e.position = r, bad
                raise
        offs.append(offs[-1] + (len(seq) + 31) // 32)
    w = np.concatenate(words) if words else np.zeros(0, dtype=np.uint64)
    return w, np.asarray(offs, dtype=np.uint64), np.asarray(hdr, dtype=np.uint64), np.asarray(lens, dtype=np.uint64)


def fastx_encode_timed(text, fasta: bool = False, path: int = PATH_AVX2, reps: int = 3):
    """(seconds, words, word_offsets) of the single-threaded CPU form of the FASTQ / FASTA row: reader + per-record encode
    in C (orc_fastx_encode), best of ``reps``.  bench tools only."""
    import time
    t = _bytes_arr(text)
    reads_cap = t.size // (2 if fasta else 4) + 1
    words = np.empty(t.size // 32 + reads_cap, dtype=np.uint64)
    wo = np.empty(reads_cap + 1, dtype=np.uint64)
    n_reads, bad, fault, e = C.c_size_t(0), C.c_uint64(0), C.c_int(0), _Err()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = lib().orc_fastx_encode(_ptr(t), t.size, int(fasta), path, _ptr(words), words.size, _ptr(wo), reads_cap, C.byref(n_reads),
                                    C.byref(bad), C.byref(fault), C.byref(e))
        best = min(best, time.perf_counter() - t0)
        if rc == -5:
            raise FastqFault(bad.value, fault.value)
        _raise(rc, e)
    n = n_reads.value
    return best, words[: int(wo[n])], wo[: n + 1]


def synth_word(seed: int, stream: int, j: int) -> int:
    return int(lib().orc_synth_word(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(j)))


def synth_ascii(seed: int, stream: int, first_base: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_synth_ascii(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(first_base), n, _ptr(out))
    return out


class CodecBench:
    """Timed CPU baseline: buffers are allocated and touched once, then reused for every repetition
    (so page faults of fresh output buffers are not billed to the codec)."""

    def __init__(self, seq: np.ndarray, threads: int):
        self.seq = _bytes_arr(seq)
        self.threads = threads
        self.ebuf = np.ones((self.seq.size + 31) // 32 + threads, dtype=np.uint64)
        self.dbuf = np.ones(self.seq.size + 32, dtype=np.uint8)

    def run(self, reps: int = 1, path: int = PATH_AVX2, do_encode: bool = True, do_decode: bool = True) -> float:
        """Best-of-``reps`` wall seconds for encode(+decode) over ``threads`` pthreads."""
        a = self.seq
        if not do_encode:
            self.ebuf[: (a.size + 31) // 32] = encode_np(a)
        return float(lib().orc_bench_codec(_ptr(a), a.size, self.threads, reps, path, int(do_encode),
                                           int(do_decode), _ptr(self.ebuf), _ptr(self.dbuf)))


def bench_codec(seq: np.ndarray, threads: int, reps: int, path: int = PATH_AVX2,
                do_encode: bool = True, do_decode: bool = True) -> float:
    return CodecBench(seq, threads).run(reps, path, do_encode, do_decode)


def bench_op(op: int, n_units: int, *, in0: np.ndarray, in1: np.ndarray | None = None, out0: np.ndarray | None = None,
             out1: np.ndarray | None = None, k: int = 0, stride: int = 0, path: int = PATH_AVX2, threads: int = 1,
             reps: int = 5, pin: bool = True):
    """Times one op of the reference's CPU path (orc_bench_op): returns (seconds per repetition, checksum).
    Buffers belong to the caller and are reused by every repetition (touch outputs before timing)."""
    d = _BenchDesc(op, path, n_units, k, stride, _ptr(in0), _ptr(in1) if in1 is not None else None,
                   _ptr(out0) if out0 is not None else None, _ptr(out1) if out1 is not None else None)
    times = (C.c_double * reps)()
    check = C.c_uint64(0)
    rc = lib().orc_bench_op(C.byref(d), threads, reps, int(pin), times, C.byref(check))
    if rc:
        raise RuntimeError(f"orc_bench_op({op}) failed: {rc}")
    return list(times), int(check.value)
