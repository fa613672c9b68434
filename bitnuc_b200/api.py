"""Host-side mirror of the reference's public API over the C ABI (host-pointer entry points).

Same names, argument meaning and error behaviour as the functions re-exported at
/root/reference/src/lib.rs:214-220.  ``Vec<u64>`` arguments are Python lists of ints that are
cleared/filled, ``Vec<u8>`` arguments are ``bytearray``s that are appended to -- exactly what the
reference does to them -- so the parity tests read like the reference's own tests.  The ``*_np``
and ``*_batch`` functions are the array forms for real workloads.

Every function runs CUDA kernels through libbitnuc_cuda.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np

from . import _lib
from ._lib import BnError, raise_for
from .errors import NucleotideError, ReferencePanic

M64 = (1 << 64) - 1


class Context:
    """A bn_ctx bound to one CUDA device (streams, staging buffers, status words)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        raise_for(self.lib.bn_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            for ref in getattr(self, "_pinned", []):  # page-locked buffers handed out by pinned_empty
                owner = ref()
                if owner is not None:
                    owner.release()
            self._pinned = []
            self.lib.bn_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        raise_for(self.lib.bn_ctx_synchronize(self.handle))

    def set_chunk_bytes(self, n: int):
        raise_for(self.lib.bn_ctx_set_chunk_bytes(self.handle, n))

    def set_compat(self, isa: str):
        """Which of the reference's per-ISA paths is mirrored where they disagree: ``"x86_64"`` (default; packing/avx.rs,
        unpacking/avx.rs) or ``"aarch64"`` (packing/aarch64.rs:173-244, unpacking/aarch64.rs:127-130) -- SURVEY.md 8f-4."""
        raise_for(self.lib.bn_ctx_set_compat(self.handle, {"x86_64": 0, "aarch64": 1}[isa]))

    @property
    def compat(self) -> str:
        return ("x86_64", "aarch64")[self.lib.bn_ctx_compat(self.handle)]

    def pinned_empty(self, n: int, dtype=np.uint8) -> np.ndarray:
        """A page-locked numpy array; host-pointer calls on it run at the PCIe link rate and overlap.  The allocation
        is owned by the array (its ``base`` chain holds a ``_Pinned`` that calls ``bn_host_free`` when the last view
        goes away) and is released at the latest by ``Context.close()`` -- do not use the array after that."""
        dtype = np.dtype(dtype)
        owner = _Pinned(self, max(1, n * dtype.itemsize))
        arr = np.frombuffer(owner, dtype=dtype, count=n)   # arr.base -> memoryview -> owner (buffer protocol)
        self._pinned = [r for r in getattr(self, "_pinned", []) if r() is not None]
        self._pinned.append(weakref.ref(owner))
        return arr


class _Pinned:
    """Owner of one bn_host_alloc allocation, exported through the buffer protocol (PEP 688)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self._lib, self._ctx_handle, self.nbytes = ctx.lib, ctx.handle, nbytes
        p = C.c_void_p()
        raise_for(ctx.lib.bn_host_alloc(ctx.handle, nbytes, C.byref(p)))
        self.ptr = p
        self._view = (C.c_uint8 * nbytes).from_address(p.value)

    def __buffer__(self, flags):
        if self.ptr is None:
            raise BufferError("pinned buffer already released")
        return memoryview(self._view)

    def release(self):
        if self.ptr is not None:
            self._lib.bn_host_free(self._ctx_handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


_default = {}
_default_lock = threading.Lock()


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = _current_device()
    with _default_lock:
        ctx = _default.get(device)
        if ctx is None:
            ctx = _default[device] = Context(device)
        return ctx


def _current_device() -> int:
    import os
    return int(os.environ.get("BITNUC_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def _u8(seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    return np.frombuffer(bytes(seq), dtype=np.uint8)


def _u64(words) -> np.ndarray:
    if isinstance(words, np.ndarray) and words.dtype == np.uint64:
        return np.ascontiguousarray(words)
    return np.array([int(w) & M64 for w in words], dtype=np.uint64)


def _check_offsets(off: np.ndarray, n_bytes: int) -> None:
    """Read offsets must never decrease and must end inside the byte buffer (checked before any arithmetic on them)."""
    if off.size and (bool(np.any(off[1:] < off[:-1])) or int(off[-1]) > n_bytes):
        raise ValueError("offsets must be non-decreasing and offsets[-1] <= len(data)")


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


# ------------------------------------------------------------------ encode / decode ----------

def encode_np(seq, ctx: Context | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """Packed words of ``seq`` (array form of ``encode_alloc``)."""
    words, n, rc, err = _encode_raw(seq, ctx, out)
    raise_for(rc, err)
    return words[:n]


def _encode_raw(seq, ctx, out=None):
    ctx = ctx or default_context()
    a = _u8(seq)
    need = (a.size + 31) // 32
    words = out if out is not None else np.empty(max(1, need), dtype=np.uint64)
    n, err = C.c_size_t(0), BnError()
    rc = ctx.lib.bn_encode(ctx.handle, _p(a), a.size, _p(words), C.byref(n), C.byref(err))
    return words, n.value, rc, err


def encode(sequence, ebuf: list, ctx: Context | None = None) -> None:
    """``encode(&[u8], &mut Vec<u64>)`` (src/utils/mod.rs:22-25): clears ``ebuf``, then fills it.
    On ``InvalidBase`` ``ebuf`` keeps the words of the chunks before the failing chunk
    (src/utils/packing/avx.rs:132,142-143)."""
    ctx = ctx or default_context()
    words, n, rc, err = _encode_raw(sequence, ctx)
    if ctx.compat == "aarch64":
        # packing/aarch64.rs:222-244: under 32 bases one word is PUSHED (nothing is cleared; nothing on error); from
        # 32 bases on the Vec is resized to ceil(n/32) words, zero-filled and overwritten up to the failing block
        size = len(_u8(sequence))
        if size < 32:
            if rc == 0:
                ebuf.append(int(words[0]))
        else:
            ebuf[:] = [int(w) for w in words[:n]] + [0] * ((size + 31) // 32 - n)
        raise_for(rc, err)
        return
    if rc != _lib.BN_ERR_EMPTY_ENCODE:  # the reference panics before it clears anything useful
        ebuf.clear()
        ebuf.extend(int(w) for w in words[:n])
    raise_for(rc, err)


def encode_alloc(sequence, ctx: Context | None = None) -> list:
    """``encode_alloc(&[u8]) -> Vec<u64>`` (src/utils/mod.rs:38-42)."""
    ebuf: list = []
    encode(sequence, ebuf, ctx)
    return ebuf


def decode_np(ebuf, n_bases: int, ctx: Context | None = None, out: np.ndarray | None = None) -> np.ndarray:
    ctx = ctx or default_context()
    w = _u64(ebuf)
    res = out if out is not None else np.empty(max(1, n_bases), dtype=np.uint8)
    err = BnError()
    raise_for(ctx.lib.bn_decode(ctx.handle, _p(w), w.size, n_bases, _p(res), C.byref(err)), err)
    return res[:n_bases]


def decode(ebuf, n_bases: int, dbuf: bytearray, ctx: Context | None = None) -> None:
    """``decode(&[u64], usize, &mut Vec<u8>)`` (src/utils/mod.rs:60-62): appends to ``dbuf`` -- or, in aarch64 mode,
    resizes it to ``n_bases`` and overwrites it (src/utils/unpacking/aarch64.rs:127-130), words missing from ``ebuf``
    reading as zero in the whole 32-base chunks (``input.get(i).copied().unwrap_or(0)``, :113) and panicking in the tail."""
    ctx = ctx or default_context()
    if ctx.compat == "aarch64":
        w = _u64(ebuf)
        need = (n_bases + 31) // 32
        if w.size < need:
            if n_bases % 32 and w.size < need:      # the scalar tail indexes input[j / 32] (:121)
                raise ReferencePanic("index out of bounds: the reference panics (unpacking/aarch64.rs:121)")
            w = np.concatenate([w, np.zeros(need - w.size, dtype=np.uint64)])
        dbuf[:] = decode_np(w, n_bases, ctx).tobytes()
        return
    dbuf.extend(decode_np(ebuf, n_bases, ctx).tobytes())


# ------------------------------------------------------------------ k-mers -------------------

def as_2bit_batch(recs, n: int, k: int, stride: int | None = None, ctx: Context | None = None) -> np.ndarray:
    """``as_2bit`` over ``n`` records of ``k`` bases laid out every ``stride`` bytes."""
    ctx = ctx or default_context()
    stride = k if stride is None else stride
    a = _u8(recs)
    if n and k <= 32 and a.size < (n - 1) * stride + k:
        raise ValueError("record buffer too small")
    out = np.empty(max(1, n), dtype=np.uint64)
    err = BnError()
    rc = ctx.lib.bn_as_2bit_batch(ctx.handle, _p(a), n, k, stride, _p(out), C.byref(err))
    if rc == 1:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.offset = int(err.record), int(err.offset)
        raise e
    raise_for(rc, err)
    return out[:n]


def as_2bit(seq, ctx: Context | None = None) -> int:
    """``as_2bit(&[u8]) -> Result<u64>`` (src/utils/packing/mod.rs:81-110)."""
    a = _u8(seq)
    return int(as_2bit_batch(a, 1, a.size, max(1, a.size), ctx)[0])


def from_2bit_batch(packed, k: int, stride: int | None = None, ctx: Context | None = None,
                    out: np.ndarray | None = None) -> np.ndarray:
    """``from_2bit`` over an array of words; returns the ``(n-1)*stride + k`` output bytes."""
    ctx = ctx or default_context()
    stride = k if stride is None else stride
    w = _u64(packed)
    n = w.size
    nbytes = (n - 1) * stride + k if n and k <= 32 else 0
    res = out if out is not None else np.zeros(max(1, nbytes), dtype=np.uint8)
    err = BnError()
    raise_for(ctx.lib.bn_from_2bit_batch(ctx.handle, _p(w), n, k, _p(res), stride, C.byref(err)), err)
    return res[:nbytes]


def from_2bit(packed: int, expected_size: int, sequence: bytearray, ctx: Context | None = None) -> None:
    """``from_2bit(u64, usize, &mut Vec<u8>)`` (src/utils/unpacking/mod.rs:119-147): appends."""
    w = np.array([int(packed) & M64], dtype=np.uint64)
    sequence.extend(from_2bit_batch(w, expected_size, max(1, expected_size), ctx).tobytes())


def from_2bit_alloc(packed: int, expected_size: int, ctx: Context | None = None) -> bytearray:
    """``from_2bit_alloc(u64, usize) -> Vec<u8>`` (src/utils/unpacking/mod.rs:178-182)."""
    seq = bytearray()
    from_2bit(packed, expected_size, seq, ctx)
    return seq


# ------------------------------------------------------------------ hamming ------------------

def hdist_total(ebuf1, ebuf2, n_bases: int, ctx: Context | None = None) -> int:
    """Exact (u64) mismatch count of two packed sequences."""
    ctx = ctx or default_context()
    a, b = _u64(ebuf1), _u64(ebuf2)
    total, err = C.c_uint64(0), BnError()
    raise_for(ctx.lib.bn_hdist(ctx.handle, _p(a), a.size, _p(b), b.size, n_bases, C.byref(total), C.byref(err)), err)
    return int(total.value)


def hdist(ebuf1, ebuf2, n_bases: int, ctx: Context | None = None) -> int:
    """``hdist(&[u64], &[u64], usize) -> Result<u32>`` (src/utils/functions/hamming/multi.rs:122-160).
    The reference accumulates in a ``u32`` (:130) that wraps in release builds; so does this."""
    return hdist_total(ebuf1, ebuf2, n_bases, ctx) & 0xFFFFFFFF


def hdist_pairs(u, v, length: int, ctx: Context | None = None) -> np.ndarray:
    """``hdist_scalar`` over arrays of words: one u32 per pair."""
    ctx = ctx or default_context()
    a, b = _u64(u), _u64(v)
    if a.size != b.size:
        raise ValueError("u and v differ in length")
    out = np.empty(max(1, a.size), dtype=np.uint32)
    err = BnError()
    raise_for(ctx.lib.bn_hdist_pairs(ctx.handle, _p(a), _p(b), a.size, length, _p(out), C.byref(err)), err)
    return out[: a.size]


def hdist_scalar(u: int, v: int, length: int, ctx: Context | None = None) -> int:
    """``hdist_scalar(u64, u64, usize) -> Result<u32>`` (src/utils/functions/hamming/scalar.rs:11-48)."""
    return int(hdist_pairs(np.array([u & M64], dtype=np.uint64), np.array([v & M64], dtype=np.uint64), length, ctx)[0])


# ------------------------------------------------------------------ analysis -----------------

def base_counts_gc(words, n_bases: int, ctx: Context | None = None):
    """([A,C,G,T], gc%) of one packed sequence (src/utils/analysis.rs:3-39)."""
    ctx = ctx or default_context()
    w = _u64(words)
    counts, gc, err = (C.c_uint64 * 4)(), C.c_double(0.0), BnError()
    raise_for(ctx.lib.bn_base_counts(ctx.handle, _p(w), w.size, n_bases, counts, C.byref(gc), C.byref(err)), err)
    return [int(x) for x in counts], float(gc.value)


def base_counts_batch(words, word_offsets, lens, ctx: Context | None = None):
    """Per-read ``base_counts`` / ``gc_content`` for a batch of packed reads.
    Returns (counts[n,4] u64, gc[n] f64, totals[4])."""
    ctx = ctx or default_context()
    w, wo, ln = _u64(words), _u64(word_offsets), _u64(lens)
    n = ln.size
    counts = np.empty((max(1, n), 4), dtype=np.uint64)
    gc = np.empty(max(1, n), dtype=np.float64)
    totals, err = (C.c_uint64 * 4)(), BnError()
    raise_for(ctx.lib.bn_base_counts_batch(ctx.handle, _p(w), w.size, _p(wo), _p(ln), n, _p(counts), _p(gc), totals,
                                           C.byref(err)), err)
    return counts[:n], gc[:n], [int(x) for x in totals]


def encode_batch(data, offsets, ctx: Context | None = None, per_read_status: bool = False):
    """``PackedSequence::new`` over a batch of reads ``data[offsets[r]:offsets[r+1]]``.
    Returns (words, word_offsets[, read_status]); raises ``InvalidBase`` (with ``.record``,
    ``.position`` and ``.offset`` attributes) for the first invalid base in input order unless
    ``per_read_status`` is set, in which case the per-read first-invalid positions are returned."""
    ctx = ctx or default_context()
    a, off = _u8(data), _u64(offsets)
    n = off.size - 1
    if n < 0:
        raise ValueError("offsets needs n_reads + 1 entries")
    _check_offsets(off, a.size)   # the C ABI takes no byte-buffer length: a bad offsets array would read past `data`
    max_words = int((off[-1] - off[0]) // np.uint64(32)) + n if n else 0
    words = np.empty(max(1, max_words), dtype=np.uint64)
    wo = np.zeros(n + 1, dtype=np.uint64)
    status = np.empty(max(1, n), dtype=np.uint32) if per_read_status else None
    err = BnError()
    rc = ctx.lib.bn_encode_batch(ctx.handle, _p(a), _p(off), n, _p(words), _p(wo), _p(status) if status is not None else None,
                                 C.byref(err))
    if rc == 1 and not per_read_status:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
        raise e
    if rc != 1:
        raise_for(rc, err)
    words = words[: int(wo[n])]
    return (words, wo, status[:n]) if per_read_status else (words, wo)


def fasta_encode(text, ctx: Context | None = None):
    """FASTA text with one sequence line per record -> (words, word_offsets, seq_offsets, seq_lens); see ``fastq_encode``."""
    return fastq_encode(text, ctx, _fasta=True)


def fastq_encode(text, ctx: Context | None = None, _fasta: bool = False):
    """FASTQ text -> (words, word_offsets, seq_offsets, seq_lens), parsed and encoded on the device: the caller's loop
    ``for record in reader { PackedSequence::new(record.seq())? }`` (README.md:160-180, src/sequence.rs:40-52) without a
    host-side parser.  Raises ``FastqError`` for the first malformed record, else ``InvalidBase`` (with ``.record``,
    ``.position``, ``.offset``) for the first byte outside ACGTacgt in file order."""
    ctx = ctx or default_context()
    t = _u8(text)
    nr, nw, err = C.c_size_t(0), C.c_size_t(0), BnError()
    scan, enc = (ctx.lib.bn_fasta_scan, ctx.lib.bn_fasta_encode) if _fasta else (ctx.lib.bn_fastq_scan, ctx.lib.bn_fastq_encode)
    raise_for(scan(ctx.handle, _p(t), t.size, C.byref(nr), C.byref(nw), C.byref(err)), err)
    n, w = nr.value, nw.value
    words = np.empty(max(1, w), dtype=np.uint64)
    wo = np.zeros(n + 1, dtype=np.uint64)
    so, sl = np.empty(max(1, n), dtype=np.uint64), np.empty(max(1, n), dtype=np.uint64)
    rc = enc(ctx.handle, _p(t), t.size, n, w, _p(words), _p(wo), _p(so), _p(sl), C.byref(err))
    if rc == 1:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
        raise e
    raise_for(rc, err)
    return words[:w], wo, so[:n], sl[:n]


def fasta_wrapped_encode(text, ctx: Context | None = None):
    """Wrapped (multi-line) FASTA text -> (words, word_offsets, header_offsets, seq_lens): a ``>`` header line, then any
    number of sequence lines per record, whose concatenation is the record's sequence -- what a FASTA reader hands to
    ``PackedSequence::new(record.seq())`` (README.md:160-180).  Raises ``FastqError`` when the text does not open with a
    header, else ``InvalidBase`` (``.record``, ``.position`` inside the record's sequence) for the first bad byte."""
    ctx = ctx or default_context()
    t = _u8(text)
    nr, nb, nw, err = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), BnError()
    rc = ctx.lib.bn_fasta_wrapped_scan(ctx.handle, _p(t), t.size, C.byref(nr), C.byref(nb), C.byref(nw), C.byref(err))
    if rc == 1:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
        raise e
    raise_for(rc, err)
    n, w = nr.value, nw.value
    words = np.empty(max(1, w), dtype=np.uint64)
    wo = np.zeros(n + 1, dtype=np.uint64)
    ho, sl = np.empty(max(1, n), dtype=np.uint64), np.empty(max(1, n), dtype=np.uint64)
    raise_for(ctx.lib.bn_fasta_wrapped_encode(ctx.handle, _p(t), t.size, n, w, _p(words), _p(wo), _p(ho), _p(sl), C.byref(err)), err)
    return words[:w], wo, ho[:n], sl[:n]


# ------------------------------------------------------------------ split_packed -------------

def split_packed_batch(words, word_offsets, lens, idx, ctx: Context | None = None):
    """``split_packed`` (src/utils/functions/split.rs:14-102) over a batch of packed reads: read ``r`` is
    ``words[word_offsets[r]:word_offsets[r+1]]`` holding ``lens[r]`` bases, split at base ``idx[r]``.
    Returns (left, left_offsets, right, right_offsets); raises ``IndexOutOfBounds`` (with ``.record``)
    for the first read with ``idx > len``."""
    ctx = ctx or default_context()
    w, wo, ln, ix = _u64(words), _u64(word_offsets), _u64(lens), _u64(idx)
    n = ln.size
    if wo.size != n + 1 or ix.size != n:
        raise ValueError("word_offsets needs n_reads + 1 entries, idx n_reads")
    left = np.empty(max(1, w.size + n), dtype=np.uint64)
    right = np.empty(max(1, w.size), dtype=np.uint64)
    lo = np.zeros(n + 1, dtype=np.uint64)
    ro = np.zeros(n + 1, dtype=np.uint64)
    err = BnError()
    rc = ctx.lib.bn_split_packed_batch(ctx.handle, _p(w), w.size, _p(wo), _p(ln), _p(ix), n, _p(left), _p(lo), _p(right), _p(ro),
                                       C.byref(err))
    if rc in (3, 4):
        e = NucleotideError.IndexOutOfBounds(err.a, err.b) if rc == 4 else NucleotideError.InvalidLength(err.a)
        e.record = int(err.record)
        raise e
    raise_for(rc, err)
    return left[: int(lo[n])], lo, right[: int(ro[n])], ro


def split_packed(ebuf, slen: int, idx: int, lbuf: list, rbuf: list, ctx: Context | None = None) -> None:
    """``split_packed(&[u64], usize, usize, &mut Vec<u64>, &mut Vec<u64>)`` (src/utils/functions/split.rs:14-20):
    validates ``idx <= slen`` first, then clears both buffers and fills them."""
    w = _u64(ebuf)
    left, _, right, _ = split_packed_batch(w, np.array([0, w.size], dtype=np.uint64), np.array([slen], dtype=np.uint64),
                                           np.array([idx], dtype=np.uint64), ctx)
    lbuf.clear()
    rbuf.clear()
    lbuf.extend(int(x) for x in left)
    rbuf.extend(int(x) for x in right)


# ------------------------------------------------------------------ get / slice --------------

def slice_batch(words, word_offsets, lens, q_read, q_start, q_end, ctx: Context | None = None):
    """``PackedSequence::slice`` (src/sequence.rs:198-212) over a batch of queries: query ``q`` = bases
    ``[q_start[q], q_end[q])`` of read ``q_read[q]``.  Returns (bytes, out_offsets); raises ``InvalidRange``
    (with ``.record``) for the first query with ``start > end`` or ``end > len``."""
    ctx = ctx or default_context()
    w, wo, ln = _u64(words), _u64(word_offsets), _u64(lens)
    qr, qs, qe = _u64(q_read), _u64(q_start), _u64(q_end)
    nq = qr.size
    if qs.size != nq or qe.size != nq:
        raise ValueError("q_read, q_start and q_end differ in length")
    cap = int(np.maximum(qe.astype(np.int64) - qs.astype(np.int64), 0).sum()) if nq else 0
    out = np.empty(max(1, cap), dtype=np.uint8)
    oo = np.zeros(nq + 1, dtype=np.uint64)
    err = BnError()
    rc = ctx.lib.bn_slice_batch(ctx.handle, _p(w), w.size, _p(wo), _p(ln), ln.size, _p(qr), _p(qs), _p(qe), nq, _p(out), cap, _p(oo),
                                C.byref(err))
    if rc == 5:
        e = NucleotideError.InvalidRange(err.a, err.b, err.c)
        e.record = int(err.record)
        raise e
    raise_for(rc, err)
    return out[: int(oo[nq])], oo


def get_batch(words, word_offsets, lens, q_read, q_index, ctx: Context | None = None) -> np.ndarray:
    """``PackedSequence::get`` (src/sequence.rs:116-135) over a batch of queries: one ASCII byte per query; raises
    ``IndexOutOfBounds`` (with ``.record``) for the first query with ``index >= len``."""
    ctx = ctx or default_context()
    w, wo, ln = _u64(words), _u64(word_offsets), _u64(lens)
    qr, qi = _u64(q_read), _u64(q_index)
    nq = qr.size
    if qi.size != nq:
        raise ValueError("q_read and q_index differ in length")
    out = np.empty(max(1, nq), dtype=np.uint8)
    err = BnError()
    rc = ctx.lib.bn_get_batch(ctx.handle, _p(w), w.size, _p(wo), _p(ln), ln.size, _p(qr), _p(qi), nq, _p(out), C.byref(err))
    if rc == 4:
        e = NucleotideError.IndexOutOfBounds(err.a, err.b)
        e.record = int(err.record)
        raise e
    raise_for(rc, err)
    return out[:nq]


# ------------------------------------------------------------------ k-mer windows ------------

def kmers(seq, k: int, ctx: Context | None = None) -> np.ndarray:
    """``[as_2bit(w)? for w in seq.windows(k)]`` (README.md:160-180): one packed word per window.  Raises the error
    of the first failing window (``.record`` = its index, ``.offset`` = the offending byte, ``.partial`` = the words of
    the windows before it)."""
    ctx = ctx or default_context()
    a = _u8(seq)
    if k <= 0:
        raise ValueError("window size must be non-zero (slice::windows panics)")
    out = np.empty(max(1, a.size - k + 1), dtype=np.uint64)
    n_out, err = C.c_size_t(0), BnError()
    rc = ctx.lib.bn_kmers(ctx.handle, _p(a), a.size, k, _p(out), C.byref(n_out), C.byref(err))
    if rc == 1:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.offset, e.partial = int(err.record), int(err.offset), out[: n_out.value]
        raise e
    raise_for(rc, err)
    return out[: n_out.value]


def kmers_batch(data, offsets, k: int, ctx: Context | None = None):
    """``kmers`` per read of a batch ``data[offsets[r]:offsets[r+1]]``: (words, out_offsets).  Windows never cross a
    read; a read shorter than ``k`` has none and is never validated.  Raises the error of the first failing window in
    (read, position) order, with ``.record`` / ``.position`` / ``.offset``."""
    ctx = ctx or default_context()
    a, off = _u8(data), _u64(offsets)
    n = off.size - 1
    if n < 0 or k <= 0:
        raise ValueError("offsets needs n_reads + 1 entries; window size must be non-zero")
    _check_offsets(off, a.size)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    cap = int(np.maximum(lens - k + 1, 0).sum()) if n else 0
    out = np.empty(max(1, cap), dtype=np.uint64)
    oo = np.zeros(n + 1, dtype=np.uint64)
    err = BnError()
    rc = ctx.lib.bn_kmers_batch(ctx.handle, _p(a), _p(off), n, k, _p(out), cap, _p(oo), C.byref(err))
    if rc == 1:
        e = NucleotideError.InvalidBase(err.base)
        e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
        raise e
    raise_for(rc, err)
    return out[:cap], oo
