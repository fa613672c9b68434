"""Device-resident variants: the same operations on torch CUDA tensors, enqueued on the current
torch stream through the ``*_dev`` entry points of the C ABI.  torch is plumbing here (device
memory, streams, torch.distributed); every kernel is ours.

ASCII buffers are ``torch.uint8`` tensors, packed buffers ``torch.int64`` tensors (bit-identical to
the reference's ``u64`` words).  Nothing synchronises unless stated: validation results live in a
device status word (``Status``) that is read back lazily.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, api
from ._lib import BnError, raise_for


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


_CUDA_STREAM_LEGACY = 0x1  # cudaStreamLegacy: the C ABI reads NULL as "the context's own stream"


def _stream():
    """torch's current stream as a cudaStream_t, so torch events and tensors order with our kernels."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream or _CUDA_STREAM_LEGACY)


def _ctx_for(t: torch.Tensor) -> api.Context:
    if not t.is_cuda:
        raise _lib.BitnucCudaError("device-resident calls need CUDA tensors (there is no CPU fallback)")
    return api.default_context(t.device.index)


def _want(t: torch.Tensor, dtype, numel: int, what: str) -> None:
    """The *_dev entry points take raw pointers: refuse tensors whose dtype, layout or size would make a kernel read or
    write outside them."""
    if t.dtype != dtype or not t.is_contiguous() or t.numel() < numel:
        raise ValueError(f"{what}: need a contiguous {dtype} tensor of at least {numel} elements, got {t.dtype}[{t.numel()}]"
                         f"{'' if t.is_contiguous() else ' (not contiguous)'}")


class Status:
    """Device-side validation status: min over invalid bases of (offset << 8 | byte)."""

    def __init__(self, device):
        self.word = torch.empty(1, dtype=torch.int64, device=device)
        self.ctx = api.default_context(self.word.device.index)   # a bare "cuda" resolves to the current device

    def check(self):
        """Synchronises the current stream; raises ``NucleotideError.InvalidBase`` if set."""
        err = BnError()
        rc = self.ctx.lib.bn_status_fetch(self.ctx.handle, _stream(), _ptr(self.word), C.byref(err))
        if rc == 1:
            e = _lib.NucleotideError.InvalidBase(err.base)
            e.offset = int(err.offset)
            raise e
        raise_for(rc, err)


def words_for(n_bases: int) -> int:
    return (n_bases + 31) // 32


def encode(seq: torch.Tensor, out: torch.Tensor | None = None, status: Status | None = None):
    """ASCII ``uint8[n]`` -> packed ``int64[ceil(n/32)]``.  Returns (words, status)."""
    ctx = _ctx_for(seq)
    n = seq.numel()
    if out is None:
        out = torch.empty(words_for(n), dtype=torch.int64, device=seq.device)
    status = status or Status(seq.device)
    raise_for(ctx.lib.bn_encode_dev(ctx.handle, _stream(), _ptr(seq), n, _ptr(out), _ptr(status.word)))
    return out, status


def decode(words: torch.Tensor, n_bases: int, out: torch.Tensor | None = None) -> torch.Tensor:
    ctx = _ctx_for(words)
    if out is None:
        out = torch.empty(n_bases, dtype=torch.uint8, device=words.device)
    rc = ctx.lib.bn_decode_dev(ctx.handle, _stream(), _ptr(words), words.numel(), n_bases, _ptr(out))
    if rc == 3:
        raise _lib.NucleotideError.InvalidLength(n_bases)
    raise_for(rc)
    return out


def as_2bit_batch(recs: torch.Tensor, n: int, k: int, stride: int | None = None, out=None, status=None):
    ctx = _ctx_for(recs)
    stride = k if stride is None else stride
    if n < 0 or stride < k:
        raise ValueError("n >= 0 and stride >= k")
    if 0 < k <= 32 and n:
        _want(recs, torch.uint8, (n - 1) * stride + k, "as_2bit_batch(recs)")
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=recs.device)
    _want(out, torch.int64, n, "as_2bit_batch(out)")
    status = status or Status(recs.device)
    rc = ctx.lib.bn_as_2bit_batch_dev(ctx.handle, _stream(), _ptr(recs), n, k, stride, _ptr(out), _ptr(status.word))
    if rc == 2:
        raise _lib.NucleotideError.SequenceTooLong(k)
    raise_for(rc)
    return out, status


def from_2bit_batch(packed: torch.Tensor, k: int, stride: int | None = None, out=None) -> torch.Tensor:
    ctx = _ctx_for(packed)
    stride = k if stride is None else stride
    n = packed.numel()
    if stride < k:
        raise ValueError("stride >= k")
    _want(packed, torch.int64, n, "from_2bit_batch(packed)")
    if out is None:
        out = torch.zeros((n - 1) * stride + k if n and k <= 32 else 0, dtype=torch.uint8, device=packed.device)
    elif n and 0 < k <= 32:
        _want(out, torch.uint8, (n - 1) * stride + k, "from_2bit_batch(out)")
    rc = ctx.lib.bn_from_2bit_batch_dev(ctx.handle, _stream(), _ptr(packed), n, k, _ptr(out), stride)
    if rc == 3:
        raise _lib.NucleotideError.InvalidLength(k)
    raise_for(rc)
    return out


def hdist(a: torch.Tensor, b: torch.Tensor, n_bases: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Exact mismatch count as a device ``int64[1]`` (``.item() & 0xFFFFFFFF`` is the reference's u32)."""
    ctx = _ctx_for(a)
    need = words_for(n_bases)
    if a.numel() < need or b.numel() < need:
        raise _lib.NucleotideError.InvalidLength(n_bases)
    if out is None:
        out = torch.empty(1, dtype=torch.int64, device=a.device)
    raise_for(ctx.lib.bn_hdist_dev(ctx.handle, _stream(), _ptr(a), _ptr(b), n_bases, _ptr(out)))
    return out


def hdist_pairs(u: torch.Tensor, v: torch.Tensor, length: int, out: torch.Tensor | None = None) -> torch.Tensor:
    ctx = _ctx_for(u)
    if out is None:
        out = torch.empty(u.numel(), dtype=torch.int32, device=u.device)
    rc = ctx.lib.bn_hdist_pairs_dev(ctx.handle, _stream(), _ptr(u), _ptr(v), u.numel(), length, _ptr(out))
    if rc == 3:
        raise _lib.NucleotideError.InvalidLength(length)
    raise_for(rc)
    return out


def base_counts(words: torch.Tensor, n_bases: int, counts=None, gc=None):
    """(counts int64[4] = [A,C,G,T], gc float64[1]) on the device."""
    ctx = _ctx_for(words)
    if words.numel() < words_for(n_bases):
        raise _lib.NucleotideError.InvalidLength(n_bases)
    counts = counts if counts is not None else torch.empty(4, dtype=torch.int64, device=words.device)
    gc = gc if gc is not None else torch.empty(1, dtype=torch.float64, device=words.device)
    raise_for(ctx.lib.bn_base_counts_dev(ctx.handle, _stream(), _ptr(words), n_bases, _ptr(counts), _ptr(gc)))
    return counts, gc


def base_counts_batch(words, word_offsets, lens, counts4=None, gc=None, totals=None, want_counts=True, want_gc=True):
    """Per-read counts ``int64[n,4]``, gc ``float64[n]`` and totals ``int64[4]`` on the device."""
    ctx = _ctx_for(words)
    n = lens.numel()
    dev = words.device
    if counts4 is None and want_counts:
        counts4 = torch.empty((n, 4), dtype=torch.int64, device=dev)
    if gc is None and want_gc:
        gc = torch.empty(n, dtype=torch.float64, device=dev)
    if totals is None:
        totals = torch.empty(4, dtype=torch.int64, device=dev)
    raise_for(ctx.lib.bn_base_counts_batch_dev(ctx.handle, _stream(), _ptr(words), words.numel(), _ptr(word_offsets),
                                               _ptr(lens), n, _ptr(counts4), _ptr(gc), _ptr(totals)))
    return counts4, gc, totals


def base_counts_fixed(words, n_reads: int, read_len: int, counts4=None, gc=None, totals=None):
    """Per-read counts / gc / totals for ``n_reads`` fixed-length reads of ``ceil(read_len/32)`` words each."""
    ctx = _ctx_for(words)
    dev = words.device
    counts4 = counts4 if counts4 is not None else torch.empty((n_reads, 4), dtype=torch.int64, device=dev)
    gc = gc if gc is not None else torch.empty(n_reads, dtype=torch.float64, device=dev)
    totals = totals if totals is not None else torch.empty(4, dtype=torch.int64, device=dev)
    if words.numel() < n_reads * words_for(read_len):
        raise _lib.NucleotideError.InvalidLength(read_len)
    raise_for(ctx.lib.bn_base_counts_fixed_dev(ctx.handle, _stream(), _ptr(words), n_reads, read_len, _ptr(counts4), _ptr(gc),
                                               _ptr(totals)))
    return counts4, gc, totals


def encode_batch(data: torch.Tensor, offsets: torch.Tensor, max_words: int | None = None, read_status: bool = False,
                 status: Status | None = None):
    """Variable-length reads -> (words, word_offsets, read_status|None, status), all on the device."""
    ctx = _ctx_for(data)
    n = offsets.numel() - 1
    dev = data.device
    if max_words is None:
        max_words = data.numel() // 32 + n
    words = torch.empty(max_words, dtype=torch.int64, device=dev)
    wo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    rs = torch.empty(n, dtype=torch.int32, device=dev) if read_status else None
    scratch = torch.empty(ctx.lib.bn_encode_batch_scratch_bytes(n, data.numel()), dtype=torch.uint8, device=dev)
    status = status or Status(dev)
    raise_for(ctx.lib.bn_encode_batch_dev(ctx.handle, _stream(), _ptr(data), _ptr(offsets), n, data.numel(), _ptr(words), _ptr(wo),
                                          _ptr(rs), _ptr(status.word), _ptr(scratch)))
    return words, wo, rs, status


class FastqStatus:
    """Device-side status of ``fastq_encode``: [0] min(text offset << 8 | byte) over invalid bases, [1] min(record << 8 | fault)."""

    def __init__(self, device):
        self.word = torch.empty(2, dtype=torch.int64, device=device)
        self.ctx = api.default_context(torch.device(device).index or 0)
        self.n_lines, self.seq_offsets, self.n_reads, self.fasta = 0, None, 0, False

    def check(self):
        """Synchronises the current stream; raises ``FastqError`` first, else ``NucleotideError.InvalidBase``."""
        err = BnError()
        fetch = self.ctx.lib.bn_fasta_status_fetch if self.fasta else self.ctx.lib.bn_fastq_status_fetch
        rc = fetch(self.ctx.handle, _stream(), _ptr(self.word), self.n_lines, _ptr(self.seq_offsets),
                                                self.n_reads, C.byref(err))
        if rc == 1:
            e = _lib.NucleotideError.InvalidBase(err.base)
            e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
            raise e
        raise_for(rc, err)


def fasta_encode(text: torch.Tensor, status: FastqStatus | None = None):
    """One-sequence-line FASTA text resident in HBM -> the same outputs as ``fastq_encode``."""
    return fastq_encode(text, status, _fasta=True)


def fastq_encode(text: torch.Tensor, status: FastqStatus | None = None, _fasta: bool = False):
    """FASTQ text resident in HBM -> (words, word_offsets, seq_offsets, seq_lens, status), all on the device.  Two small
    read-backs size the outputs (number of lines, number of words); ``status.check()`` reports faults."""
    ctx = _ctx_for(text)
    dev, n_bytes = text.device, text.numel()
    status = status or FastqStatus(dev)
    scratch = torch.empty(max(16, ctx.lib.bn_fastq_scratch_bytes(n_bytes)), dtype=torch.uint8, device=dev)
    n_lines_t = torch.zeros(1, dtype=torch.int64, device=dev)
    L = ctx.lib
    count, index, encode_ = (L.bn_fasta_count_dev, L.bn_fasta_index_dev, L.bn_fasta_encode_dev) if _fasta else \
        (L.bn_fastq_count_dev, L.bn_fastq_index_dev, L.bn_fastq_encode_dev)
    raise_for(count(ctx.handle, _stream(), _ptr(text), n_bytes, _ptr(scratch), _ptr(n_lines_t)))
    n_lines = int(n_lines_t.item())
    n = n_lines // (2 if _fasta else 4)
    iscratch = torch.empty(max(16, ctx.lib.bn_fastq_index_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    so = torch.empty(max(1, n), dtype=torch.int64, device=dev)
    sl = torch.empty(max(1, n), dtype=torch.int64, device=dev)
    wo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    raise_for(index(ctx.handle, _stream(), _ptr(text), n_bytes, n, _ptr(scratch), _ptr(iscratch), _ptr(so), _ptr(sl),
                                         _ptr(wo), _ptr(status.word)))
    n_words = int(wo[n].item())
    words = torch.empty(max(1, n_words), dtype=torch.int64, device=dev)
    raise_for(encode_(ctx.handle, _stream(), _ptr(text), n_bytes, n, _ptr(scratch), _ptr(so), _ptr(sl), _ptr(wo),
                                          _ptr(words), _ptr(status.word)))
    status.n_lines, status.seq_offsets, status.n_reads, status.fasta = n_lines, so, n, _fasta
    return words[:n_words], wo, so[:n], sl[:n], status


class SplitStatus:
    """Device-side status of ``split_packed_batch``: min over failing reads of (read << 1 | kind)."""

    def __init__(self, device):
        self.word = torch.empty(1, dtype=torch.int64, device=device)

    def check(self, lens: torch.Tensor, idx: torch.Tensor):
        """Synchronises (reads the word back); raises the first failing read's error."""
        key = int(self.word.item()) & api.M64
        if key == api.M64:
            return
        r = key >> 1
        if key & 1:
            e = _lib.NucleotideError.InvalidLength(int(lens[r].item()))
        else:
            e = _lib.NucleotideError.IndexOutOfBounds(int(idx[r].item()), int(lens[r].item()))
        e.record = r
        raise e


def split_packed_batch(words: torch.Tensor, word_offsets: torch.Tensor, lens: torch.Tensor, idx: torch.Tensor,
                       status: SplitStatus | None = None):
    """Batched ``split_packed`` on the device: returns (left, left_offsets, right, right_offsets, status).
    ``left`` / ``right`` are over-allocated (n_words + n_reads / n_words); the valid prefix length is the last
    entry of the offsets arrays."""
    ctx = _ctx_for(words if words.numel() else lens)
    n = lens.numel()
    dev = lens.device
    if word_offsets.numel() != n + 1 or idx.numel() != n:
        raise ValueError("word_offsets needs n_reads + 1 entries, idx n_reads")
    left = torch.empty(words.numel() + n, dtype=torch.int64, device=dev)
    right = torch.empty(words.numel(), dtype=torch.int64, device=dev)
    lo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(ctx.lib.bn_split_packed_scratch_bytes(n), dtype=torch.uint8, device=dev)
    status = status or SplitStatus(dev)
    raise_for(ctx.lib.bn_split_packed_batch_dev(ctx.handle, _stream(), _ptr(words), _ptr(word_offsets), _ptr(lens), _ptr(idx), n,
                                                _ptr(left), _ptr(lo), _ptr(right), _ptr(ro), _ptr(status.word), _ptr(scratch)))
    return left, lo, right, ro, status


def kmers(seq: torch.Tensor, k: int, out: torch.Tensor | None = None, status: Status | None = None):
    """Every k-mer of an ASCII sequence: (int64[n-k+1], status); window i = ``as_2bit(seq[i:i+k])``."""
    ctx = _ctx_for(seq)
    n = seq.numel()
    if out is None:
        out = torch.empty(max(0, n - k + 1), dtype=torch.int64, device=seq.device)
    status = status or Status(seq.device)
    rc = ctx.lib.bn_kmers_dev(ctx.handle, _stream(), _ptr(seq), n, k, _ptr(out), _ptr(status.word))
    if rc == 2:
        raise _lib.NucleotideError.SequenceTooLong(k)
    raise_for(rc)
    return out, status


def kmers_batch(data: torch.Tensor, offsets: torch.Tensor, k: int, out_words: int | None = None, status: Status | None = None):
    """Per-read k-mer windows on the device: (words int64[out_words], out_offsets int64[n+1], status)."""
    ctx = _ctx_for(data)
    n, dev = offsets.numel() - 1, data.device
    out = torch.empty(data.numel() if out_words is None else out_words, dtype=torch.int64, device=dev)
    oo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(ctx.lib.bn_kmers_batch_scratch_bytes(n, data.numel()), dtype=torch.uint8, device=dev)
    status = status or Status(dev)
    rc = ctx.lib.bn_kmers_batch_dev(ctx.handle, _stream(), _ptr(data), _ptr(offsets), n, data.numel(), k, _ptr(out), _ptr(oo),
                                    _ptr(status.word), _ptr(scratch))
    if rc == 2:
        raise _lib.NucleotideError.SequenceTooLong(k)
    raise_for(rc)
    return out, oo, status


class QueryStatus:
    """Device-side status of ``slice_batch`` / ``get_batch``: the smallest failing query index."""

    def __init__(self, device):
        self.word = torch.empty(1, dtype=torch.int64, device=device)

    def first_failing(self):
        """Synchronises (reads the word back): index of the first failing query, or None."""
        q = int(self.word.item()) & api.M64
        return None if q == api.M64 else q


def slice_batch(words, word_offsets, lens, q_read, q_start, q_end, out_bytes: int, status: QueryStatus | None = None):
    """Batched ``PackedSequence::slice`` on the device: (bytes uint8[out_bytes], out_offsets int64[nq+1], status).
    ``out_bytes`` must cover the sum of the valid range lengths (failing queries take no room)."""
    ctx = _ctx_for(lens)
    nq, dev = q_read.numel(), lens.device
    out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    oo = torch.empty(nq + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(ctx.lib.bn_slice_batch_scratch_bytes(nq), dtype=torch.uint8, device=dev)
    status = status or QueryStatus(dev)
    raise_for(ctx.lib.bn_slice_batch_dev(ctx.handle, _stream(), _ptr(words), _ptr(word_offsets), _ptr(lens), lens.numel(), _ptr(q_read),
                                         _ptr(q_start), _ptr(q_end), nq, _ptr(out), _ptr(oo), _ptr(status.word), _ptr(scratch)))
    return out, oo, status


def get_batch(words, word_offsets, lens, q_read, q_index, status: QueryStatus | None = None):
    """Batched ``PackedSequence::get`` on the device: (bytes uint8[nq], status)."""
    ctx = _ctx_for(lens)
    nq, dev = q_read.numel(), lens.device
    out = torch.empty(nq, dtype=torch.uint8, device=dev)
    status = status or QueryStatus(dev)
    raise_for(ctx.lib.bn_get_batch_dev(ctx.handle, _stream(), _ptr(words), _ptr(word_offsets), _ptr(lens), lens.numel(), _ptr(q_read),
                                       _ptr(q_index), nq, _ptr(out), _ptr(status.word)))
    return out, status


def synth_words(seed: int, stream_id: int, first_word: int, n_words: int, device="cuda") -> torch.Tensor:
    ctx = api.default_context(torch.device(device).index or 0)
    out = torch.empty(n_words, dtype=torch.int64, device=device)
    raise_for(ctx.lib.bn_synth_words_dev(ctx.handle, _stream(), seed, stream_id, first_word, n_words, _ptr(out)))
    return out


def synth_ascii(seed: int, stream_id: int, first_base: int, n: int, device="cuda") -> torch.Tensor:
    ctx = api.default_context(torch.device(device).index or 0)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    raise_for(ctx.lib.bn_synth_ascii_dev(ctx.handle, _stream(), seed, stream_id, first_base, n, _ptr(out)))
    return out
