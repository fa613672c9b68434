"""``MultiContext`` -- the Python face of ``bn_multi`` (include/bitnuc_cuda.h, "multi-GPU"): one process, N devices.

The host-pointer methods have the signatures of the module-level array functions (``encode_np``, ``decode_np``, ...)
and return what those return: the library cuts the input into contiguous shards (base ranges on 64-base boundaries,
records / pairs / reads by index, variable-length reads by byte volume on read boundaries), runs every shard on its own
device at once -- one ``bn_ctx`` with its own streams and pinned staging and one host thread per device -- and merges
the results (first error in input order, counters summed by the collective: ``ncclAllReduce`` or our mailbox kernel
over NVLink peer memory).  The ``*_dev`` methods take one device tensor per shard.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BnError, raise_for
from .api import Context, _check_offsets, _p, _u8, _u64
from .errors import NucleotideError

REDUCE_NCCL, REDUCE_P2P = 0, 1


class _ShardContext(Context):
    """A view of the bn_ctx of one shard (owned by the bn_multi, never destroyed from here)."""

    def __init__(self, lib, handle, device):
        self.lib, self.handle, self.device = lib, handle, device

    def close(self):
        self.handle = None


class MultiContext:
    def __init__(self, devices=None, reduce: str | int = "nccl"):
        """``devices``: list of device ordinals (``None``: every visible device; an int n: devices 0..n-1).  A device may
        be named more than once (how a one-GPU box exercises the sharding) -- then ``reduce`` must be ``"p2p"``."""
        self.lib = _lib.load()
        mode = {"nccl": REDUCE_NCCL, "p2p": REDUCE_P2P}.get(reduce, reduce)
        if isinstance(devices, int):
            devices = list(range(devices))
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.bn_multi_create(None, 0, mode, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.bn_multi_create(arr, len(devices), mode, C.byref(h))
        raise_for(rc)
        self.handle = h
        self.n = self.lib.bn_multi_size(h)
        self.contexts = []
        for i in range(self.n):
            ch = C.c_void_p(self.lib.bn_multi_ctx(h, i))
            self.contexts.append(_ShardContext(self.lib, ch, self.lib.bn_ctx_device(ch)))
        self.devices = [c.device for c in self.contexts]

    def close(self):
        if getattr(self, "handle", None):
            for c in self.contexts:
                c.close()
            self.lib.bn_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def reduce(self) -> str:
        return ("nccl", "p2p")[self.lib.bn_multi_reduce(self.handle)]

    @property
    def nccl_version(self) -> int:
        return self.lib.bn_multi_nccl_version(self.handle)

    def set_chunk_bytes(self, n: int):
        raise_for(self.lib.bn_multi_set_chunk_bytes(self.handle, n))

    def synchronize(self):
        raise_for(self.lib.bn_multi_synchronize(self.handle))

    def shard_units(self, n_units: int, align: int = 1) -> list[int]:
        st = (C.c_size_t * (self.n + 1))()
        raise_for(self.lib.bn_multi_shard_units(self.handle, n_units, align, st))
        return list(st)

    def shard_reads(self, offsets) -> list[int]:
        off = _u64(offsets)
        st = (C.c_size_t * (self.n + 1))()
        raise_for(self.lib.bn_multi_shard_reads(self.handle, _p(off), max(0, off.size - 1), st))
        return list(st)

    # ------------------------------------------------------------------ host-pointer calls
    def encode_np(self, seq, out: np.ndarray | None = None) -> np.ndarray:
        a = _u8(seq)
        need = (a.size + 31) // 32
        words = out if out is not None else np.empty(max(1, need), dtype=np.uint64)
        if words.dtype != np.uint64 or words.size < need:
            raise ValueError("out needs ceil(n/32) uint64 words")
        n_words, err = C.c_size_t(0), BnError()
        rc = self.lib.bn_multi_encode(self.handle, _p(a), a.size, _p(words), C.byref(n_words), C.byref(err))
        if rc == 1:
            e = NucleotideError.InvalidBase(err.base)
            e.offset, e.n_words = int(err.offset), int(n_words.value)
            raise e
        raise_for(rc, err)
        return words[: n_words.value]

    def decode_np(self, ebuf, n_bases: int, out: np.ndarray | None = None) -> np.ndarray:
        w = _u64(ebuf)
        res = out if out is not None else np.empty(max(1, n_bases), dtype=np.uint8)
        if res.dtype != np.uint8 or res.size < n_bases:
            raise ValueError("out needs n_bases bytes")
        err = BnError()
        raise_for(self.lib.bn_multi_decode(self.handle, _p(w), w.size, n_bases, _p(res), C.byref(err)), err)
        return res[:n_bases]

    def as_2bit_batch(self, recs, n: int, k: int, stride: int | None = None) -> np.ndarray:
        a = _u8(recs)
        stride = k if stride is None else stride
        if n and 0 < k <= 32 and (stride < k or a.size < (n - 1) * stride + k):
            raise ValueError("recs is shorter than (n - 1) * stride + k")
        out = np.empty(max(1, n), dtype=np.uint64)
        err = BnError()
        rc = self.lib.bn_multi_as_2bit_batch(self.handle, _p(a), n, k, stride, _p(out), C.byref(err))
        if rc == 1:
            e = NucleotideError.InvalidBase(err.base)
            e.record, e.offset = int(err.record), int(err.offset)
            raise e
        raise_for(rc, err)
        return out[:n]

    def from_2bit_batch(self, packed, k: int, stride: int | None = None) -> np.ndarray:
        w = _u64(packed)
        stride = k if stride is None else stride
        n = w.size
        out = np.zeros(max(1, (n - 1) * stride + k if n and k <= 32 else 0), dtype=np.uint8)
        err = BnError()
        raise_for(self.lib.bn_multi_from_2bit_batch(self.handle, _p(w), n, k, _p(out), stride, C.byref(err)), err)
        return out[: (n - 1) * stride + k if n and k else 0]

    def hdist_total(self, ebuf1, ebuf2, n_bases: int) -> int:
        a, b = _u64(ebuf1), _u64(ebuf2)
        total, err = C.c_uint64(0), BnError()
        raise_for(self.lib.bn_multi_hdist(self.handle, _p(a), a.size, _p(b), b.size, n_bases, C.byref(total), C.byref(err)), err)
        return int(total.value)

    def hdist(self, ebuf1, ebuf2, n_bases: int) -> int:
        return self.hdist_total(ebuf1, ebuf2, n_bases) & 0xFFFFFFFF   # the reference's u32 accumulator (multi.rs:130)

    def hdist_pairs(self, u, v, length: int, out: np.ndarray | None = None) -> np.ndarray:
        a, b = _u64(u), _u64(v)
        if a.size != b.size:
            raise ValueError("u and v differ in length")
        res = out if out is not None else np.empty(max(1, a.size), dtype=np.uint32)
        err = BnError()
        raise_for(self.lib.bn_multi_hdist_pairs(self.handle, _p(a), _p(b), a.size, length, _p(res), C.byref(err)), err)
        return res[: a.size]

    def base_counts_gc(self, words, n_bases: int):
        w = _u64(words)
        counts, gc, err = (C.c_uint64 * 4)(), C.c_double(0.0), BnError()
        raise_for(self.lib.bn_multi_base_counts(self.handle, _p(w), w.size, n_bases, counts, C.byref(gc), C.byref(err)), err)
        return [int(c) for c in counts], float(gc.value)

    def base_counts_batch(self, words, word_offsets, lens):
        w, wo, ln = _u64(words), _u64(word_offsets), _u64(lens)
        n = ln.size
        if wo.size < n:
            raise ValueError("word_offsets needs one entry per read")
        counts4 = np.empty((max(1, n), 4), dtype=np.uint64)
        gc = np.empty(max(1, n), dtype=np.float64)
        totals, err = (C.c_uint64 * 4)(), BnError()
        rc = self.lib.bn_multi_base_counts_batch(self.handle, _p(w), w.size, _p(wo), _p(ln), n, _p(counts4), _p(gc), totals, C.byref(err))
        if rc == 3:
            e = NucleotideError.InvalidLength(err.a)
            e.record = int(err.record)
            raise e
        raise_for(rc, err)
        return counts4[:n], gc[:n], [int(t) for t in totals]

    def encode_batch(self, data, offsets, per_read_status: bool = False, out_words: np.ndarray | None = None):
        a, off = _u8(data), _u64(offsets)
        n = off.size - 1
        if n < 0:
            raise ValueError("offsets needs n_reads + 1 entries")
        _check_offsets(off, a.size)
        max_words = int((off[-1] - off[0]) // np.uint64(32)) + n if n else 0
        words = out_words if out_words is not None else np.empty(max(1, max_words), dtype=np.uint64)
        if words.dtype != np.uint64 or words.size < max_words:
            raise ValueError("out_words needs (offsets[-1] - offsets[0]) // 32 + n_reads words")
        wo = np.zeros(n + 1, dtype=np.uint64)
        status = np.empty(max(1, n), dtype=np.uint32) if per_read_status else None
        err = BnError()
        rc = self.lib.bn_multi_encode_batch(self.handle, _p(a), _p(off), n, _p(words), _p(wo), _p(status) if status is not None else None,
                                            C.byref(err))
        if rc == 1 and not per_read_status:
            e = NucleotideError.InvalidBase(err.base)
            e.record, e.position, e.offset = int(err.record), int(err.b), int(err.offset)
            raise e
        if rc != 1:
            raise_for(rc, err)
        words = words[: int(wo[n])]
        return (words, wo, status[:n]) if per_read_status else (words, wo)

    # ------------------------------------------------------------------ device-resident sharded reductions
    def _ptrs(self, tensors, allow_none=False):
        if tensors is None:
            return None
        if len(tensors) != self.n:
            raise ValueError(f"need one tensor per shard ({self.n})")
        arr = (C.c_void_p * self.n)()
        for i, t in enumerate(tensors):
            if t is None:
                if not allow_none:
                    raise ValueError("missing shard tensor")
                arr[i] = None
                continue
            if not t.is_cuda or t.device.index != self.devices[i] or not t.is_contiguous():
                raise ValueError(f"shard {i}: need a contiguous CUDA tensor on device {self.devices[i]}")
            arr[i] = t.data_ptr()
        return arr

    def _sizes(self, values):
        if len(values) != self.n:
            raise ValueError(f"need one size per shard ({self.n})")
        return (C.c_size_t * self.n)(*[int(v) for v in values])

    def base_counts_dev(self, words, n_bases, counts, gc=None):
        """Shard i = ``n_bases[i]`` bases at ``words[i]`` (int64 tensor on device i).  Enqueue-only; afterwards every
        ``counts[i]`` (int64[4]) holds the global [A,C,G,T] and ``gc[i]`` (float64[1]) the global gc_content."""
        for i, t in enumerate(words):
            if t.numel() < (int(n_bases[i]) + 31) // 32:
                raise NucleotideError.InvalidLength(int(n_bases[i]))
        raise_for(self.lib.bn_multi_base_counts_dev(self.handle, self._ptrs(words), self._sizes(n_bases), self._ptrs(counts),
                                                    self._ptrs(gc, True)))

    def base_counts_fixed_dev(self, words, n_reads, read_len: int, totals, gc=None, counts4=None, gc_reads=None):
        wpr = (read_len + 31) // 32
        for i, t in enumerate(words):
            if t.numel() < int(n_reads[i]) * wpr:
                raise NucleotideError.InvalidLength(read_len)
        raise_for(self.lib.bn_multi_base_counts_fixed_dev(self.handle, self._ptrs(words), self._sizes(n_reads), read_len,
                                                          self._ptrs(counts4, True), self._ptrs(gc_reads, True), self._ptrs(totals),
                                                          self._ptrs(gc, True)))

    def hdist_dev(self, a, b, n_bases, total):
        for i in range(self.n):
            need = (int(n_bases[i]) + 31) // 32
            if a[i].numel() < need or b[i].numel() < need:
                raise NucleotideError.InvalidLength(int(n_bases[i]))
        raise_for(self.lib.bn_multi_hdist_dev(self.handle, self._ptrs(a), self._ptrs(b), self._sizes(n_bases), self._ptrs(total)))

    def allreduce_u64_dev(self, bufs, count: int):
        raise_for(self.lib.bn_multi_allreduce_u64_dev(self.handle, self._ptrs(bufs), count))

    def last_ms(self) -> list[float]:
        ms = (C.c_float * self.n)()
        raise_for(self.lib.bn_multi_last_ms(self.handle, ms))
        return [float(x) for x in ms]
