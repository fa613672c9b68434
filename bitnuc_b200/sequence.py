"""PackedSequence + the GCContent / BaseCount traits -- mirrors /root/reference/src/sequence.rs and
/root/reference/src/utils/analysis.rs.  Construction, ``to_vec`` and the analysis methods ride the
CUDA path; ``get`` / ``slice`` are O(1)/O(k) host-side bit pokes on the owned words, as in the
reference (src/sequence.rs:116-135, :198-212)."""
from __future__ import annotations

import numpy as np

from . import api
from .errors import NucleotideError

_ASCII = b"ACGT"


class PackedSequence:
    """``struct PackedSequence { data: Vec<u64>, length: usize }`` (src/sequence.rs:5-9).
    Equality and hashing are over ``(data, length)`` like the derived ``PartialEq, Eq, Hash``."""

    __slots__ = ("data", "length", "_ctx")

    def __init__(self, seq, ctx: api.Context | None = None):
        """``PackedSequence::new`` (src/sequence.rs:40-52): the empty sequence skips ``encode``."""
        seq = bytes(seq) if not isinstance(seq, np.ndarray) else seq
        n = len(seq)
        self._ctx = ctx
        self.data = np.zeros(0, dtype=np.uint64) if n == 0 else api.encode_np(seq, ctx)
        self.length = n

    @classmethod
    def new(cls, seq, ctx: api.Context | None = None) -> "PackedSequence":
        return cls(seq, ctx)

    def __len__(self) -> int:
        return self.length

    def len(self) -> int:
        return self.length

    def is_empty(self) -> bool:
        return self.length == 0

    def get(self, index: int) -> int:
        """src/sequence.rs:116-135"""
        if index < 0 or index >= self.length:
            raise NucleotideError.IndexOutOfBounds(index, self.length)
        return _ASCII[(int(self.data[index // 32]) >> ((index % 32) * 2)) & 3]

    def slice(self, start: int, end: int) -> bytes:
        """src/sequence.rs:198-212 (``range.start > range.end || range.end > self.length``)."""
        if start > end or end > self.length:
            raise NucleotideError.InvalidRange(start, end, self.length)
        return bytes(self.get(i) for i in range(start, end))

    def to_vec(self) -> bytes:
        """src/sequence.rs:260-262: the whole sequence, decoded on the GPU."""
        if self.length == 0:
            return b""
        return api.decode_np(self.data, self.length, self._ctx).tobytes()

    def base_counts(self) -> list:
        """``BaseCount::base_counts`` (src/utils/analysis.rs:19-39) -> [A, C, G, T]."""
        return api.base_counts_gc(self.data, self.length, self._ctx)[0]

    def gc_content(self) -> float:
        """``GCContent::gc_content`` (src/utils/analysis.rs:3-17)."""
        return api.base_counts_gc(self.data, self.length, self._ctx)[1]

    def __eq__(self, other):
        return (isinstance(other, PackedSequence) and self.length == other.length
                and np.array_equal(self.data, other.data))

    def __hash__(self):
        return hash((self.data.tobytes(), self.length))

    def __repr__(self):
        return f"PackedSequence {{ data: {[int(w) for w in self.data]}, length: {self.length} }}"
