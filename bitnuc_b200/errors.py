"""NucleotideError and friends -- mirrors /root/reference/src/error.rs:4-47."""
from __future__ import annotations


class NucleotideError(Exception):
    """One exception class, six variants, same payloads and Display strings as the reference enum.

    ``variant`` is the Rust variant name; ``payload`` its fields in declaration order.
    Equality compares (variant, payload), like the reference's ``#[derive(PartialEq, Eq)]``.
    """

    VARIANTS = ("InvalidBase", "SequenceTooLong", "InvalidLength", "IndexOutOfBounds", "InvalidRange",
                "Unsupported")

    def __init__(self, variant: str, *payload: int):
        assert variant in self.VARIANTS, variant
        self.variant = variant
        self.payload = tuple(int(p) for p in payload)
        super().__init__(self._display())

    # constructors named like the Rust variants -------------------------------------------------
    @classmethod
    def InvalidBase(cls, base: int):
        return cls("InvalidBase", base)

    @classmethod
    def SequenceTooLong(cls, length: int):
        return cls("SequenceTooLong", length)

    @classmethod
    def InvalidLength(cls, length: int):
        return cls("InvalidLength", length)

    @classmethod
    def IndexOutOfBounds(cls, index: int, length: int):
        return cls("IndexOutOfBounds", index, length)

    @classmethod
    def InvalidRange(cls, start: int, end: int, length: int):
        return cls("InvalidRange", start, end, length)

    @classmethod
    def Unsupported(cls):
        return cls("Unsupported")

    def _display(self) -> str:  # src/error.rs:20-45
        p = self.payload
        if self.variant == "InvalidBase":
            return f"Invalid nucleotide base: {p[0]}"  # the byte as a decimal integer
        if self.variant == "SequenceTooLong":
            return f"Sequence length {p[0]} exceeds maximum"
        if self.variant == "InvalidLength":
            return f"Invalid length: {p[0]}"
        if self.variant == "IndexOutOfBounds":
            return f"Index {p[0]} out of bounds for sequence of length {p[1]}"
        if self.variant == "InvalidRange":
            return f"Invalid range {p[0]}..{p[1]} for sequence of length {p[2]}"
        return "Unsupported architecture"

    def key(self):
        return (self.variant,) + self.payload

    def __eq__(self, other):
        return isinstance(other, NucleotideError) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())

    def __repr__(self):
        return f"NucleotideError::{self.variant}{self.payload if self.payload else ''}"


class ReferencePanic(RuntimeError):
    """Raised where the reference panics (e.g. ``encode(b"")``, src/utils/packing/avx.rs:138)."""


class BitnucCudaError(RuntimeError):
    """CUDA / argument failure below the C ABI: no reference variant exists for these and there is
    no CPU fallback, so they surface loudly."""


class FastqError(ValueError):
    """Malformed FASTQ / FASTA text (``BN_ERR_FASTQ``): ``record`` is the first faulty record in file order, ``fault`` is
    1 header without '@', 2 separator without '+', 3 quality and sequence lengths differ, 4 text ends inside the record."""

    FAULTS = {1: "header line does not start with '@' (FASTA: '>')", 2: "separator line does not start with '+'",
              3: "quality and sequence lengths differ", 4: "text ends inside the record"}

    def __init__(self, record: int, fault: int):
        super().__init__(f"record {record}: {self.FAULTS.get(fault, 'malformed record')}")
        self.record, self.fault = int(record), int(fault)

    def key(self):
        return (self.record, self.fault)
