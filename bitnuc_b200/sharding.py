"""Multi-GPU host logic: one process per GPU (torch.distributed), contiguous range sharding.

The codec / k-mer / pair kernels have no exchange step: every 32-base word, record and pair is
independent (/root/reference/src/utils/packing/avx.rs:138-145 carries nothing between words), so
ranks just take contiguous shards and there is no data-path collective.  The only reductions are
  * base-count totals: sum of 4 x u64  (all_reduce SUM over NCCL; the north-star's single collective),
  * whole-sequence hdist: sum of one u64,
  * error parity: min over ranks of the first invalid global offset.
Everything here is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

NO_ERROR = (1 << 63) - 1  # int64 sentinel for "no invalid base on this rank"


def shard_range(n_units: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous shard [start, stop) of ``n_units`` for ``rank``; interior cuts are multiples of
    ``align`` and shard sizes differ by at most ``align``."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    blocks = -(-n_units // align)
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * align, n_units), min(b1 * align, n_units)


def shard_bases(n_bases: int, rank: int, world: int) -> tuple[int, int]:
    """Base range of a rank, cut on multiples of 64 bases: the shard's ASCII starts 64-byte aligned
    and its packed words start 16-byte aligned, so every rank keeps 128-bit accesses."""
    return shard_range(n_bases, rank, world, align=64)


def shard_reads_by_volume(offsets: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Cut a read batch into ``world`` contiguous read ranges of near-equal byte volume (cfg 5):
    cuts fall on read boundaries at the first read whose start reaches the ideal byte cut."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    lo, hi = int(offsets[0]), int(offsets[-1])
    cuts = [0]
    for g in range(1, world):
        target = lo + (hi - lo) * g // world
        cuts.append(max(cuts[-1], min(n, int(np.searchsorted(offsets[:-1], target, side="left")))))
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def fastq_record_start(text, pos: int) -> int:
    """The first record boundary at or after byte ``pos`` of a FASTQ text (``len(text)`` when there is none): the
    start of a line that opens with '@' and whose second-next line opens with '+'.  The second condition is what
    tells a header from a quality line that happens to start with '@' (its second-next line is a sequence)."""
    t = memoryview(text).cast("B") if not isinstance(text, np.ndarray) else text
    n = len(t)
    p = pos
    if p > 0:  # move to the start of the next line
        while p < n and t[p - 1] != 10:
            p += 1
    while p < n:
        if t[p] == 64:  # '@'
            q, lines = p, 0
            while q < n and lines < 2:
                if t[q] == 10:
                    lines += 1
                q += 1
            if lines == 2 and q < n and t[q] == 43:  # '+'
                return p
        while p < n and t[p] != 10:
            p += 1
        p += 1
    return n


def fasta_record_start(text, pos: int) -> int:
    """The first record boundary at or after byte ``pos`` of a FASTA text: the start of a line that opens with '>'
    (a sequence line never does)."""
    t = memoryview(text).cast("B") if not isinstance(text, np.ndarray) else text
    n = len(t)
    p = pos
    if p > 0:
        while p < n and t[p - 1] != 10:
            p += 1
    while p < n and t[p] != 62:  # '>'
        while p < n and t[p] != 10:
            p += 1
        p += 1
    return min(p, n)


def shard_fasta_text(text, world: int) -> list[tuple[int, int]]:
    """``shard_fastq_text`` for FASTA text (``bn_fasta_*``)."""
    n = len(text)
    cuts = [0]
    for g in range(1, world):
        cuts.append(max(cuts[-1], fasta_record_start(text, n * g // world)))
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def shard_fastq_text(text, world: int) -> list[tuple[int, int]]:
    """Cut a FASTQ text into ``world`` contiguous byte ranges of near-equal size on record boundaries: every rank
    parses and encodes its own range (``bn_fastq_*``), read indices / word offsets of rank g are then offset by the
    totals of ranks < g, and the first fault is the MIN over ranks.  No data-path collective."""
    n = len(text)
    cuts = [0]
    for g in range(1, world):
        cuts.append(max(cuts[-1], fastq_record_start(text, n * g // world)))
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the four base counters (int64[4]) over all ranks, in place.  32 bytes: latency-bound."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def allreduce_sum(value: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(value, op=dist.ReduceOp.SUM, group=group)
    return value


def gc_from_counts(counts) -> float:
    """Global gc_content from the reduced integer counts, in the reference's operation order
    (/root/reference/src/utils/analysis.rs:14): (gc as f64 / len as f64) * 100.0."""
    a, c, g, t = (int(x) for x in counts)
    n = a + c + g + t
    if n == 0:
        return 0.0
    return float((np.float64(c + g) / np.float64(n)) * np.float64(100.0))


def first_error_across_ranks(local_key: int | None, shard_start: int, device=None, group=None):
    """Error parity across shards: every rank passes its local (offset << 8 | byte) status key (or
    None) and the global offset of its shard; returns (global_offset, byte) of the first invalid base
    in sequence order, or None.  One MIN all_reduce of one int64."""
    if local_key is None:
        key = NO_ERROR
    else:
        key = (((local_key >> 8) + shard_start) << 8) | (local_key & 0xFF)
    t = torch.tensor([key], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    k = int(t.item())
    return None if k == NO_ERROR else (k >> 8, k & 0xFF)
