#!/usr/bin/env python
"""BASELINE.json configs[3] under torchrun: hdist over 2^30 pairs of packed 32-mers plus base_counts / gc_content
on 10 M x 150 bp reads, sharded over N GPUs (one process per GPU) with ONE NCCL all-reduce of the four base
counters (and one of the hdist partial).  Every rank generates only its own shard (counter-based streams), so
there is no data-path collective; the reduced totals are checked against closed-form expectations that do not
depend on N: the same totals must come out at N = 1, 2, 4, 8.

    torchrun --nproc-per-node N tools/bench_multi.py [--scale 1.0] [--reps 10]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
_STDOUT = os.fdopen(os.dup(1), "w")  # fd 1 goes to stderr from here on (NCCL prints its version banner on stdout);
os.dup2(2, 1)                        # the JSON line is written to the real stdout at the end

import torch
import torch.distributed as dist

from bitnuc_b200 import device as dv
from bitnuc_b200 import sharding as sh

SEED = 0x5EEDB17C0DE5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return float(t.item())

    out = {"n_gpus": world}
    # ---- hdist over pairs, sharded by pair index on even boundaries (16-byte aligned shards)
    n_pairs = int((1 << 30) * args.scale)
    p0, p1 = sh.shard_range(n_pairs, rank, world, align=2)
    u = dv.synth_words(SEED, 2, p0, p1 - p0, device=dev)
    v = dv.synth_words(SEED, 3, p0, p1 - p0, device=dev)
    d = torch.empty(p1 - p0, dtype=torch.int32, device=dev)
    tot = torch.empty(1, dtype=torch.int64, device=dev)

    def hd():
        dv.hdist_pairs(u, v, 32, out=d)
        dv.hdist(u, v, 32 * (p1 - p0), out=tot)
        sh.allreduce_sum(tot)

    ms = timed(hd)
    total = int(tot.item())
    check = d.sum(dtype=torch.int64)
    sh.allreduce_sum(check)
    assert int(check.item()) == total, "sum of per-pair distances != whole-sequence distance"
    out["hdist"] = {"pairs": n_pairs, "ms": ms, "Gpairs_s": n_pairs / ms / 1e6, "total_mismatches": total,
                    "GB_s_aggregate": (20 + 16) * n_pairs / ms / 1e6}
    del u, v, d

    # ---- base_counts / gc on fixed-length reads, sharded by read index, NCCL all-reduce of the four counters
    n_reads = int(10_000_000 * args.scale)
    r0, r1 = sh.shard_range(n_reads, rank, world)
    words = dv.synth_words(SEED, 4, 5 * r0, 5 * (r1 - r0), device=dev)
    counts4 = torch.empty((r1 - r0, 4), dtype=torch.int64, device=dev)
    gcs = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    totals = torch.empty(4, dtype=torch.int64, device=dev)

    def bc():
        dv.base_counts_fixed(words, r1 - r0, 150, counts4=counts4, gc=gcs, totals=totals)
        sh.allreduce_counts(totals)

    ms = timed(bc)
    t = totals.tolist()
    assert sum(t) == 150 * n_reads
    out["base_counts"] = {"reads": n_reads, "ms": ms, "Greads_s": n_reads / ms / 1e6, "totals": t, "gc_global": sh.gc_from_counts(t),
                          "GB_s_aggregate": 80 * n_reads / ms / 1e6}
    del words, counts4, gcs

    # ---- FASTQ text -> records -> packed reads, strong scaling: 2 x 10^7 reads of 150 bp in all (6.5 GB of text), every
    # rank builds and parses only its own contiguous range of records (what sharding.shard_fastq_text cuts from a real
    # file); the number of reads and of output words is all-reduced, no data-path collective
    n_reads = int(20_000_000 * args.scale)
    r0, r1 = sh.shard_range(n_reads, rank, world)
    nr, rl, hdr = r1 - r0, 150, 23
    rec = hdr + rl + 1 + 2 + rl + 1
    t2 = torch.full((nr, rec), ord("I"), dtype=torch.uint8, device=dev)
    t2[:, 0] = ord("@")
    t2[:, 1 : hdr - 1] = ord("h")
    t2[:, hdr - 1] = 10
    b0 = (r0 * rl) // 32 * 32                                   # the generator starts on a word boundary
    seqs = dv.synth_ascii(SEED, 7, b0, nr * rl + 32, device=dev)[r0 * rl - b0 : r0 * rl - b0 + nr * rl]
    t2[:, hdr : hdr + rl] = seqs.view(nr, rl)
    del seqs
    t2[:, hdr + rl] = 10
    t2[:, hdr + rl + 1] = ord("+")
    t2[:, hdr + rl + 2] = 10
    t2[:, rec - 1] = 10
    text = t2.view(-1)
    ctx = dv.api.default_context(local)
    n_bytes = text.numel()
    P = dv._ptr
    scratch = torch.empty(ctx.lib.bn_fastq_scratch_bytes(n_bytes), dtype=torch.uint8, device=dev)
    iscratch = torch.empty(ctx.lib.bn_fastq_index_scratch_bytes(nr), dtype=torch.uint8, device=dev)
    n_lines = torch.zeros(1, dtype=torch.int64, device=dev)
    so, sl = torch.empty(nr, dtype=torch.int64, device=dev), torch.empty(nr, dtype=torch.int64, device=dev)
    wo = torch.empty(nr + 1, dtype=torch.int64, device=dev)
    wpr = (rl + 31) // 32
    fwords = torch.empty(nr * wpr, dtype=torch.int64, device=dev)
    st = dv.FastqStatus(dev)

    def fq():
        dv.raise_for(ctx.lib.bn_fastq_count_dev(ctx.handle, dv._stream(), P(text), n_bytes, P(scratch), P(n_lines)))
        dv.raise_for(ctx.lib.bn_fastq_index_dev(ctx.handle, dv._stream(), P(text), n_bytes, nr, P(scratch), P(iscratch), P(so), P(sl), P(wo),
                                                P(st.word)))
        dv.raise_for(ctx.lib.bn_fastq_encode_dev(ctx.handle, dv._stream(), P(text), n_bytes, nr, P(scratch), P(so), P(sl), P(wo), P(fwords),
                                                 P(st.word)))

    ms = timed(fq)
    st.n_lines, st.seq_offsets, st.n_reads = int(n_lines.item()), so, nr
    st.check()
    sizes = torch.tensor([st.n_lines // 4, int(wo[-1].item())], dtype=torch.int64, device=dev)
    sh.allreduce_sum(sizes)
    assert sizes.tolist() == [n_reads, n_reads * wpr]
    # the first packed word of this rank's first read is the generator's own word (reads are 150 bases: word-aligned every 16 reads)
    if (r0 * rl) % 32 == 0:
        assert int(fwords[0].item()) == int(dv.synth_words(SEED, 7, r0 * rl // 32, 1, device=dev)[0].item())
    text_total = n_reads * rec
    out["fastq"] = {"reads": n_reads, "text_bytes": text_total, "ms": ms, "Gbases_s": n_reads * rl / ms / 1e6,
                    "text_GB_s_aggregate": text_total / ms / 1e6}
    if rank == 0:
        print(json.dumps(out), file=_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
