#!/usr/bin/env python
"""BASELINE.json configs[4] end to end: a variable-length read batch (50 bp - 10 kbp, offset-indexed, ~32 Gbases at
--scale 1) with injected N bases, encoded through the HOST-pointer entry point (bn_encode_batch: pinned host buffers,
PCIe H2D + kernels + D2H inside the timed region), sharded over N GPUs by byte volume on read boundaries.

    torchrun --nproc-per-node N tools/bench_cfg5_e2e.py [--scale 0.25] [--reps 3]

Every rank lays out the whole batch (lengths come from a counter hash), takes its contiguous read range, fills its
bytes with the device generator and copies them to pinned host memory.  Two timed legs: a clean batch, and the batch
with the injected N bases (per-read status variant: nothing short-circuits).  Error parity: the first invalid base in
input order over the whole batch = MIN over ranks of (global byte offset << 8 | byte), compared with the closed form.
Time = wall clock around the call, barrier on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
_STDOUT = os.fdopen(os.dup(1), "w")  # fd 1 goes to stderr from here on (NCCL prints its version banner on stdout);
os.dup2(2, 1)                        # the JSON line is written to the real stdout at the end

import numpy as np
import torch
import torch.distributed as dist

import bitnuc_b200 as bn
from bitnuc_b200 import device as dv
from bitnuc_b200 import sharding as sh
from bitnuc_b200 import synth
from bitnuc_b200._lib import BnError

SEED = synth.DEFAULT_SEED


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--reps", type=int, default=3)
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

    lens = synth.cfg5_read_lengths(int(32e9 * args.scale), SEED)
    n_total = lens.size
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    r_lo, r_hi = sh.shard_reads_by_volume(offsets, world)[rank]
    b_lo, b_hi = int(offsets[r_lo]), int(offsets[r_hi])
    n_reads, n_bytes = r_hi - r_lo, b_hi - b_lo
    ctx = bn.Context(local)
    h_bytes = ctx.pinned_empty(n_bytes, np.uint8)
    a_lo = b_lo // 32 * 32
    h_bytes[:] = dv.synth_ascii(SEED, 5, a_lo, b_hi - a_lo, device=dev)[b_lo - a_lo:].cpu().numpy()
    h_off = ctx.pinned_empty(n_reads + 1, np.uint64)
    h_off[:] = offsets[r_lo : r_hi + 1] - np.uint64(b_lo)
    h_words = ctx.pinned_empty(n_bytes // 32 + n_reads, np.uint64)
    h_wo = ctx.pinned_empty(n_reads + 1, np.uint64)
    h_rs = ctx.pinned_empty(n_reads, np.uint32)

    def call(with_status):
        err = BnError()
        rc = ctx.lib.bn_encode_batch(ctx.handle, h_bytes.ctypes.data_as(C.c_void_p), h_off.ctypes.data_as(C.c_void_p), n_reads,
                                     h_words.ctypes.data_as(C.c_void_p), h_wo.ctypes.data_as(C.c_void_p),
                                     h_rs.ctypes.data_as(C.c_void_p) if with_status else None, C.byref(err))
        return rc, err

    def timed(with_status):
        call(with_status)  # warm-up: sizes the device staging buffers
        best = float("inf")
        for _ in range(args.reps):
            barrier()
            t0 = time.perf_counter()
            rc, err = call(with_status)
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            barrier()
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best, rc, err

    # ---- clean batch
    t_clean, rc, err = timed(False)
    assert rc == 0, rc
    n_words = int(h_wo[n_reads])
    assert n_words == int(((lens[r_lo:r_hi] + np.uint64(31)) // np.uint64(32)).sum())
    for r in {0, n_reads // 2, n_reads - 1}:  # round trip of a few reads through the device decode
        w0, ln, o = int(h_wo[r]), int(lens[r_lo + r]), int(h_off[r])
        back = dv.decode(torch.from_numpy(h_words[w0 : w0 + (ln + 31) // 32].view(np.int64)).to(dev), ln)
        assert np.array_equal(back.cpu().numpy(), h_bytes[o : o + ln])

    # ---- injected N: read r gets 'N' at h2(r) mod len iff h1(r) mod 100003 == 0 (whole-batch rule, rank-local bytes)
    victims, pos = synth.cfg5_injected_n(n_total, lens)
    if victims.size == 0:
        victims, pos = np.array([n_total // 3]), np.array([int(lens[n_total // 3]) // 2], dtype=np.int64)
    mine = (victims >= r_lo) & (victims < r_hi)
    h_bytes[(offsets[victims[mine]] - np.uint64(b_lo)).astype(np.int64) + pos[mine]] = ord("N")
    t_inj, rc, err = timed(True)
    local_key = (int(err.offset) << 8 | int(err.base)) if rc == 1 else None
    assert (rc == 1) == bool(mine.any()), (rc, int(mine.sum()))
    first = sh.first_error_across_ranks(local_key, b_lo, device=dev)
    expect_off = int(offsets[victims[0]]) + int(pos[0])
    assert first == (expect_off, ord("N")), (first, expect_off)
    bad = np.flatnonzero(h_rs[:n_reads] != 0xFFFFFFFF)
    assert np.array_equal(bad + r_lo, victims[mine]) and np.array_equal(h_rs[bad].astype(np.int64), pos[mine])

    total_bases = int(offsets[-1])
    h2d = total_bases + 8 * (n_total + world)
    d2h = 8 * int(((lens + np.uint64(31)) // np.uint64(32)).sum()) + 8 * (n_total + world)
    if rank == 0:
        print(json.dumps({
            "workload": "BASELINE.json configs[4]: variable-length read batch encode, end to end incl. PCIe", "n_gpus": world,
            "scale": args.scale, "reads": int(n_total), "bases": total_bases, "h2d_bytes": h2d, "d2h_bytes": d2h,
            "clean": {"ms": t_clean * 1e3, "Gbases_s": total_bases / t_clean / 1e9, "pcie_GB_s_aggregate": (h2d + d2h) / t_clean / 1e9},
            "injected_N_per_read_status": {"ms": t_inj * 1e3, "Gbases_s": total_bases / t_inj / 1e9, "injected": int(victims.size),
                                           "first_error": {"record": int(victims[0]), "position": int(pos[0]), "offset": expect_off,
                                                           "byte": ord("N")}, "parity": "ok"},
            "timing": "wall clock around bn_encode_batch (pinned host buffers), barrier both sides, max over ranks, best of reps"}),
              file=_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
