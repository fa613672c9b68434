#!/usr/bin/env python
"""What bounds the end-to-end number at N GPUs: the host's PCIe / memory path, measured with all N ranks at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_probe.py

Every rank owns one GPU.  Part 1 times raw pinned copies of the e2e step's sizes (1.25 GB up, 1.25 GB down) on all
ranks at once -- upload only, download only, both directions -- with plain and write-combined source buffers: the
ceiling.  Part 2 times the bn_encode + bn_decode round trip of bench.py's e2e leg under the knobs that could move it:
stage chunk size, rank <-> core pinning, the two-thread pipelined form, and staggered directions (odd ranks decode while
even ranks encode).  Wall clock, barrier on both sides, max over ranks; one JSON line per measurement on rank 0."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch
import torch.distributed as dist

import bitnuc_b200 as bn
from bitnuc_b200 import device as dv

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

N_BASES = int(os.environ.get("PROBE_BASES", 1_000_000_000))
UP = N_BASES + dv.words_for(N_BASES) * 8  # bytes up (= bytes down) in one encode + decode step
REPS = 4


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(**kw):
    if rank == 0:
        print(json.dumps({"n_gpus": world, **kw}), flush=True)


rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaFreeHost.argtypes = [C.c_void_p]


def host_alloc(nbytes: int, flags: int) -> int:
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, rc
    C.memset(p, 1, nbytes)
    return p.value


def raw_copies():
    d_up = torch.empty(UP, dtype=torch.uint8, device=dev)
    d_dn = torch.ones(UP, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    for label, flags in (("pinned", 1), ("write_combined", 1 | 4)):
        h_up = host_alloc(UP, flags)
        h_dn = host_alloc(UP, 1)

        def run(up, dn, pieces=1):
            barrier()
            t0 = time.perf_counter()
            for _ in range(REPS):
                step = (UP // pieces + 255) & ~255
                for o in range(0, UP, step):
                    m = min(step, UP - o)
                    if up:
                        rt.cudaMemcpyAsync(d_up.data_ptr() + o, h_up + o, m, 1, s_up.cuda_stream)
                    if dn:
                        rt.cudaMemcpyAsync(h_dn + o, d_dn.data_ptr() + o, m, 2, s_dn.cuda_stream)
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t0) / REPS)

        run(True, True)
        t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
        t_both16 = run(True, True, 16)
        say(probe="raw_copies", source=label, bytes_each_way=UP,
            h2d_gbs_aggregate=world * UP / t_up / 1e9, d2h_gbs_aggregate=world * UP / t_dn / 1e9,
            both_gbs_aggregate=2 * world * UP / t_both / 1e9, both_ms=t_both * 1e3,
            both_16_pieces_gbs_aggregate=2 * world * UP / t_both16 / 1e9)
        rt.cudaFreeHost(h_up)
        rt.cudaFreeHost(h_dn)


def pin_to_share():
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    mine = cpus[local * per:(local + 1) * per] or cpus
    os.sched_setaffinity(0, mine)
    return mine


def e2e(chunk_mib: int, mode: str, pinned_cores: bool):
    all_cpus = os.sched_getaffinity(0)
    if pinned_cores:
        pin_to_share()
    ctx_a, ctx_b = bn.Context(local), bn.Context(local)
    for c in (ctx_a, ctx_b):
        c.set_chunk_bytes(chunk_mib << 20)
    asc = dv.synth_ascii(0x5EEDB17C0DE5, 0, 0, N_BASES, device=dev)
    h_seq = ctx_a.pinned_empty(N_BASES, np.uint8)
    h_seq[:] = asc.cpu().numpy()
    del asc
    nw = dv.words_for(N_BASES)
    h_words = [ctx_a.pinned_empty(nw, np.uint64) for _ in range(2)]
    h_back = ctx_b.pinned_empty(N_BASES, np.uint8)
    for c in (ctx_a, ctx_b):
        bn.encode_np(h_seq, c, out=h_words[0])
        bn.decode_np(h_words[0], N_BASES, c, out=h_back)
    bn.encode_np(h_seq, ctx_a, out=h_words[1])
    steps = 6
    barrier()
    t0 = time.perf_counter()
    if mode == "serial":
        for _ in range(steps):
            bn.encode_np(h_seq, ctx_a, out=h_words[0])
            bn.decode_np(h_words[0], N_BASES, ctx_a, out=h_back)
    elif mode == "staggered":  # odd ranks run decode-then-encode: their downloads meet the even ranks' uploads
        for _ in range(steps):
            if local % 2 == 0:
                bn.encode_np(h_seq, ctx_a, out=h_words[0])
                bn.decode_np(h_words[0], N_BASES, ctx_a, out=h_back)
            else:
                bn.decode_np(h_words[1], N_BASES, ctx_a, out=h_back)
                bn.encode_np(h_seq, ctx_a, out=h_words[1])
    else:  # pipelined: thread A encodes step i+1 while thread B decodes step i
        ready = [threading.Semaphore(0), threading.Semaphore(0)]
        free = [threading.Semaphore(1), threading.Semaphore(1)]

        def enc():
            torch.cuda.set_device(local)
            for i in range(steps):
                free[i % 2].acquire()
                bn.encode_np(h_seq, ctx_a, out=h_words[i % 2])
                ready[i % 2].release()

        def dec():
            torch.cuda.set_device(local)
            for i in range(steps):
                ready[i % 2].acquire()
                bn.decode_np(h_words[i % 2], N_BASES, ctx_b, out=h_back)
                free[i % 2].release()

        th = [threading.Thread(target=enc), threading.Thread(target=dec)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    torch.cuda.synchronize()
    dt = max_over_ranks((time.perf_counter() - t0) / steps)
    ok = bool(np.array_equal(h_back[:1 << 20], h_seq[:1 << 20]))
    say(probe="e2e", mode=mode, chunk_mib=chunk_mib, pinned_cores=pinned_cores, ms_per_step=dt * 1e3,
        gbases_s=2 * N_BASES * world / dt / 1e9, pcie_gbs_aggregate=2 * UP * world / dt / 1e9, ok=ok)
    os.sched_setaffinity(0, all_cpus)
    del h_seq, h_words, h_back
    ctx_a.close()
    ctx_b.close()


if rank == 0:
    print(json.dumps({"host_threads": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}), flush=True)
if os.environ.get('PROBE_RAW', '1') == '1':
    raw_copies()
which = os.environ.get("PROBE_E2E", "full")
if which.startswith("sweep"):   # PROBE_E2E=sweep:8,16,32 -> the pipelined and serial schedules at those chunk sizes
    for mib in [int(x) for x in which.split(":")[1].split(",")]:
        e2e(mib, "pipelined", False)
        e2e(mib, "serial", False)
elif which != "none":
    e2e(64, "serial", False)
    e2e(64, "pipelined", False)
    e2e(64, "staggered", False)
    if which == "full":
        e2e(64, "serial", True)
        e2e(64, "pipelined", True)
        e2e(16, "serial", False)
        e2e(256, "serial", False)
        e2e(256, "pipelined", False)
        e2e(256, "staggered", False)
if world > 1:
    dist.destroy_process_group()
