#!/usr/bin/env python
"""bench.py -- the headline benchmark of the bitnuc hot path on B200.

Workload (BASELINE.json configs[1]): encode + decode of one contiguous 1 Gbase random sequence,
device-resident, per GPU.  A "step" encodes the sequence and decodes it back: every base goes through
the encode kernel once and the decode kernel once, so a step codes 2 x n_bases bases and
``value`` = 2 x n_bases x n_gpus / step time, in Gbases/s.  At N > 1 every rank owns its own
1 Gbase shard of an N-Gbase stream (contiguous range sharding on 64-base boundaries, no data-path
collective) -> "scaling": "weak".  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Inputs (1 GB ASCII, 250 MB packed) are larger than the
126 MB L2, so no flush is needed between iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bases B]

``--impl reference`` times the reference's own CPU algorithm on the host cores: the reference is a
Rust crate and there is no Rust toolchain in this image, so the arm runs the oracle's C/AVX2
restatement of it (oracle/bitnuc_oracle.c, kind "port"), chunked over all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "encode+decode Gbases/s (device-timed; each base encoded once and decoded once per step)"
UNIT = "Gbases/s"
BYTES_PER_BASE = 1.25  # encode: 1 B read + 0.25 B written; decode: 0.25 B read + 1 B written
SEED = 0x5EEDB17C0DE5
SAMPLE_BASES = 1 << 28  # bounded CPU sample (reference arm / cpu_baseline)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_codec_gbases(sample_bases: int, threads: int, reps: int):
    """Times the oracle's AVX2 restatement of the reference (encode + decode) on the host cores."""
    import oracle
    from oracle import oracle_np as onp
    seq = onp.synth_ascii(SEED, 0, sample_bases)
    path = oracle.PATH_AVX2 if oracle.have_avx2() else oracle.PATH_NAIVE
    t = oracle.CodecBench(seq, threads).run(reps=reps, path=path)
    if t <= 0:
        raise RuntimeError("oracle bench failed")
    return 2 * sample_bases / t / 1e9, ("avx2" if path == oracle.PATH_AVX2 else "scalar")


def run_reference(args):
    """The reference arm: the reference's CPU implementation (C/AVX2 restatement) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    try:
        oracle.build(native=True, force=True)  # the analogue of -C target-cpu=native on this box
    except Exception:
        oracle.build()
    threads = host_threads()
    sample = min(args.bases, SAMPLE_BASES)
    from oracle import oracle_np as onp
    seq = onp.synth_ascii(SEED, 0, sample)
    path = oracle.PATH_AVX2 if oracle.have_avx2() else oracle.PATH_NAIVE
    cb = oracle.CodecBench(seq, threads)
    for _ in range(max(1, args.warmup)):
        cb.run(1, path)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if cb.run(1, path) <= 0:
            raise RuntimeError("oracle bench failed")
    dt = time.perf_counter() - t0
    value = 2 * sample * args.steps / dt / 1e9
    single, _ = cpu_codec_gbases(min(sample, 1 << 26), 1, 2)
    line = {
        "impl": "reference", "metric": METRIC, "timing": "host wall-clock around the CPU implementation (there is no device in this arm)",
        "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: encode + decode of one contiguous random sequence, device-resident",
                   "sample": f"CPU arm: a bounded sample of {sample} bases per step of the 1e9-base workload", "bases_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} bases/step, {args.steps} steps, oracle C restatement of packing/avx.rs + "
                                   f"unpacking/avx.rs ({'AVX2' if path == oracle.PATH_AVX2 else 'scalar'}), chunked over {threads} threads; "
                                   f"single thread: {single:.3f} Gbases/s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Point fd 1 at stderr for the rest of the process and keep the real stdout for the one JSON line: libraries
    (NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment) cannot interleave with it."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_ours(args):
    claim_stdout()
    import numpy as np
    import torch
    import torch.distributed as dist

    import bitnuc_b200 as bn
    from bitnuc_b200 import device as dv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.bases
    # rank r owns bases [r*n, (r+1)*n) of stream 0: contiguous shards on 64-base boundaries
    first_base = rank * ((n + 63) // 64 * 64)
    asc = dv.synth_ascii(SEED, 0, first_base, n, device=dev)
    words = torch.empty(dv.words_for(n), dtype=torch.int64, device=dev)
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    status = dv.Status(dev)

    def step():
        dv.encode(asc, out=words, status=status)
        dv.decode(words, n, out=back)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    status.check()
    # parity inside the bench: the encode of a generated stream is the word stream itself
    expect = dv.synth_words(SEED, 0, first_base // 32, dv.words_for(n), device=dev)
    if n % 32:
        expect[-1] &= (1 << (2 * (n % 32))) - 1
    if not (torch.equal(words, expect) and torch.equal(back, asc)):
        raise SystemExit("bench.py: encode/decode output is wrong; refusing to report a number")
    del expect

    K = args.steps
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: exactly K steps between two events, barrier + synchronize on both sides
    barrier()
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(K):
        dv.encode(asc, out=words, status=status)
        dv.decode(words, n, out=back)
    end.record()
    barrier()
    total_ms = start.elapsed_time(end)
    # ---- same K steps again with an event around every launch: per-kernel durations for the roofline
    # (an event record between two kernels costs ~3 us, so this pass is kept out of the headline)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(K):
        ev[i][0].record()
        dv.encode(asc, out=words, status=status)
        ev[i][1].record()
        dv.decode(words, n, out=back)
        ev[i][2].record()
    barrier()
    # The timed region lasts a few milliseconds, less than one nvidia-smi sampling period: keep the same step running
    # (untimed) for ~0.6 s so that the clock samples are taken under exactly this load.
    t_load = time.time()
    while time.time() - t_load < 0.6:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if clocks:
        clocks["window"] = "the K timed steps, the per-kernel pass and ~0.6 s of the same step repeated untimed"
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    status.check()

    # ---- end to end through the public host API: pinned host buffers, H2D + kernels + D2H timed
    e2e_steps, e2e_serial_s, e2e_s, e2e_pageable_s = 0, float("nan"), float("nan"), float("nan")
    if not args.skip_e2e:
        e2e_steps, e2e_serial_s, e2e_s, e2e_pageable_s = run_e2e(bn, dv, np, torch, barrier, asc, n, local, K)

    times = torch.tensor([total_ms, enc_ms, dec_ms, e2e_s * 1e3, e2e_serial_s * 1e3, e2e_pageable_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms, e2e_ms, e2e_serial_ms, e2e_pageable_ms = times.tolist()
    report(args, world, rank, n, K, total_ms, enc_ms, dec_ms, e2e_ms, e2e_serial_ms, e2e_pageable_ms, e2e_steps, clocks, dv)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(bn, dv, np, torch, barrier, asc, n, local, K):
    """End to end through the host-pointer API (bn_encode / bn_decode) on pinned host buffers.

    Two legs, both with every H2D and D2H copy inside the timed region:
      serial    -- one host thread: encode(step i) then decode(step i);
      pipelined -- the streaming form a caller with a queue of sequences uses: host thread A encodes step i+1
                   while host thread B decodes step i (one bn_ctx per thread, as include/bitnuc_cuda.h asks for
                   concurrency), so the encode's upload and the decode's download share the full-duplex PCIe link.
    Returns (steps, serial seconds per step, pipelined seconds per step)."""
    print(f"[bench] rank-local e2e leg: {host_threads()} host threads visible", file=sys.stderr)
    ctx_a, ctx_b = bn.Context(local), bn.Context(local)
    h_seq = ctx_a.pinned_empty(n, np.uint8)
    h_words = [ctx_a.pinned_empty(dv.words_for(n), np.uint64) for _ in range(2)]
    h_back = ctx_b.pinned_empty(n, np.uint8)
    h_seq[:] = asc.cpu().numpy()
    e2e_steps = max(1, min(K, 10))   # enough steps for the encode-ahead pipeline to amortise its fill and drain
    for ctx in (ctx_a, ctx_b):  # warm-up: allocates the staging buffers of both contexts
        bn.encode_np(h_seq, ctx, out=h_words[0])
        bn.decode_np(h_words[0], n, ctx, out=h_back)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        bn.encode_np(h_seq, ctx_a, out=h_words[0])
        bn.decode_np(h_words[0], n, ctx_a, out=h_back)
    torch.cuda.synchronize()
    serial_s = (time.perf_counter() - t0) / e2e_steps
    if not np.array_equal(h_back, h_seq):
        raise SystemExit("bench.py: end-to-end round trip is wrong")

    h_back[:] = 0
    ready = [threading.Semaphore(0), threading.Semaphore(0)]
    free = [threading.Semaphore(1), threading.Semaphore(1)]
    errors = []

    def encoder():
        try:
            torch.cuda.set_device(local)
            for i in range(e2e_steps):
                free[i % 2].acquire()
                bn.encode_np(h_seq, ctx_a, out=h_words[i % 2])
                ready[i % 2].release()
        except BaseException as ex:  # surfaced below: a failed leg must not report a number
            errors.append(ex)
            for s in ready:
                s.release()

    def decoder():
        try:
            torch.cuda.set_device(local)
            for i in range(e2e_steps):
                ready[i % 2].acquire()
                bn.decode_np(h_words[i % 2], n, ctx_b, out=h_back)
                free[i % 2].release()
        except BaseException as ex:
            errors.append(ex)
            for s in free:
                s.release()

    barrier()
    threads = [threading.Thread(target=encoder), threading.Thread(target=decoder)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    piped_s = (time.perf_counter() - t0) / e2e_steps
    if errors or not np.array_equal(h_back, h_seq):
        raise SystemExit(f"bench.py: pipelined end-to-end round trip is wrong {errors[:1]}")
    # informational third leg: PAGEABLE host buffers (what a drop-in caller holding a plain Vec / ndarray passes);
    # the library bounces them through its pinned stage buffers with a multi-threaded memcpy
    p_seq, p_words, p_back = np.array(h_seq), np.empty(dv.words_for(n), dtype=np.uint64), np.empty(n, dtype=np.uint8)
    bn.encode_np(p_seq, ctx_a, out=p_words)
    bn.decode_np(p_words, n, ctx_a, out=p_back)
    t0 = time.perf_counter()
    for _ in range(2):
        bn.encode_np(p_seq, ctx_a, out=p_words)
        bn.decode_np(p_words, n, ctx_a, out=p_back)
    pageable_s = (time.perf_counter() - t0) / 2
    if not np.array_equal(p_back, h_seq):
        raise SystemExit("bench.py: pageable end-to-end round trip is wrong")
    return e2e_steps, serial_s, piped_s, pageable_s


def report(args, world, rank, n, K, total_ms, enc_ms, dec_ms, e2e_ms, e2e_serial_ms, e2e_pageable_ms, e2e_steps, clocks, dv):
    if rank == 0:
        ms_per_step = total_ms / K
        value = 2.0 * n * world / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        dom = "encode_kernel" if enc_ms >= dec_ms else "decode_kernel"
        dom_ms = max(enc_ms, dec_ms)
        achieved = BYTES_PER_BASE * n / (dom_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: encode + decode of one contiguous random sequence, device-resident",
                       "bases_per_gpu": n, "sharding": "contiguous base ranges on 64-base boundaries, no data-path collective",
                       "l2": "inputs larger than L2 (1 GB ASCII + 0.25 GB packed per GPU vs 126 MB), no flush between iterations",
                       "generator": "splitmix64 counter stream 0 (SURVEY.md 8d)"},
            "kernels": {"timing": "second pass of the same K steps with a CUDA event around every launch (same stream)",
                        "encode_ms": enc_ms, "decode_ms": dec_ms,
                        "encode_gbases_s": n / (enc_ms * 1e-3) / 1e9, "decode_gbases_s": n / (dec_ms * 1e-3) / 1e9,
                        "encode_gbs": BYTES_PER_BASE * n / (enc_ms * 1e-3) / 1e9, "decode_gbs": BYTES_PER_BASE * n / (dec_ms * 1e-3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": BYTES_PER_BASE * n,
                         "traffic": (traffic or {}).get(dom) if traffic else None},
            "e2e": {"value": 2.0 * n * world / (min(e2e_ms, e2e_serial_ms) * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": n + dv.words_for(n) * 8, "d2h_bytes_per_step": dv.words_for(n) * 8 + n,
                    "ms_per_step": min(e2e_ms, e2e_serial_ms), "steps": e2e_steps,
                    "mode": "pipelined" if e2e_ms <= e2e_serial_ms else "serial",
                    "pipelined_value": 2.0 * n * world / (e2e_ms * 1e-3) / 1e9, "pipelined_ms_per_step": e2e_ms,
                    "serial_value": 2.0 * n * world / (e2e_serial_ms * 1e-3) / 1e9, "serial_ms_per_step": e2e_serial_ms,
                    "pageable_value": 2.0 * n * world / (e2e_pageable_ms * 1e-3) / 1e9,
                    "api": "bitnuc_b200.encode_np + decode_np (bn_encode/bn_decode, pinned host buffers, chunked 3-stage pipeline "
                           "inside each call), both legs measured with every H2D/D2H copy inside the timed region. pipelined: two "
                           "host threads, one bn_ctx each -- encode of step i+1 overlaps decode of step i over the full-duplex PCIe "
                           "link; serial: one thread, encode then decode. value = the faster leg (the serial one wins when the "
                           "host's aggregate PCIe path is already saturated, e.g. 4+ GPUs on this box). pageable_value (informational): "
                           "the serial leg on pageable numpy buffers, bounced through pinned stage buffers by the library"},
            "gpu_launches": 2 * K,
            "clocks": clocks,
        }
        if args.skip_e2e:
            line["e2e"] = None
        if world == 1 and not args.skip_cpu:
            try:
                import oracle
                try:
                    oracle.build(native=True, force=True)
                except Exception:
                    oracle.build()
                threads = host_threads()
                v_all, isa = cpu_codec_gbases(SAMPLE_BASES, threads, 3)
                v_one, _ = cpu_codec_gbases(1 << 26, 1, 3)
                v_cfg0, _ = cpu_codec_gbases(1_000_000, 1, 5)   # BASELINE.json configs[0]: the reference's own CPU-runnable case
                line["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port",
                                        "single_thread_value": v_one,
                                        "configs0_1e6_bases_single_thread_value": v_cfg0,
                                        "sample": f"{SAMPLE_BASES} bases encode+decode, best of 3, C restatement of the reference's "
                                                  f"{isa} path (oracle/bitnuc_oracle.c), chunked over {threads} host threads"}
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"error": str(ex)}
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=1_000_000_000, help="bases per GPU (BASELINE configs[1]: 1e9)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
