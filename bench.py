#!/usr/bin/env python
"""bench.py -- the headline benchmark of the bitnuc hot path on B200.

Headline workload (BASELINE.json configs[1]): encode + decode of one contiguous 1 Gbase random sequence,
device-resident, per GPU.  A "step" encodes the sequence and decodes it back: every base goes through
the encode kernel once and the decode kernel once, so a step codes 2 x n_bases bases and
``value`` = 2 x n_bases x n_gpus / step time, in Gbases/s.  At N > 1 every rank owns its own
1 Gbase shard of an N-Gbase stream (contiguous range sharding on 64-base boundaries, no data-path
collective) -> "scaling": "weak".  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Inputs (1 GB ASCII, 250 MB packed) are larger than the
126 MB L2, so no flush is needed between iterations.

The same JSON line carries, beside the headline (all measured in this run, all max over ranks):
  strong   -- the strict strong-scaling figure: ONE 1 Gbase sequence cut over the N GPUs;
  e2e      -- the round trip through the host-pointer C ABI on pinned host buffers, every H2D / D2H copy timed,
              next to the PCIe / host-memory ceiling of this box measured with all N ranks copying at once;
  configs  -- BASELINE.json configs[2] (2^28 31-mers as_2bit / from_2bit), configs[3] (hdist over 2^30 pairs,
              base_counts / gc on 10 M x 150 bp reads with the NCCL all-reduce of the counters inside the timed
              region) and configs[4] (variable-length read batch, 4 Gbases per GPU = 32 Gbases on 8 GPUs, with injected
              N bases, device-resident and end to end), each with a recomputable roofline and its CPU baseline;
  multi    -- the single-process N-device layer of the C ABI (bn_multi_*): the sharded calls a C / Rust caller makes,
              with the library's own collective (NCCL or the NVLink mailbox all-reduce) inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bases B]

``--impl reference`` times the reference's own CPU algorithm on the host cores: the reference is a
Rust crate and there is no Rust toolchain in this image, so the arm runs the oracle's C/AVX2
restatement of it (oracle/bitnuc_oracle.c, kind "port"), chunked over all host threads, on the same
1e9-base sequence.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
M62 = (1 << 62) - 1


def workload_config(n: int) -> dict:
    """The ``config`` of the line -- the same object in both arms (the CPU arm's sample is described in cpu_baseline)."""
    return {"workload": "BASELINE.json configs[1]: encode + decode of one contiguous random sequence, device-resident",
            "bases_per_gpu": n, "sharding": "contiguous base ranges on 64-base boundaries, no data-path collective",
            "l2": "inputs larger than L2 (1 GB ASCII + 0.25 GB packed per GPU vs 126 MB), no flush between iterations",
            "generator": "splitmix64 counter stream 0 (SURVEY.md 8d)"}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the kernels from the committed ncu captures (profiles/roofline_traffic.json): a constant
    read from the repo, not measured in this run."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


def roofline(kernel: str, algorithmic_bytes: float, ms: float, traffic_key: str | None = None) -> dict:
    peak, src = measured_peak()
    achieved = algorithmic_bytes / (ms * 1e-3) / 1e9
    rec = (recorded_traffic() or {}).get(traffic_key or kernel)
    t = None
    if isinstance(rec, dict) and rec.get("algorithmic") and abs(algorithmic_bytes - rec["algorithmic"]) <= 0.01 * rec["algorithmic"]:
        t = rec["traffic"]   # the committed capture is of a launch of this size
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": src, "frac_of_nominal_8000": achieved / 8000.0, "algorithmic_bytes_per_launch": algorithmic_bytes,
            "ms_per_launch": ms, "traffic": t,
            "traffic_source": ("profiles/roofline_traffic.json: dram bytes of one launch of this size from the ncu --set full capture committed "
                               "with the repo -- a constant of the repo, NOT measured in this run") if t else None}


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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ====================================================================================== the reference arm (CPU)
def build_oracle():
    import oracle
    try:
        oracle.build(native=True, force=True)  # the analogue of -C target-cpu=native on this box
    except Exception:
        oracle.build()
    return oracle


def run_reference(args):
    """The reference arm: the reference's CPU implementation (C/AVX2 restatement) on all host threads, on the same
    1e9-base sequence as the GPU arm; each step = encode of the whole sequence + decode of the whole sequence."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = build_oracle()
    import numpy as np
    from oracle import baselines as B
    threads = host_threads()
    n = args.bases
    path = oracle.PATH_AVX2 if oracle.have_avx2() else oracle.PATH_NAIVE
    isa = "AVX2" if path == oracle.PATH_AVX2 else "scalar"
    seq = B.synth_ascii_mt(0, 0, n)
    words = np.ones((n + 31) // 32 + 8, dtype=np.uint64)
    back = np.ones(n + 32, dtype=np.uint8)

    def step():
        te, _ = oracle.bench_op(oracle.OP_ENCODE, n, in0=seq, out0=words, path=path, threads=threads, reps=1)
        td, _ = oracle.bench_op(oracle.OP_DECODE, n, in0=words, out0=back, path=path, threads=threads, reps=1)
        return te[0] + td[0]

    for _ in range(max(1, args.warmup)):
        step()
    times = [step() for _ in range(args.steps)]
    if not np.array_equal(back[:n], seq):
        raise SystemExit("bench.py --impl reference: round trip is wrong")
    dt = sum(times)
    value = 2 * n * args.steps / dt / 1e9
    single = B.Suite(threads=1, reps=3).codec(n_single=1 << 26, seq=seq[: 1 << 26], grid=[(path, 1)])[0]
    line = {
        "impl": "reference", "metric": METRIC,
        "timing": "host clock around the CPU implementation (threads created and pinned before the clock starts; there is no device in this arm)",
        "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "best_step": 2 * n / min(times) / 1e9, "worst_step": 2 * n / max(times) / 1e9,
                         "median_step": 2 * n / statistics.median(times) / 1e9,
                         "single_thread_value": single["value"],
                         "sample": f"the whole workload: {n} bases encoded and decoded per step, {args.steps} steps; oracle C restatement of "
                                   f"packing/avx.rs + unpacking/avx.rs ({isa}, -march=native), chunked on 128-base boundaries over "
                                   f"{threads} pinned threads (the chunking is the harness's: the reference is single-threaded)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ====================================================================================== our arm
class Job:
    """Per-process state of our arm: device, ranks, barriers, reductions."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        self.np, self.torch, self.dist, self.args = np, torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.cpu_group = dist.new_group(backend="gloo")  # host-side waits that leave the GPUs idle (multi leg)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def rmax(self, *vals):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def rsum_int(self, *vals):
        t = self.torch.tensor([int(v) for v in vals], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]

    def ev_ms(self, fn, reps: int, warm: int = 3) -> float:
        """Device time of one call of fn: CUDA events on the launching stream around `reps` calls, after `warm` untimed ones,
        barrier + synchronize on both sides; the caller takes the max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) / reps

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def leg_codec(job: Job, n: int, first_base: int, K: int, warm: int, sample_clocks: bool):
    """K steps of encode + decode of this rank's n bases (device-resident).  Returns local times and the buffers."""
    torch = job.torch
    from bitnuc_b200 import device as dv
    dev = job.dev
    asc = dv.synth_ascii(SEED, 0, first_base, n, device=dev)
    words = torch.empty(dv.words_for(n), dtype=torch.int64, device=dev)
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    status = dv.Status(dev)

    def step():
        dv.encode(asc, out=words, status=status)
        dv.decode(words, n, out=back)

    for _ in range(warm):
        step()
    job.barrier()
    status.check()
    # self-consistency inside the bench (the oracle comparison is tests/ and smoke()): the encode of a generated stream is the
    # generator's word stream itself (bn_synth_words_dev, pinned to the oracle's generator in tests/test_gpu_parity.py)
    expect = dv.synth_words(SEED, 0, first_base // 32, dv.words_for(n), device=dev)
    if n % 32:
        expect[-1] &= (1 << (2 * (n % 32))) - 1
    if not (torch.equal(words, expect) and torch.equal(back, asc)):
        raise SystemExit("bench.py: encode/decode output is wrong; refusing to report a number")
    del expect
    sampler = ClockSampler(job.local) if (sample_clocks and job.rank == 0) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: exactly K steps between two events, barrier + synchronize on both sides
    job.barrier()
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(K):
        dv.encode(asc, out=words, status=status)
        dv.decode(words, n, out=back)
    end.record()
    job.barrier()
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
    job.barrier()
    clocks = None
    if sample_clocks:
        # The timed region lasts a few milliseconds, less than one nvidia-smi sampling period: keep the same step running
        # (untimed) for ~0.6 s so that the clock samples are taken under exactly this load.
        t_load = time.time()
        while time.time() - t_load < 0.6:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        job.barrier()
        clocks = sampler.stop(t_wall0, time.time()) if sampler else None
        if clocks:
            clocks["window"] = "the K timed steps, the per-kernel pass and ~0.6 s of the same step repeated untimed"
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    status.check()
    return {"total_ms": total_ms, "enc_ms": enc_ms, "dec_ms": dec_ms, "clocks": clocks, "asc": asc}


def cudart():
    """libcudart through ctypes (cudaMemcpyAsync on raw pointers for the copy probes): the copy torch has already loaded is found by
    its soname; otherwise the wheels' and the toolkit's copies are tried."""
    import glob
    import site
    cands = ["libcudart.so.12", "libcudart.so"]
    for base in site.getsitepackages() + ["/usr/local/cuda/lib64"]:
        cands += sorted(glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib", "libcudart.so*"))) + sorted(glob.glob(os.path.join(base, "libcudart.so*")))
    for c in cands:
        try:
            rt = C.CDLL(c)
            rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
            rt.cudaMemcpyAsync.restype = C.c_int
            return rt
        except OSError:
            continue
    raise SystemExit("bench.py: libcudart not found for the PCIe probes")


def leg_pcie_ceiling(job: Job, h_up, h_dn, nbytes: int):
    """The denominator of e2e: raw pinned copies of `nbytes` each way on ALL ranks at once -- upload only, download only,
    both directions (separate streams).  Aggregate GB/s over the N ranks, max-over-ranks time, best of 3."""
    torch = job.torch
    rt = cudart()
    d_up = torch.empty(nbytes, dtype=torch.uint8, device=job.dev)
    d_dn = torch.ones(nbytes, dtype=torch.uint8, device=job.dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    p_up, p_dn = h_up.ctypes.data, h_dn.ctypes.data

    def run(up, dn):
        best = float("inf")
        for _ in range(4):
            job.barrier()
            t0 = time.perf_counter()
            rc = 0
            if up:
                rc |= rt.cudaMemcpyAsync(d_up.data_ptr(), p_up, nbytes, 1, s_up.cuda_stream)
            if dn:
                rc |= rt.cudaMemcpyAsync(p_dn, d_dn.data_ptr(), nbytes, 2, s_dn.cuda_stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rc:
                raise SystemExit("bench.py: cudaMemcpyAsync failed in the PCIe probe")
            best = min(best, job.rmax(dt)[0])
        return best

    t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
    w = job.world
    return {"h2d": w * nbytes / t_up / 1e9, "d2h": w * nbytes / t_dn / 1e9, "both": 2 * w * nbytes / t_both / 1e9,
            "bytes_each_way_per_gpu": nbytes, "unit": "GB/s aggregate over the N GPUs",
            "how": "one cudaMemcpyAsync per direction per rank from / to the pinned buffers of the e2e leg, all ranks at once, "
                   "wall clock between barriers, max over ranks, best of 4"}


def leg_pcie_pattern(job: Job, h_big_up, h_small_dn, h_small_up, h_big_dn, n: int, steps: int):
    """The same denominator, for the COPY PATTERN of the pipelined schedule rather than for two monolithic copies: thread A
    moves the chunks of an encode (64 MiB up, 16 MiB down per chunk), thread B those of a decode (16 MiB up, 64 MiB down),
    each through three rotating streams exactly as bn_encode / bn_decode issue them -- but with NO kernel in between.  What
    this gives is all the link offers the pipelined round trip (tools/pcie_pattern.cu is the C++ form of this probe)."""
    torch = job.torch
    rt = cudart()
    chunk = 64 << 20
    n_chunks = (n + chunk - 1) // chunk
    d_a = torch.empty(3 * chunk + 3 * (chunk // 4), dtype=torch.uint8, device=job.dev)
    d_b = torch.empty(3 * chunk + 3 * (chunk // 4), dtype=torch.uint8, device=job.dev)

    def call(d, h_up, h_dn, up_unit, dn_unit, streams, events):
        # chunk c: up_unit bytes per base-chunk up, dn_unit down (1 or 1/4 byte per base)
        for c in range(n_chunks):
            k = c % 3
            bases = min(chunk, n - c * chunk)
            if c >= 3:
                events[k].synchronize()
            up, dn = int(bases * up_unit) & ~7, int(bases * dn_unit) & ~7
            big, small = d.data_ptr() + k * chunk, d.data_ptr() + 3 * chunk + k * (chunk // 4)   # the stage's two device buffers
            rt.cudaMemcpyAsync(big if up_unit == 1.0 else small, h_up.ctypes.data + int(c * chunk * up_unit), up, 1, streams[k].cuda_stream)
            rt.cudaMemcpyAsync(h_dn.ctypes.data + int(c * chunk * dn_unit), big if dn_unit == 1.0 else small, dn, 2, streams[k].cuda_stream)
            events[k].record(streams[k])
        for k in range(min(3, n_chunks)):
            events[k].synchronize()

    sa, sb = [torch.cuda.Stream() for _ in range(3)], [torch.cuda.Stream() for _ in range(3)]
    ea, eb = [torch.cuda.Event(blocking=True) for _ in range(3)], [torch.cuda.Event(blocking=True) for _ in range(3)]

    def run():
        def a():
            torch.cuda.set_device(job.local)
            for _ in range(steps):
                call(d_a, h_big_up, h_small_dn.view(job.np.uint8), 1.0, 0.25, sa, ea)

        def b():
            torch.cuda.set_device(job.local)
            for _ in range(steps):
                call(d_b, h_small_up.view(job.np.uint8), h_big_dn, 0.25, 1.0, sb, eb)

        job.barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=a), threading.Thread(target=b)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    run()
    (t,) = job.rmax(run())
    return t


def leg_e2e(job: Job, asc, n: int, K: int):
    """End to end through the host-pointer API (bn_encode / bn_decode) on pinned host buffers, every H2D and D2H copy inside
    the timed region.  Three schedules of the same calls:
      pipelined -- THE e2e figure: the streaming form a caller with a queue of sequences uses; host thread A encodes step
                   i+1 while host thread B decodes step i (one bn_ctx per thread), so uploads and downloads share the
                   full-duplex link;
      serial    -- one host thread: encode(step i) then decode(step i);
      staggered -- serial, with odd ranks running decode-then-encode so their downloads meet the even ranks' uploads."""
    np, torch = job.np, job.torch
    import bitnuc_b200 as bn
    from bitnuc_b200 import device as dv
    local = job.local
    ctx_a, ctx_b = bn.Context(local), bn.Context(local)
    nw = dv.words_for(n)
    h_seq = ctx_a.pinned_empty(n, np.uint8)
    h_words = [ctx_a.pinned_empty(nw, np.uint64) for _ in range(2)]
    h_back = ctx_b.pinned_empty(n, np.uint8)
    h_seq[:] = asc.cpu().numpy()
    steps = max(2, min(K, 32))  # the K steps of the run (bounded): the encode-ahead pipeline pays one encode alone to fill and
                                # one decode alone to drain, once per leg -- over 8 steps that was 5 % of the figure
    for ctx in (ctx_a, ctx_b):  # warm-up: allocates the staging buffers of both contexts
        bn.encode_np(h_seq, ctx, out=h_words[0])
        bn.decode_np(h_words[0], n, ctx, out=h_back)
    bn.encode_np(h_seq, ctx_a, out=h_words[1])

    pcie = leg_pcie_ceiling(job, h_seq, h_back, n)
    pcie["pattern_s"] = leg_pcie_pattern(job, h_seq, h_words[0], h_words[1], h_back, n, steps)   # h_words[0] / h_back now hold device scratch:
                                                                                             # every schedule below rewrites them first

    def wall(fn):
        job.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    def serial():
        for _ in range(steps):
            bn.encode_np(h_seq, ctx_a, out=h_words[0])
            bn.decode_np(h_words[0], n, ctx_a, out=h_back)

    def staggered():
        for _ in range(steps):
            if job.rank % 2 == 0:
                bn.encode_np(h_seq, ctx_a, out=h_words[0])
                bn.decode_np(h_words[0], n, ctx_a, out=h_back)
            else:
                bn.decode_np(h_words[1], n, ctx_a, out=h_back)
                bn.encode_np(h_seq, ctx_a, out=h_words[1])

    errors = []

    def pipelined():
        ready = [threading.Semaphore(0), threading.Semaphore(0)]
        free = [threading.Semaphore(1), threading.Semaphore(1)]

        def encoder():
            try:
                torch.cuda.set_device(local)
                for i in range(steps):
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
                for i in range(steps):
                    ready[i % 2].acquire()
                    bn.decode_np(h_words[i % 2], n, ctx_b, out=h_back)
                    free[i % 2].release()
            except BaseException as ex:
                errors.append(ex)
                for s in free:
                    s.release()

        th = [threading.Thread(target=encoder), threading.Thread(target=decoder)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    out = {}
    for name, fn in (("serial", serial), ("staggered", staggered), ("pipelined", pipelined)):
        h_back[:4096] = 0
        out[name] = wall(fn)
        if errors or not np.array_equal(h_back, h_seq):
            raise SystemExit(f"bench.py: {name} end-to-end round trip is wrong {errors[:1]}")
    # informational: PAGEABLE host buffers (what a drop-in caller holding a plain Vec / ndarray passes);
    # the library bounces them through its pinned stage buffers with a multi-threaded memcpy
    p_seq, p_words, p_back = np.array(h_seq), np.empty(nw, dtype=np.uint64), np.empty(n, dtype=np.uint8)
    bn.encode_np(p_seq, ctx_a, out=p_words)
    bn.decode_np(p_words, n, ctx_a, out=p_back)
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        bn.encode_np(p_seq, ctx_a, out=p_words)
        bn.decode_np(p_words, n, ctx_a, out=p_back)
    out["pageable"] = (time.perf_counter() - t0) / 2
    if not np.array_equal(p_back, h_seq):
        raise SystemExit("bench.py: pageable end-to-end round trip is wrong")
    out["steps"] = steps
    out["pcie"] = pcie
    del h_seq, h_words, h_back
    ctx_a.close()
    ctx_b.close()
    return out


# ------------------------------------------------------------------------------------------ BASELINE configs[2..4]
def leg_kmers(job: Job, reps: int):
    """configs[2]: batched as_2bit / from_2bit of 2^28 random 31-mers (one u64 each), tight 31-byte records, cut over the ranks
    by record index.  39 B per k-mer per kernel (31 + 8)."""
    torch = job.torch
    from bitnuc_b200 import device as dv
    from bitnuc_b200 import sharding as sh
    n_total = 1 << 28
    r0, r1 = sh.shard_range(n_total, job.rank, job.world, 64)
    n = r1 - r0
    words = dv.synth_words(SEED, 1, r0, n, device=job.dev)
    expect = words & M62
    recs = torch.zeros(n * 31, dtype=torch.uint8, device=job.dev)
    packed = torch.empty(n, dtype=torch.int64, device=job.dev)
    st = dv.Status(job.dev)
    ms_f = job.ev_ms(lambda: dv.from_2bit_batch(words, 31, 31, out=recs), reps)
    ms_a = job.ev_ms(lambda: dv.as_2bit_batch(recs, n, 31, 31, out=packed, status=st), reps)
    st.check()
    if not torch.equal(packed, expect):
        raise SystemExit("bench.py: as_2bit(from_2bit(w)) != w & (2^62 - 1)")
    # from_2bit against the definition on a strided sample of records: base i of record r = "ACGT"[(W_r >> 2i) & 3]
    idx = torch.arange(0, n, max(1, n // 4096), device=job.dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=job.dev)
    sh_ = torch.arange(31, device=job.dev, dtype=torch.int64) * 2
    want = lut[((words[idx].unsqueeze(1) >> sh_) & 3)]
    got = recs.view(n, 31)[idx]
    if not torch.equal(want, got):
        raise SystemExit("bench.py: from_2bit differs from its definition")
    ms_a, ms_f = job.rmax(ms_a, ms_f)
    del words, expect, recs, packed
    job.free()
    return {"n_total": n_total, "n_local": n, "as_2bit_ms": ms_a, "from_2bit_ms": ms_f}


def leg_hdist_counts(job: Job, reps: int):
    """configs[3]: hdist over 2^30 pairs of packed 32-mers (per-pair distances, and the whole-sequence total with its sum
    all-reduced), base_counts / gc_content on 10 M x 150 bp reads with the NCCL all-reduce of the four counters inside the
    timed region; pairs and reads cut over the ranks by index (strong scaling)."""
    np, torch, dist = job.np, job.torch, job.dist
    from bitnuc_b200 import device as dv
    from bitnuc_b200 import sharding as sh
    dev = job.dev
    n_total = 1 << 30
    r0, r1 = sh.shard_range(n_total, job.rank, job.world, 64)
    n = r1 - r0
    u, v = dv.synth_words(SEED, 2, r0, n, device=dev), dv.synth_words(SEED, 3, r0, n, device=dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    tot = torch.empty(1, dtype=torch.int64, device=dev)

    def hdist_total():
        dv.hdist(u, v, 32 * n, out=tot)
        if job.world > 1:
            dist.all_reduce(tot)

    ms_pairs = job.ev_ms(lambda: dv.hdist_pairs(u, v, 32, out=out), reps)
    ms_total = job.ev_ms(hdist_total, reps)
    dv.hdist(u, v, 32 * n, out=tot)
    local_total = int(tot.item())
    if int(out.sum(dtype=torch.int64).item()) != local_total:  # checksum of checksums
        raise SystemExit("bench.py: sum of hdist_pairs != hdist total")
    # per-pair distances against the definition (popcount of the per-base mismatch mask) on a strided sample of the whole range
    idx = torch.arange(0, n, max(1, n // 65536), device=dev)
    x = (u[idx] ^ v[idx]).cpu().numpy().view(np.uint64)
    m = (x | (x >> np.uint64(1))) & np.uint64(0x5555555555555555)
    ref = np.unpackbits(m.view(np.uint8).reshape(-1, 8), axis=1).sum(axis=1).astype(np.int32)
    if not np.array_equal(out[idx].cpu().numpy(), ref):
        raise SystemExit("bench.py: hdist_pairs differs from its definition")
    (global_total,) = job.rsum_int(local_total)
    del u, v, out
    job.free()

    reads_total = 10_000_000
    q0, q1 = sh.shard_range(reads_total, job.rank, job.world, 1)
    reads = q1 - q0
    words = dv.synth_words(SEED, 4, 5 * q0, 5 * reads, device=dev)
    words.view(reads, 5)[:, 4] &= (1 << 44) - 1  # 150 = 4*32 + 22 bases: the tail of word 4 is zero padding
    wo = torch.arange(reads, dtype=torch.int64, device=dev) * 5
    lens = torch.full((reads,), 150, dtype=torch.int64, device=dev)
    counts4 = torch.empty((reads, 4), dtype=torch.int64, device=dev)
    gcs = torch.empty(reads, dtype=torch.float64, device=dev)
    totals = torch.empty(4, dtype=torch.int64, device=dev)

    def counts_indexed():
        dv.base_counts_batch(words, wo, lens, counts4=counts4, gc=gcs, totals=totals)
        if job.world > 1:
            dist.all_reduce(totals)

    def counts_fixed():
        dv.base_counts_fixed(words, reads, 150, counts4=counts4, gc=gcs, totals=totals)
        if job.world > 1:
            dist.all_reduce(totals)

    def comm_only():
        if job.world > 1:
            dist.all_reduce(totals)

    ms_idx = job.ev_ms(counts_indexed, reps)
    ms_fix = job.ev_ms(counts_fixed, reps)
    ms_comm = job.ev_ms(comm_only, reps) if job.world > 1 else 0.0
    ms_kernel_only = job.ev_ms(lambda: dv.base_counts_fixed(words, reads, 150, counts4=counts4, gc=gcs, totals=totals), reps)
    counts_fixed()
    g_totals = [int(x) for x in totals.tolist()]
    if sum(g_totals) != 150 * reads_total or int(counts4.sum().item()) != 150 * reads:
        raise SystemExit("bench.py: base counts do not add up to the number of bases")
    gc_np = (counts4[:, 1] + counts4[:, 2]).cpu().numpy().astype(np.float64)
    if not np.array_equal(gcs.cpu().numpy(), (gc_np / np.float64(150.0)) * np.float64(100.0)):  # the reference's operation order
        raise SystemExit("bench.py: per-read gc_content differs from (gc / len) * 100")
    ms_pairs, ms_total, ms_idx, ms_fix, ms_comm, ms_kernel_only = job.rmax(ms_pairs, ms_total, ms_idx, ms_fix, ms_comm, ms_kernel_only)
    del words, wo, lens, counts4, gcs
    job.free()
    return {"pairs_total": n_total, "pairs_local": n, "pairs_ms": ms_pairs, "total_ms": ms_total, "hdist_total": global_total,
            "reads_total": reads_total, "reads_local": reads, "counts_indexed_ms": ms_idx, "counts_fixed_ms": ms_fix,
            "comm_ms": ms_comm, "counts_fixed_kernel_only_ms": ms_kernel_only, "totals": g_totals,
            "gc_global": sh.gc_from_counts(g_totals)}


def leg_batch(job: Job, reps: int, bases_per_gpu: int):
    """configs[4]: variable-length read batch (50 bp - 10 kbp, offset-indexed), `bases_per_gpu` per GPU (4e9 -> 32 Gbases on
    8 GPUs), cut over the ranks by byte volume on read boundaries; device-resident kernel time, end to end through
    bn_encode_batch on pinned host buffers, and a run with injected N bases for InvalidBase parity."""
    np, torch = job.np, job.torch
    import bitnuc_b200 as bn
    from bitnuc_b200 import device as dv
    from bitnuc_b200 import sharding as sh
    from bitnuc_b200 import synth
    from bitnuc_b200._lib import BnError
    dev = job.dev
    lens = synth.cfg5_read_lengths(bases_per_gpu * job.world, SEED)
    n_all = lens.size
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    r_lo, r_hi = sh.shard_reads_by_volume(offsets, job.world)[job.rank]
    b_lo, b_hi = int(offsets[r_lo]), int(offsets[r_hi])
    n_reads, n_bytes = r_hi - r_lo, b_hi - b_lo
    a_lo = b_lo // 32 * 32
    d_all = dv.synth_ascii(SEED, 5, a_lo, b_hi - a_lo, device=dev)
    d_bytes = d_all[b_lo - a_lo:]
    if d_bytes.data_ptr() % 16:
        d_bytes = d_bytes.clone()
    del d_all
    rel = (offsets[r_lo: r_hi + 1] - np.uint64(b_lo)).astype(np.uint64)
    d_off = torch.from_numpy(rel.view(np.int64)).to(dev)
    my_lens = lens[r_lo:r_hi]
    n_words = int(((my_lens + np.uint64(31)) // np.uint64(32)).sum())

    # ---- device-resident
    ctx = dv.api.default_context(job.local)
    words = torch.empty(n_bytes // 32 + n_reads, dtype=torch.int64, device=dev)
    wo = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
    rs = torch.empty(n_reads, dtype=torch.int32, device=dev)
    scratch = torch.empty(ctx.lib.bn_encode_batch_scratch_bytes(n_reads, n_bytes), dtype=torch.uint8, device=dev)
    st = dv.Status(dev)

    def run_dev():
        dv.raise_for(ctx.lib.bn_encode_batch_dev(ctx.handle, dv._stream(), dv._ptr(d_bytes), dv._ptr(d_off), n_reads, n_bytes,
                                                 dv._ptr(words), dv._ptr(wo), dv._ptr(rs), dv._ptr(st.word), dv._ptr(scratch)))

    ms_dev = job.ev_ms(run_dev, max(2, reps // 2), warm=2)
    st.check()
    if int(wo[-1].item()) != n_words:
        raise SystemExit("bench.py: encode_batch wrote the wrong number of words")
    for ridx in (0, n_reads // 2, n_reads - 1):  # round trip of a few reads through decode
        w0, ln = int(wo[ridx].item()), int(my_lens[ridx])
        back = dv.decode(words[w0: w0 + (ln + 31) // 32].contiguous(), ln)
        if not torch.equal(back, d_bytes[int(rel[ridx]): int(rel[ridx]) + ln]):
            raise SystemExit("bench.py: encode_batch -> decode round trip is wrong")
    words_dev_sample = words[:4096].cpu().numpy().view(np.uint64).copy()
    del words, wo, rs, scratch

    # ---- end to end: pinned host buffers through bn_encode_batch (chunked 3-stage pipeline inside the call)
    hctx = bn.Context(job.local)
    h_bytes = hctx.pinned_empty(n_bytes, np.uint8)
    h_bytes[:] = d_bytes.cpu().numpy()
    del d_bytes, d_off
    job.free()
    h_off = hctx.pinned_empty(n_reads + 1, np.uint64)
    h_off[:] = rel
    h_words = hctx.pinned_empty(n_bytes // 32 + n_reads, np.uint64)
    h_wo = hctx.pinned_empty(n_reads + 1, np.uint64)
    h_rs = hctx.pinned_empty(n_reads, np.uint32)

    def call(with_status):
        err = BnError()
        rc = hctx.lib.bn_encode_batch(hctx.handle, h_bytes.ctypes.data_as(C.c_void_p), h_off.ctypes.data_as(C.c_void_p), n_reads,
                                      h_words.ctypes.data_as(C.c_void_p), h_wo.ctypes.data_as(C.c_void_p),
                                      h_rs.ctypes.data_as(C.c_void_p) if with_status else None, C.byref(err))
        return rc, err

    def timed(with_status, n_rep):
        call(with_status)  # warm-up: sizes the device staging buffers
        ts, rc, err = [], 0, None
        for _ in range(n_rep):
            job.barrier()
            t0 = time.perf_counter()
            rc, err = call(with_status)
            ts.append(job.rmax(time.perf_counter() - t0)[0])
        return statistics.median(ts), rc, err

    t_clean, rc, err = timed(False, 3)
    if rc != 0 or int(h_wo[n_reads]) != n_words or not np.array_equal(h_words[:4096], words_dev_sample):
        raise SystemExit(f"bench.py: bn_encode_batch failed or differs from the device-resident call (rc {rc})")
    # ---- injected N: read r gets 'N' at h2(r) mod len iff h1(r) mod 100003 == 0 (whole-batch rule, rank-local bytes)
    victims, pos = synth.cfg5_injected_n(n_all, lens)
    if victims.size == 0:
        victims, pos = np.array([n_all // 3]), np.array([int(lens[n_all // 3]) // 2], dtype=np.int64)
    mine = (victims >= r_lo) & (victims < r_hi)
    h_bytes[(offsets[victims[mine]] - np.uint64(b_lo)).astype(np.int64) + pos[mine]] = ord("N")
    t_inj, rc, err = timed(True, 3)
    local_key = (int(err.offset) << 8 | int(err.base)) if rc == 1 else None
    if (rc == 1) != bool(mine.any()):
        raise SystemExit("bench.py: InvalidBase reported on the wrong rank")
    first = sh.first_error_across_ranks(local_key, b_lo, device=dev)
    expect_off = int(offsets[victims[0]]) + int(pos[0])
    bad = np.flatnonzero(h_rs[:n_reads] != 0xFFFFFFFF)
    if first != (expect_off, ord("N")) or not np.array_equal(bad + r_lo, victims[mine]) or \
            not np.array_equal(h_rs[bad].astype(np.int64), pos[mine]):
        raise SystemExit(f"bench.py: injected-N parity failed: first {first}, expected offset {expect_off}")
    (ms_dev,) = job.rmax(ms_dev)
    tot_bytes, tot_words, tot_reads = job.rsum_int(n_bytes, n_words, n_reads)
    del h_bytes, h_off, h_words, h_wo, h_rs
    hctx.close()
    job.free()
    return {"reads": tot_reads, "bases": tot_bytes, "words": tot_words, "local": {"reads": n_reads, "bases": n_bytes, "words": n_words},
            "dev_ms": ms_dev, "e2e_clean_s": t_clean, "e2e_injected_s": t_inj, "injected": int(victims.size),
            "first_error": {"record": int(victims[0]), "position": int(pos[0]), "offset": expect_off, "byte": ord("N")}}


def leg_multi(job: Job, reps: int, n_bases: int):
    """The single-process N-device layer of the C ABI, run by rank 0 alone while the other ranks wait on a host-side (gloo)
    barrier with their GPUs idle: what a C / Rust caller of bn_multi_* gets from one process.
      counts  -- bn_multi_base_counts_fixed_dev on the 10 M x 150 bp reads cut over the N devices, the collective of the
                 library inside the timed region (NCCL ncclAllReduce, and the NVLink mailbox all-reduce kernel);
      hdist   -- bn_multi_hdist_dev over 2^28 word pairs cut over the devices, total all-reduced;
      e2e     -- bn_multi_encode + bn_multi_decode of ONE n_bases sequence from pinned host memory (strong scaling)."""
    np, torch = job.np, job.torch
    out = None
    job.cpu_barrier()
    if job.rank == 0:
        from bitnuc_b200 import device as dv
        from bitnuc_b200.multi import MultiContext
        n_dev = job.world
        out = {"n_devices": n_dev, "process": "one (rank 0); the other ranks idle on a gloo barrier"}
        for mode in (["nccl", "p2p"] if n_dev > 1 else ["p2p"]):
            try:
                m = MultiContext(n_dev, reduce=mode)
            except Exception as ex:
                out[mode] = {"error": str(ex)}
                continue
            reads_total = 10_000_000
            starts = m.shard_units(reads_total, 1)
            words, totals, gcg, n_reads = [], [], [], []
            for i, d in enumerate(m.devices):
                with torch.cuda.device(d):
                    dd = torch.device("cuda", d)
                    r = starts[i + 1] - starts[i]
                    w = dv.synth_words(SEED, 4, 5 * starts[i], 5 * r, device=dd)
                    w.view(r, 5)[:, 4] &= (1 << 44) - 1
                    words.append(w)
                    totals.append(torch.zeros(4, dtype=torch.int64, device=dd))
                    gcg.append(torch.zeros(1, dtype=torch.float64, device=dd))
                    n_reads.append(r)
                    torch.cuda.synchronize()
            ms = []
            for it in range(3 + reps):
                m.base_counts_fixed_dev(words, n_reads, 150, totals, gc=gcg)
                t = max(m.last_ms())
                if it >= 3:
                    ms.append(t)
            g = [int(x) for x in totals[-1].tolist()]
            if sum(g) != 150 * reads_total or any([int(x) for x in t.tolist()] != g for t in totals):
                raise SystemExit("bench.py: bn_multi base counts are wrong")
            res = {"reduce": m.reduce, "nccl_version": m.nccl_version,
                   "base_counts_fixed_ms": statistics.median(ms), "base_counts_fixed_ms_best": min(ms),
                   "totals": g, "gc_global": float(gcg[0].item()),
                   "timing": "bn_multi_last_ms: CUDA events on every device's stream, first launch to end of the collective, max over devices; median of reps"}
            del words
            # whole-sequence hdist over 2^28 word pairs
            pw = 1 << 28
            ps = m.shard_units(pw, 2)
            ua, ub, tt, nb = [], [], [], []
            for i, d in enumerate(m.devices):
                with torch.cuda.device(d):
                    dd = torch.device("cuda", d)
                    k = ps[i + 1] - ps[i]
                    ua.append(dv.synth_words(SEED, 2, ps[i], k, device=dd))
                    ub.append(dv.synth_words(SEED, 3, ps[i], k, device=dd))
                    tt.append(torch.zeros(1, dtype=torch.int64, device=dd))
                    nb.append(32 * k)
                    torch.cuda.synchronize()
            ms = []
            for it in range(3 + reps):
                m.hdist_dev(ua, ub, nb, tt)
                t = max(m.last_ms())
                if it >= 3:
                    ms.append(t)
            res["hdist_2p28_words_ms"] = statistics.median(ms)
            res["hdist_total"] = int(tt[0].item())
            if any(int(t.item()) != res["hdist_total"] for t in tt):
                raise SystemExit("bench.py: bn_multi hdist totals differ between devices")
            del ua, ub
            if mode == "p2p":  # host-pointer round trip, once (the reduce mode plays no part in it)
                ctx0 = m.contexts[0]
                h_seq = ctx0.pinned_empty(n_bases, np.uint8)
                with torch.cuda.device(m.devices[0]):
                    h_seq[:] = dv.synth_ascii(SEED, 0, 0, n_bases, device=torch.device("cuda", m.devices[0])).cpu().numpy()
                h_words = ctx0.pinned_empty(dv.words_for(n_bases), np.uint64)
                h_back = ctx0.pinned_empty(n_bases, np.uint8)
                m.encode_np(h_seq, out=h_words)
                m.decode_np(h_words, n_bases, out=h_back)
                steps = 5
                t0 = time.perf_counter()
                for _ in range(steps):
                    m.encode_np(h_seq, out=h_words)
                    m.decode_np(h_words, n_bases, out=h_back)
                dt = (time.perf_counter() - t0) / steps
                if not np.array_equal(h_back, h_seq):
                    raise SystemExit("bench.py: bn_multi round trip is wrong")
                out["e2e_strong"] = {"value": 2 * n_bases / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "bases": n_bases,
                                     "api": "bn_multi_encode + bn_multi_decode (one process, one host worker thread + bn_ctx per device), "
                                            "pinned host buffers, serial encode then decode, every copy inside the wall-clock region"}
                del h_seq, h_words, h_back
            out[mode] = res
            m.close()
            job.free()
    job.cpu_barrier()
    return out


def cpu_suite_main(args):
    """`bench.py --cpu-suite`: the cpu_baseline legs in a process of their own (no CUDA context, no torch threads -- the
    same conditions as the --impl reference arm), one JSON object on stdout."""
    oracle = build_oracle()
    from oracle import baselines as B
    s = B.Suite(reps=5)
    t0 = time.time()
    seq = B.synth_ascii_mt(0, 0, args.bases)
    path = s.paths[0]
    rows = {"codec_full": s.codec(seq=seq, grid=[(path, s.threads)]),                      # the whole workload, as the reference arm
            "codec": B.Suite(reps=5).codec(n_many=1 << 28, seq=seq[: 1 << 28]),
            "kmers": s.kmers(), "hdist": s.hdist(), "base_counts": s.base_counts(), "encode_batch": B.encode_batch(s)}
    rows["cfg0"] = B.Suite(threads=1, reps=7).codec(n_single=1_000_000, n_many=1_000_000, seq=seq[:1_000_000], grid=[(path, 1)])  # configs[0]
    rows["seconds"] = time.time() - t0
    model = next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "unknown")
    rows["host"] = {"cpu": model, "threads": s.threads, "compiler": "gcc -O3 -march=native -ffp-contract=off (oracle/Makefile, target `native`)"}
    print(json.dumps(rows))


def cpu_baselines(args):
    """cpu_baseline legs (rank 0, N = 1): every row's CPU form, AVX2 and scalar, one pinned thread and all threads, timed in a
    fresh process (`--cpu-suite`)."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-suite", "--bases", str(args.bases)], capture_output=True, text=True,
                       timeout=600)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-400:])
    rows = json.loads(r.stdout.strip().splitlines()[-1])
    log(f"cpu baselines took {rows['seconds']:.1f} s")
    return rows


def run_ours(args):
    claim_stdout()
    job = Job(args)
    np, torch = job.np, job.torch
    from bitnuc_b200 import device as dv
    from bitnuc_b200 import sharding as sh
    world, rank = job.world, job.rank
    n, K, W = args.bases, args.steps, max(args.warmup, 3)
    t_start = time.time()

    # ---- headline: weak scaling, rank r owns bases [r*n, (r+1)*n) of stream 0 (contiguous shards on 64-base boundaries)
    first_base = rank * ((n + 63) // 64 * 64)
    main = leg_codec(job, n, first_base, K, W, sample_clocks=True)
    total_ms, enc_ms, dec_ms = job.rmax(main["total_ms"], main["enc_ms"], main["dec_ms"])
    asc = main.pop("asc")
    log(f"headline done {time.time() - t_start:.1f} s")

    # ---- end to end + PCIe ceiling
    e2e = None
    if not args.skip_e2e:
        e2e = leg_e2e(job, asc, n, K)
        for k in ("serial", "staggered", "pipelined", "pageable"):
            (e2e[k],) = job.rmax(e2e[k])
        log(f"e2e done {time.time() - t_start:.1f} s")
    del asc
    job.free()

    # ---- strict strong scaling: ONE n-base sequence cut over the ranks on 64-base boundaries
    strong = None
    if not args.skip_configs:
        s0, s1 = sh.shard_bases(n, rank, world)
        st = leg_codec(job, s1 - s0, s0, K, W, sample_clocks=False)
        st.pop("asc")
        job.free()
        s_total, s_enc, s_dec = job.rmax(st["total_ms"], st["enc_ms"], st["dec_ms"])
        strong = {"value": 2.0 * n / (s_total / K * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": s_total / K, "total_bases": n,
                  "encode_ms": s_enc, "decode_ms": s_dec, "scaling": "strong",
                  "what": f"one {n}-base sequence cut over {world} GPU(s) on 64-base boundaries; same kernels, device-timed, max over ranks"}

    cfg2 = cfg3 = cfg4 = multi = None
    if not args.skip_configs:
        cfg2 = leg_kmers(job, 10)
        log(f"configs[2] done {time.time() - t_start:.1f} s")
        cfg3 = leg_hdist_counts(job, 10)
        log(f"configs[3] done {time.time() - t_start:.1f} s")
        cfg4 = leg_batch(job, 6, args.batch_bases)
        log(f"configs[4] done {time.time() - t_start:.1f} s")
        multi = leg_multi(job, 10, n)
        log(f"multi done {time.time() - t_start:.1f} s")

    if rank == 0:
        report(args, job, n, K, W, total_ms, enc_ms, dec_ms, main["clocks"], e2e, strong, cfg2, cfg3, cfg4, multi)
    if world > 1:
        job.dist.destroy_process_group()


def report(args, job, n, K, W, total_ms, enc_ms, dec_ms, clocks, e2e, strong, cfg2, cfg3, cfg4, multi):
    world = job.world
    from bitnuc_b200 import device as dv
    ms_per_step = total_ms / K
    value = 2.0 * n * world / (ms_per_step * 1e-3) / 1e9
    dom = "encode_kernel" if enc_ms >= dec_ms else "decode_kernel"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(n),
        "kernels": {"timing": "second pass of the same K steps with a CUDA event around every launch (same stream)",
                    "encode_ms": enc_ms, "decode_ms": dec_ms,
                    "encode_gbases_s": n / (enc_ms * 1e-3) / 1e9, "decode_gbases_s": n / (dec_ms * 1e-3) / 1e9,
                    "encode_gbs": BYTES_PER_BASE * n / (enc_ms * 1e-3) / 1e9, "decode_gbs": BYTES_PER_BASE * n / (dec_ms * 1e-3) / 1e9},
        "roofline": roofline(dom, BYTES_PER_BASE * n, max(enc_ms, dec_ms)),
        "in_bench_check": "self-consistency (encode == the generator's word stream, decode == the input); parity against the oracle is "
                          "tests/ (-m gpu) and smoke()",
        "gpu_launches": 2 * K,
        "clocks": clocks,
    }
    if strong:
        line["strong"] = strong
    if e2e:
        nw8 = dv.words_for(n) * 8
        per_step = n + nw8

        def gb(s):  # Gbases/s of a schedule
            return 2.0 * n * world / s / 1e9
        pc = e2e["pcie"]
        moved = 2 * per_step * world / e2e["pipelined"] / 1e9
        line["e2e"] = {
            "value": gb(e2e["pipelined"]), "unit": UNIT, "h2d_bytes_per_step": per_step, "d2h_bytes_per_step": per_step,
            "ms_per_step": e2e["pipelined"] * 1e3, "steps": e2e["steps"], "mode": "pipelined",
            "serial_value": gb(e2e["serial"]), "serial_ms_per_step": e2e["serial"] * 1e3,
            "staggered_value": gb(e2e["staggered"]), "staggered_ms_per_step": e2e["staggered"] * 1e3,
            "pageable_value": gb(e2e["pageable"]),
            "pcie_ceiling_gbs": {k: v for k, v in pc.items() if k != "pattern_s"}, "pcie_gbs_moved": moved,
            "frac_of_pcie_ceiling": moved / pc["both"],
            "pcie_pattern_ms_per_step": pc["pattern_s"] * 1e3, "pcie_pattern_gbs": 2 * per_step * world / pc["pattern_s"] / 1e9,
            "frac_of_pcie_pattern": pc["pattern_s"] / e2e["pipelined"],
            "api": "bitnuc_b200.encode_np + decode_np (bn_encode / bn_decode, pinned host buffers, chunked 3-stage pipeline inside each "
                   "call), every H2D / D2H copy inside the timed region.  value = the pipelined schedule (two host threads, one bn_ctx "
                   "each: encode of step i+1 overlaps decode of step i over the full-duplex link) at every N.  serial_value: one "
                   "thread, encode then decode; staggered_value: serial with odd ranks decoding first; pageable_value: the serial "
                   "schedule on pageable numpy buffers, bounced through pinned stage buffers by the library (informational).  "
                   "pcie_ceiling_gbs: raw pinned copies of the same bytes on all N ranks at once -- the host of this box, not the "
                   "kernels, bounds e2e (frac_of_pcie_ceiling = bytes moved per second / the both-directions ceiling of two monolithic "
                   "copies).  pcie_pattern_*: the chunked copy sequence of the pipelined schedule itself (64 MiB up + 16 MiB down per "
                   "encode chunk, the reverse per decode chunk, three rotating streams per call) issued with NO kernels: what the link "
                   "gives that pattern; frac_of_pcie_pattern = its time / the e2e step time"}
    else:
        line["e2e"] = None

    cpu_rows = None
    if world == 1 and not args.skip_cpu:
        try:
            cpu_rows = cpu_baselines(args)
            allc = cpu_rows["codec_full"][0]
            one = next((r for r in cpu_rows["codec"] if r["isa"] == allc["isa"] and r["cores"] == 1), cpu_rows["codec"][0])
            line["cpu_baseline"] = {
                "value": allc["value"], "unit": UNIT, "cores": allc["cores"], "kind": "port",
                "best": allc["best"], "worst": allc["worst"], "reps": allc["reps"],
                "single_thread_value": one["value"], "configs0_1e6_bases_single_thread_value": cpu_rows["cfg0"][0]["value"],
                "sample": f"the whole workload ({allc['sample']}), encode + decode, median of {allc['reps']} repetitions on {allc['cores']} pinned "
                          f"threads in a process of its own; C restatement of the reference's {allc['isa']} path (oracle/bitnuc_oracle.c, "
                          f"-march=native)",
                "rows": cpu_rows["codec"], "host": cpu_rows.get("host")}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"error": str(ex)}

    def cpu(name):
        return cpu_rows[name] if cpu_rows else None

    configs = {}
    if cfg2:
        nl, nt = cfg2["n_local"], cfg2["n_total"]
        both = cfg2["as_2bit_ms"] + cfg2["from_2bit_ms"]
        configs["configs[2]"] = {
            "workload": "batched as_2bit / from_2bit of 2^28 random 31-mers (one u64 each), tight 31-byte records, device-resident; "
                        "records cut over the GPUs by index (strong)",
            "value": 2 * nt / (both * 1e-3) / 1e9, "unit": "Gkmers/s (as_2bit + from_2bit, each k-mer through both)",
            "as_2bit_ms": cfg2["as_2bit_ms"], "from_2bit_ms": cfg2["from_2bit_ms"],
            "as_2bit_gkmers_s": nt / (cfg2["as_2bit_ms"] * 1e-3) / 1e9, "from_2bit_gkmers_s": nt / (cfg2["from_2bit_ms"] * 1e-3) / 1e9,
            "roofline": roofline("as_2bit_tight_kernel", 39.0 * nl, cfg2["as_2bit_ms"]),
            "roofline_from_2bit": roofline("from_2bit_tight31_kernel", 39.0 * nl, cfg2["from_2bit_ms"]),
            "bytes_per_unit": "39 B per k-mer per kernel (31 ASCII + 8 packed); x the k-mers of one rank's launch",
            "check": "as_2bit(from_2bit(W)) == W & (2^62 - 1) on every record; from_2bit against its definition on a strided sample",
            "cpu_baseline": cpu("kmers")}
    if cfg3:
        pl, pt, rl, rt_ = cfg3["pairs_local"], cfg3["pairs_total"], cfg3["reads_local"], cfg3["reads_total"]
        configs["configs[3]"] = {
            "workload": "hdist over 2^30 pairs of packed 32-mers + base_counts / gc_content on 10 M x 150 bp reads, pairs and reads cut "
                        "over the GPUs by index (strong), counters all-reduced with NCCL inside the timed region",
            "hdist_pairs": {"value": pt / (cfg3["pairs_ms"] * 1e-3) / 1e9, "unit": "Gpairs/s", "ms": cfg3["pairs_ms"],
                            "roofline": roofline("hdist_pairs_kernel", 20.0 * pl, cfg3["pairs_ms"]),
                            "bytes_per_unit": "20 B per pair (8 + 8 in, 4 out)"},
            "hdist_total": {"value": 32 * pt / (cfg3["total_ms"] * 1e-3) / 1e9, "unit": "Gbases/s", "ms": cfg3["total_ms"],
                            "total_mismatches": cfg3["hdist_total"],
                            "includes": "hdist_sum_kernel + all_reduce(SUM) of the u64 total" if world > 1 else "hdist_sum_kernel",
                            "roofline": roofline("hdist_sum_kernel", 16.0 * pl, cfg3["total_ms"]),
                            "bytes_per_unit": "0.5 B per base (two packed streams)"},
            "base_counts_gc": {"value": rt_ / (cfg3["counts_fixed_ms"] * 1e-3) / 1e9, "unit": "Greads/s", "ms": cfg3["counts_fixed_ms"],
                               "form": "fixed-length reads (bn_base_counts_fixed_dev), per-read [u64;4] + f64 gc out, totals all-reduced",
                               "includes": "kernel + ncclAllReduce(4 x u64, sum)" if world > 1 else "kernel (one GPU: nothing to reduce)",
                               "kernel_only_ms": cfg3["counts_fixed_kernel_only_ms"],
                               "offset_indexed_ms": cfg3["counts_indexed_ms"],
                               "offset_indexed_value": rt_ / (cfg3["counts_indexed_ms"] * 1e-3) / 1e9,
                               "totals": cfg3["totals"], "gc_global": cfg3["gc_global"],
                               "roofline": roofline("base_counts_batch_kernel<fixed>", 80.0 * rl, cfg3["counts_fixed_kernel_only_ms"]),
                               "roofline_offset_indexed": roofline("base_counts_batch_kernel<indexed>", 96.0 * rl, cfg3["counts_indexed_ms"]),
                               "bytes_per_unit": "80 B per read (40 in, 32 counts + 8 gc out); the offset-indexed form also reads "
                                                 "16 B per read of word offsets and lengths: 96 B"},
            "comm": {"collective": "all_reduce(SUM) of 4 x u64 (32 bytes) over NCCL", "ms": cfg3["comm_ms"], "ranks": world,
                     "note": "latency-bound; timed alone on the same stream (K calls between two events)"},
            "check": "sum(hdist_pairs) == hdist total; pairs against the definition on a strided sample of the whole range; counts add up to "
                     "150 x reads; per-read gc == (gc / len) * 100 bit for bit",
            "cpu_baseline": {"hdist": cpu("hdist"), "base_counts": cpu("base_counts")}}
    if cfg4:
        nbytes_local = cfg4["local"]["bases"] + 8 * cfg4["local"]["words"] + 16 * cfg4["local"]["reads"]
        h2d = cfg4["bases"] + 8 * (cfg4["reads"] + world)
        d2h = 8 * cfg4["words"] + 8 * (cfg4["reads"] + world)
        configs["configs[4]"] = {
            "workload": f"variable-length read batch (50 bp - 10 kbp, offset-indexed), {args.batch_bases} bases per GPU "
                        f"({cfg4['bases']} in all; 8 GPUs = the stated ~32 Gbases), cut by byte volume on read boundaries (weak)",
            "reads": cfg4["reads"], "bases": cfg4["bases"],
            "device_resident": {"value": cfg4["bases"] / (cfg4["dev_ms"] * 1e-3) / 1e9, "unit": UNIT, "ms": cfg4["dev_ms"],
                                "roofline": roofline("encode_batch_kernel", float(nbytes_local), cfg4["dev_ms"]),
                                "bytes_per_unit": "1 B per base + 8 B per output word + 16 B per read of offsets (in + out)",
                                "includes": "bn_encode_batch_dev: the word-offset scan + the encode kernel"},
            "e2e": {"value": cfg4["bases"] / cfg4["e2e_clean_s"] / 1e9, "unit": UNIT, "ms": cfg4["e2e_clean_s"] * 1e3,
                    "h2d_bytes": h2d, "d2h_bytes": d2h, "pcie_gbs_moved": (h2d + d2h) / cfg4["e2e_clean_s"] / 1e9,
                    "api": "bn_encode_batch on pinned host buffers (chunks of whole reads through the 3-stage pipeline), wall clock, "
                           "barrier before, max over ranks, median of 3 after one warm-up call"},
            "e2e_injected_n": {"value": cfg4["bases"] / cfg4["e2e_injected_s"] / 1e9, "unit": UNIT, "ms": cfg4["e2e_injected_s"] * 1e3,
                               "injected": cfg4["injected"], "first_error": cfg4["first_error"],
                               "parity": "InvalidBase(78) at the closed-form first offset (MIN over ranks) and per-read status == the injected set"},
            "cpu_baseline": cpu("encode_batch")}
    if configs:
        line["configs"] = configs
    if multi:
        line["multi"] = multi
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=1_000_000_000, help="bases per GPU (BASELINE configs[1]: 1e9)")
    ap.add_argument("--batch-bases", type=int, default=4_000_000_000, help="configs[4]: bases per GPU (8 GPUs x 4e9 = 32 Gbases)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: skip the cpu_baseline legs")
    ap.add_argument("--skip-configs", action="store_true", help="profiling runs only: skip strong / configs[2..4] / multi")
    ap.add_argument("--cpu-suite", action="store_true", help="internal: run the cpu_baseline legs and print them as JSON")
    args = ap.parse_args()
    if args.cpu_suite:
        cpu_suite_main(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
