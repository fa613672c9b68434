"""Timed CPU baselines of every hot-path row (bench.py's cpu_baseline / --impl reference legs only).  TEST INFRASTRUCTURE.

The reference is single-threaded Rust that cannot be built here (no cargo); what is timed is the oracle's C restatement
of it (oracle/bitnuc_oracle.c): the AVX2 path the reference selects on this host (``-C target-cpu=native`` ->
``-march=native``) and the scalar path (its ``nosimd`` feature, /root/reference/Cargo.toml:13-14), each on one pinned thread
and chunked over all host threads (the chunking is the harness's).  Inputs are the same counter-based synthetic streams
the GPU arm uses.  Every figure is the median of ``reps`` repetitions with min / max beside it."""
from __future__ import annotations

import ctypes as C
import os
import statistics
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import oracle

SEED = 0x5EEDB17C0DE5


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth_ascii_mt(stream: int, first_base: int, n: int, threads: int | None = None) -> np.ndarray:
    """orc_synth_ascii over several Python threads (ctypes drops the GIL): 1e9 bases in well under a second."""
    threads = threads or host_threads()
    out = np.empty(n, dtype=np.uint8)
    L = oracle.lib()
    per = ((n + threads - 1) // threads + 31) // 32 * 32
    base = out.ctypes.data

    def fill(t):
        o = t * per
        m = min(per, n - o)
        if m > 0:
            L.orc_synth_ascii(C.c_uint64(SEED), C.c_uint64(stream), C.c_uint64(first_base + o), m, C.c_void_p(base + o))

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(fill, range(threads)))
    return out


def synth_words(stream: int, first_word: int, n: int) -> np.ndarray:
    """splitmix64((seed ^ stream * golden) + j) for j in [first_word, first_word + n) -- vectorised."""
    with np.errstate(over="ignore"):
        z = (np.uint64(SEED) ^ (np.uint64(stream) * np.uint64(0x9E3779B97F4A7C15))) + np.arange(first_word, first_word + n, dtype=np.uint64)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _summ(times, units, unit, threads, path, what, sample):
    med = statistics.median(times)
    return {"value": units / med / 1e9, "unit": unit, "cores": threads, "kind": "port",
            "isa": {oracle.PATH_AVX2: "avx2", oracle.PATH_NAIVE: "scalar"}[path],
            "best": units / min(times) / 1e9, "worst": units / max(times) / 1e9, "reps": len(times),
            "median_ms": med * 1e3, "pinned_threads": True, "what": what, "sample": sample}


class Suite:
    """Builds each op's buffers once (outputs touched before timing), then times it for every (isa, threads) asked for."""

    def __init__(self, threads: int | None = None, reps: int = 5, scale: float = 1.0):
        self.threads = threads or host_threads()
        self.reps, self.scale = reps, scale
        self.paths = ([oracle.PATH_AVX2] if oracle.have_avx2() else []) + [oracle.PATH_NAIVE]

    def _grid(self):
        for path in self.paths:
            for th in sorted({1, self.threads}):
                yield path, th

    def _n(self, single: int, many: int, threads: int, path: int) -> int:
        n = many if threads > 1 else single
        if path == oracle.PATH_NAIVE:
            n //= 2
        return max(4096, int(n * self.scale))

    def codec(self, n_single=1 << 26, n_many=1 << 28, seq: np.ndarray | None = None, grid=None):
        """encode + decode of one contiguous sequence (packing/avx.rs:130-151, unpacking/avx.rs:117-153 | naive.rs)."""
        out = []
        cap = max(n_single, n_many) if seq is None else seq.size
        seq = synth_ascii_mt(0, 0, cap) if seq is None else seq
        words = np.ones((cap + 31) // 32 + 8, dtype=np.uint64)
        back = np.ones(cap + 32, dtype=np.uint8)
        for path, th in (grid or self._grid()):
            n = min(cap, self._n(n_single, n_many, th, path)) if grid is None else cap
            te, _ = oracle.bench_op(oracle.OP_ENCODE, n, in0=seq, out0=words, path=path, threads=th, reps=self.reps)
            td, _ = oracle.bench_op(oracle.OP_DECODE, n, in0=words, out0=back, path=path, threads=th, reps=self.reps)
            if not np.array_equal(back[:n][-4096:], seq[:n][-4096:]):
                raise RuntimeError("cpu baseline: codec round trip is wrong")
            r = _summ([a + b for a, b in zip(te, td)], 2 * n, "Gbases/s", th, path, "encode + decode round trip",
                      f"{n} bases of stream 0")
            r["encode_gbases_s"] = n / statistics.median(te) / 1e9
            r["decode_gbases_s"] = n / statistics.median(td) / 1e9
            out.append(r)
        return out

    def kmers(self, k=31, n_single=1 << 21, n_many=1 << 24):
        """batched as_2bit / from_2bit of k-mers, tight records (packing/avx.rs:76-128, unpacking/avx.rs:50-114 | naive.rs)."""
        out = []
        cap = int(max(n_single, n_many) * self.scale)
        packed_in = synth_words(1, 0, cap) & np.uint64((1 << (2 * k)) - 1)
        recs = np.ones(cap * k + 32, dtype=np.uint8)
        oracle.bench_op(oracle.OP_FROM_2BIT, cap, in0=packed_in, out0=recs, k=k, stride=k, path=self.paths[0], threads=self.threads, reps=1)
        packed = np.ones(cap, dtype=np.uint64)
        for path, th in self._grid():
            n = min(cap, self._n(n_single, n_many, th, path))
            ta, _ = oracle.bench_op(oracle.OP_AS_2BIT, n, in0=recs, out0=packed, k=k, stride=k, path=path, threads=th, reps=self.reps)
            if not np.array_equal(packed[:n], packed_in[:n]):
                raise RuntimeError("cpu baseline: as_2bit is wrong")
            tf, _ = oracle.bench_op(oracle.OP_FROM_2BIT, n, in0=packed_in, out0=recs, k=k, stride=k, path=path, threads=th, reps=self.reps)
            r = _summ([a + b for a, b in zip(ta, tf)], 2 * n, "Gkmers/s", th, path, f"as_2bit + from_2bit of {k}-mers",
                      f"{n} records of stream 1")
            r["as_2bit_gkmers_s"] = n / statistics.median(ta) / 1e9
            r["from_2bit_gkmers_s"] = n / statistics.median(tf) / 1e9
            out.append(r)
        return out

    def hdist(self, n_single=1 << 23, n_many=1 << 26):
        """whole-sequence hdist (hamming/multi.rs:122-160; AVX2 :12-67) and per-pair hdist_scalar (hamming/scalar.rs:11-48)."""
        out = []
        cap = int(max(n_single, n_many) * self.scale)
        u, v = synth_words(2, 0, cap), synth_words(3, 0, cap)
        d = np.ones(cap, dtype=np.uint32)
        for path, th in self._grid():
            n = min(cap, self._n(n_single, n_many, th, path))
            tt, total = oracle.bench_op(oracle.OP_HDIST, 32 * n, in0=u, in1=v, path=path, threads=th, reps=self.reps)
            out.append(_summ(tt, 32 * n, "Gbases/s", th, path, "hdist, whole sequence", f"{n} word pairs of streams 2, 3"))
            out[-1]["total"] = total
            if path == self.paths[0]:  # hdist_scalar has one code path
                tp, s = oracle.bench_op(oracle.OP_HDIST_PAIRS, n, in0=u, in1=v, out0=d, k=32, threads=th, reps=self.reps)
                if s != total:
                    raise RuntimeError("cpu baseline: hdist_pairs sum differs from the hdist total")
                out.append(_summ(tp, n, "Gpairs/s", th, oracle.PATH_NAIVE, "hdist_scalar per pair, len 32", f"{n} pairs of streams 2, 3"))
        return out

    def base_counts(self, read_len=150, n_single=1 << 16, n_many=1 << 19):
        """base_counts() + gc_content() per read, each through its own to_vec() (analysis.rs:3-39, sequence.rs:116-135, 198-212)."""
        out = []
        cap = int(max(n_single, n_many) * self.scale)
        wpr = (read_len + 31) // 32
        words = synth_words(4, 0, wpr * cap)
        if read_len % 32:
            words.reshape(cap, wpr)[:, -1] &= np.uint64((1 << (2 * (read_len % 32))) - 1)
        counts = np.ones(4 * cap, dtype=np.uint64)
        gc = np.ones(cap, dtype=np.float64)
        for th in sorted({1, self.threads}):
            n = min(cap, self._n(n_single, n_many, th, oracle.PATH_AVX2))
            t, s = oracle.bench_op(oracle.OP_BASE_COUNTS_GC, n, in0=words, out0=counts, out1=gc, k=read_len, path=oracle.PATH_NAIVE,
                                   threads=th, reps=self.reps)
            c4 = counts[: 4 * n].reshape(n, 4)
            if int(c4.sum()) != read_len * n or s != int(c4[:, 1:3].sum()):
                raise RuntimeError("cpu baseline: base_counts is wrong")
            out.append(_summ(t, n, "Greads/s", th, oracle.PATH_NAIVE, f"base_counts + gc_content per {read_len} bp read",
                             f"{n} reads of stream 4"))
        return out


def pick(rows, isa: str, cores_gt_1: bool):
    for r in rows:
        if r["isa"] == isa and (r["cores"] > 1) == cores_gt_1:
            return r
    return None


def encode_batch(suite: Suite, target_bases: int = 1 << 28):
    """BASELINE configs[4] on the CPU: a variable-length read batch (50 bp - 10 kbp, the cfg-5 length profile), one
    PackedSequence::new per read (sequence.rs:40-52 -> packing/avx.rs:130-151 | naive.rs:22-43), threads cut by volume."""
    with np.errstate(over="ignore"):
        n_guess = int(target_bases / 5025 * 1.05) + 16
        r = np.arange(n_guess, dtype=np.uint64) + np.uint64(SEED + 5)
        z = r + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        lens = np.uint64(50) + (z ^ (z >> np.uint64(31))) % np.uint64(9951)
    cum = np.cumsum(lens)
    n = int(np.searchsorted(cum, target_bases)) + 1
    lens = lens[:n]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    woff = np.concatenate([[0], np.cumsum((lens + np.uint64(31)) // np.uint64(32))]).astype(np.uint64)
    total = int(off[-1])
    data = synth_ascii_mt(5, 0, total)
    words = np.ones(int(woff[-1]) + 8, dtype=np.uint64)
    out = []
    for path, th in suite._grid():
        t, s = oracle.bench_op(oracle.OP_ENCODE_BATCH, n, in0=data, in1=off, out0=words, out1=woff, path=path, threads=th,
                               reps=suite.reps if th > 1 else max(2, suite.reps // 2))
        if s != int(woff[-1]):
            raise RuntimeError("cpu baseline: encode_batch wrote the wrong number of words")
        out.append(_summ(t, total, "Gbases/s", th, path, "variable-length read batch encode (50 bp - 10 kbp)",
                         f"{n} reads, {total} bases of stream 5"))
    return out
