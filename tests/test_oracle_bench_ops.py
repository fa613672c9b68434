"""The timed CPU baselines (orc_bench_op) compute what the oracle's per-item functions compute: a baseline that
timed the wrong thing would be worse than none.  Small sizes; both code paths; one and several threads."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

SEED = 0x5EEDB17C0DE5
PATHS = [oracle.PATH_NAIVE] + ([oracle.PATH_AVX2] if oracle.have_avx2() else [])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("threads", [1, 3])
def test_codec_ops(path, threads):
    n = 100_003
    seq = onp.synth_ascii(SEED, 0, n)
    words = np.zeros((n + 31) // 32, dtype=np.uint64)
    times, _ = oracle.bench_op(oracle.OP_ENCODE, n, in0=seq, out0=words, path=path, threads=threads, reps=2)
    assert len(times) == 2 and all(t > 0 for t in times)
    assert np.array_equal(words, oracle.encode_np(seq))
    back = np.zeros(n + 32, dtype=np.uint8)
    oracle.bench_op(oracle.OP_DECODE, n, in0=words, out0=back, path=path, threads=threads, reps=1)
    assert np.array_equal(back[:n], seq)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("k,stride", [(31, 31), (31, 32), (16, 16), (7, 8)])
def test_kmer_ops(path, k, stride):
    n = 5000
    recs = onp.synth_ascii(SEED, 1, n * stride)
    out = np.zeros(n, dtype=np.uint64)
    oracle.bench_op(oracle.OP_AS_2BIT, n, in0=recs, out0=out, k=k, stride=stride, path=path, threads=2, reps=1)
    assert [int(x) for x in out[::37]] == [oracle.as_2bit(recs[r * stride : r * stride + k]) for r in range(0, n, 37)]
    asc = np.zeros(n * stride + 32, dtype=np.uint8)
    oracle.bench_op(oracle.OP_FROM_2BIT, n, in0=out, out0=asc, k=k, stride=stride, path=path, threads=2, reps=1)
    for r in range(0, n, 41):
        assert asc[r * stride : r * stride + k].tobytes() == recs[r * stride : r * stride + k].tobytes()


@pytest.mark.parametrize("path", PATHS)
def test_hdist_ops(path):
    n = 64 * 1000 + 17
    a = oracle.encode_np(onp.synth_ascii(SEED, 2, n))
    b = oracle.encode_np(onp.synth_ascii(SEED, 3, n))
    _, total = oracle.bench_op(oracle.OP_HDIST, n, in0=a, in1=b, path=path, threads=3, reps=1)
    assert total == oracle.hdist(a, b, n, wide=True)
    out = np.zeros(a.size, dtype=np.uint32)
    _, s = oracle.bench_op(oracle.OP_HDIST_PAIRS, a.size, in0=a, in1=b, out0=out, k=32, threads=2, reps=1)
    assert s == int(out.sum()) and [int(x) for x in out[:50]] == [oracle.hdist_scalar(int(a[i]), int(b[i]), 32) for i in range(50)]


def test_base_counts_gc_op():
    n_reads, read_len = 400, 150
    wpr = (read_len + 31) // 32
    words = np.zeros(n_reads * wpr, dtype=np.uint64)
    for r in range(n_reads):
        words[r * wpr : (r + 1) * wpr] = oracle.encode_np(onp.synth_ascii(SEED, 4, read_len + 32 * r)[-read_len:])
    counts = np.zeros(4 * n_reads, dtype=np.uint64)
    gc = np.zeros(n_reads, dtype=np.float64)
    _, s = oracle.bench_op(oracle.OP_BASE_COUNTS_GC, n_reads, in0=words, out0=counts, out1=gc, k=read_len, threads=2, reps=1)
    for r in range(0, n_reads, 13):
        w = words[r * wpr : (r + 1) * wpr]
        assert [int(x) for x in counts[4 * r : 4 * r + 4]] == list(oracle.base_counts(w, read_len))
        assert gc[r] == oracle.gc_content(w, read_len)
    assert s == int(counts.reshape(-1, 4)[:, 1:3].sum())
