"""Parity of the multi-device layer (``bn_multi_*`` through the C ABI) with the CPU oracle: sharded results must be
those of the single-device call -- and therefore the oracle's -- on the whole input, including which shard's error wins.

Every test runs on a one-GPU box with a context list that names device 0 three times (the sharding, the worker
threads, the rebasing of offsets / records and the mailbox all-reduce are all exercised; NCCL refuses one device twice);
with two or more devices the same tests also run across all of them with the NCCL all-reduce and with the NVLink
mailbox kernel.
"""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
ACGT_MIXED = np.frombuffer(b"ACGTacgt", dtype=np.uint8)


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


MODES = ["one_device_x3_p2p", "all_devices_nccl", "all_devices_p2p", "single"]


@pytest.fixture(scope="module", params=MODES)
def multi(request):
    import bitnuc_b200 as bn
    nd = bn._lib.load().bn_device_count()
    assert nd >= 1
    mode = request.param
    if mode.startswith("all_devices") and nd < 2:
        pytest.skip("needs >= 2 CUDA devices")
    if mode == "one_device_x3_p2p":
        m = bn.MultiContext([0, 0, 0], "p2p")
    elif mode == "single":
        m = bn.MultiContext([0], "nccl")
    else:
        m = bn.MultiContext(list(range(nd)), "nccl" if mode.endswith("nccl") else "p2p")
    m.set_chunk_bytes(64 << 10)   # many chunks per shard: the per-device pipelines run too
    yield m
    m.close()


def rand_seq(rng, n, mixed=False):
    a = ACGT_MIXED if mixed else ACGT
    return a[rng.integers(0, a.size, n)]


def test_modes(multi):
    import bitnuc_b200 as bn
    assert multi.n == len(multi.devices) >= 1
    if multi.n > 1 and len(set(multi.devices)) == multi.n and multi.reduce == "nccl":
        assert multi.nccl_version >= 20000
    st = multi.shard_units(1_000_003, 64)
    assert st[0] == 0 and st[-1] == 1_000_003 and all(a <= b for a, b in zip(st, st[1:]))
    assert all(s % 64 == 0 for s in st[1:-1])
    if multi.n > 1:   # NCCL cannot take one device twice: the library says so instead of hanging
        with pytest.raises(bn.BitnucCudaError):
            bn.MultiContext([0, 0], "nccl")


@pytest.mark.parametrize("n", [1, 31, 64, 65, 1000, 333_333, 1_000_000, (1 << 21) + 17])
def test_encode_decode_match_oracle(multi, n):
    rng = np.random.default_rng(n)
    seq = rand_seq(rng, n, mixed=True)
    words = multi.encode_np(seq)
    assert np.array_equal(words, oracle.encode_np(seq))
    assert np.array_equal(multi.decode_np(words, n), oracle.decode_np(words, n))
    import bitnuc_b200 as bn
    with pytest.raises(bn.NucleotideError) as ei:
        multi.decode_np(words[:-1], n)
    assert ei.value.key() == ("InvalidLength", n)


def test_encode_first_invalid_base_across_shards(multi):
    import bitnuc_b200 as bn
    rng = np.random.default_rng(5)
    n = 900_001
    seq = rand_seq(rng, n)
    st = multi.shard_units(n, 64)
    # one bad byte in every shard (different bytes), then progressively clean the earlier shards
    positions = [min(st[i + 1] - 1, st[i] + (st[i + 1] - st[i]) // 2 + i) for i in range(multi.n) if st[i + 1] > st[i]]
    bad = seq.copy()
    for j, p in enumerate(positions):
        bad[p] = ord("N") if j % 2 == 0 else ord("x")
    exp_words = oracle.encode_np(seq)
    for j, p in enumerate(positions):
        with pytest.raises(bn.NucleotideError) as ei:
            multi.encode_np(bad)
        e = ei.value
        assert e.key() == ("InvalidBase", int(bad[p])) and e.offset == p
        assert e.n_words == p // 32   # the reference leaves the words of the chunks before the failing chunk (avx.rs:142-143)
        out = np.zeros(exp_words.size, dtype=np.uint64)
        with pytest.raises(bn.NucleotideError):
            multi.encode_np(bad, out=out)
        assert np.array_equal(out[: p // 32], exp_words[: p // 32])
        bad[p] = seq[p]
    assert np.array_equal(multi.encode_np(bad), exp_words)
    with pytest.raises(bn.ReferencePanic):
        multi.encode_np(np.zeros(0, dtype=np.uint8))


@pytest.mark.parametrize("k,stride", [(31, 31), (31, 32), (32, 32), (21, 24), (1, 1), (7, 40)])
def test_kmer_batches(multi, k, stride):
    import bitnuc_b200 as bn
    rng = np.random.default_rng(k * 100 + stride)
    n = 20_011
    recs = rand_seq(rng, (n - 1) * stride + k, mixed=True)
    got = multi.as_2bit_batch(recs, n, k, stride)
    exp = np.array([oracle.as_2bit(recs[r * stride : r * stride + k]) for r in range(0, n, 97)], dtype=np.uint64)
    assert np.array_equal(got[::97], exp)
    single = bn.as_2bit_batch(recs, n, k, stride)
    assert np.array_equal(got, single)
    back = multi.from_2bit_batch(got, k, stride)
    for r in range(0, n, 501):
        assert bytes(back[r * stride : r * stride + k]) == bytes(oracle.from_2bit_alloc(int(got[r]), k))
    # first failing record in index order, offsets rebased to the caller's buffer
    victims = sorted(rng.choice(n, size=4, replace=False).tolist())
    bad = recs.copy()
    for r in victims:
        bad[r * stride + (k - 1)] = ord("N")
    with pytest.raises(bn.NucleotideError) as ei:
        multi.as_2bit_batch(bad, n, k, stride)
    assert ei.value.key() == ("InvalidBase", ord("N"))
    assert (ei.value.record, ei.value.offset) == (victims[0], victims[0] * stride + k - 1)
    with pytest.raises(bn.NucleotideError) as ei:
        multi.as_2bit_batch(recs, n, 33, 33)
    assert ei.value.key() == ("SequenceTooLong", 33)
    with pytest.raises(bn.NucleotideError) as ei:
        multi.from_2bit_batch(got, 33, 33)
    assert ei.value.key() == ("InvalidLength", 33)


@pytest.mark.parametrize("n", [1, 32, 4097, 700_001])
def test_hdist(multi, n):
    rng = np.random.default_rng(n + 1)
    a, b = oracle.encode_np(rand_seq(rng, n)), oracle.encode_np(rand_seq(rng, n))
    assert multi.hdist_total(a, b, n) == oracle.hdist(a, b, n, wide=True)
    assert multi.hdist(a, b, n) == oracle.hdist(a, b, n)
    for length in (32, 17, 0):
        got = multi.hdist_pairs(a, b, length)
        idx = list(range(0, a.size, max(1, a.size // 300)))
        assert [int(got[i]) for i in idx] == [oracle.hdist_scalar(int(a[i]), int(b[i]), length) for i in idx]
    import bitnuc_b200 as bn
    with pytest.raises(bn.NucleotideError) as ei:
        multi.hdist_total(a[:-1], b, n)
    assert ei.value.key() == ("InvalidLength", n)
    with pytest.raises(bn.NucleotideError) as ei:
        multi.hdist_pairs(a, b, 33)
    assert ei.value.key() == ("InvalidLength", 33)


@pytest.mark.parametrize("n", [1, 63, 64, 100_003, 1_500_017])
def test_base_counts_allreduce(multi, n):
    rng = np.random.default_rng(n + 2)
    words = oracle.encode_np(rand_seq(rng, n))
    counts, gc = multi.base_counts_gc(words, n)
    assert counts == oracle.base_counts(words, n)
    assert gc == oracle.gc_content(words, n)   # f64 ==: (gc as f64 / len as f64) * 100.0 from the reduced integer counts


def test_base_counts_batch_allreduce(multi):
    rng = np.random.default_rng(11)
    lens = np.where(rng.random(30_000) < 0.05, 0, rng.integers(1, 400, 30_000))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rand_seq(rng, int(offs[-1]))
    words, wo = multi.encode_batch(data, offs)
    counts4, gc, totals = multi.base_counts_batch(words, wo[:-1], lens.astype(np.uint64))
    tot = [0, 0, 0, 0]
    for r in range(0, lens.size):
        c = [int(x) for x in np.bincount(np.searchsorted(ACGT, data[int(offs[r]) : int(offs[r + 1])]), minlength=4)]
        tot = [a + b for a, b in zip(tot, c)]
        if r % 53 == 0:
            ps = oracle.PackedSequence(data[int(offs[r]) : int(offs[r + 1])].tobytes())
            assert [int(x) for x in counts4[r]] == ps.base_counts() == c
            assert gc[r] == ps.gc_content()
    assert totals == tot
    import bitnuc_b200 as bn
    bad_lens = lens.astype(np.uint64).copy()
    victim = int(lens.size * 0.8)
    bad_lens[victim] = 10**9   # needs more words than the buffer holds: InvalidLength for that read
    with pytest.raises(bn.NucleotideError) as ei:
        multi.base_counts_batch(words, wo[:-1], bad_lens)
    assert ei.value.key() == ("InvalidLength", 10**9) and ei.value.record == victim


@pytest.mark.parametrize("profile", ["short", "cfg5", "with_empties", "few"])
def test_encode_batch_sharded_by_volume(multi, profile):
    import bitnuc_b200 as bn
    rng = np.random.default_rng(29)
    lens = {
        "short": rng.integers(1, 160, 20_000),
        "cfg5": 50 + rng.integers(0, 9951, 700),
        "with_empties": np.where(rng.random(9000) < 0.4, 0, rng.integers(1, 300, 9000)),
        "few": np.array([5, 0]),   # fewer reads than shards
    }[profile]
    offs = (np.concatenate([[0], np.cumsum(lens)]) + 7).astype(np.uint64)   # the batch starts at an odd byte
    data = np.concatenate([np.full(7, ord("#"), dtype=np.uint8), rand_seq(rng, int(lens.sum()), mixed=True)])
    st = multi.shard_reads(offs)
    assert st[0] == 0 and st[-1] == lens.size and all(a <= b for a, b in zip(st, st[1:]))
    words, wo = multi.encode_batch(data, offs)
    exp_words, exp_wo = [], [0]
    for r in range(lens.size):
        ps = oracle.PackedSequence(data[int(offs[r]) : int(offs[r + 1])].tobytes())
        exp_words.extend(ps.data)
        exp_wo.append(exp_wo[-1] + len(ps.data))
    assert np.array_equal(wo, np.array(exp_wo, dtype=np.uint64))
    assert np.array_equal(words, np.array(exp_words, dtype=np.uint64))
    nonempty = [r for r in range(lens.size) if lens[r] > 0]
    # one injected N per shard that has a non-empty read; the first in input order wins with its global read index
    victims = []
    for i in range(multi.n):
        cand = [r for r in nonempty if st[i] <= r < st[i + 1]]
        if cand:
            victims.append(cand[len(cand) // 2])
    bad, where = data.copy(), {}
    for r in victims:
        where[r] = int(rng.integers(0, lens[r]))
        bad[int(offs[r]) + where[r]] = ord("N")
    with pytest.raises(bn.NucleotideError) as ei:
        multi.encode_batch(bad, offs)
    e, r0 = ei.value, victims[0]
    assert e.key() == ("InvalidBase", ord("N")) and (e.record, e.position, e.offset) == (r0, where[r0], int(offs[r0]) + where[r0])
    _, wo2, status = multi.encode_batch(bad, offs, per_read_status=True)
    expect = np.full(lens.size, 0xFFFFFFFF, dtype=np.uint32)
    for r, pos in where.items():
        expect[r] = pos
    assert np.array_equal(status, expect) and np.array_equal(wo2, wo)
    with pytest.raises(ValueError):
        multi.encode_batch(data, offs[::-1].copy())


def test_device_resident_reductions(multi):
    import torch
    import bitnuc_b200 as bn
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(3)
    n = multi.n
    devs = [torch.device("cuda", d) for d in multi.devices]
    sizes = [200_003 + 64 * 1000 * i for i in range(n)]
    seqs = [rand_seq(rng, s) for s in sizes]
    others = [rand_seq(rng, s) for s in sizes]
    w = [torch.from_numpy(oracle.encode_np(s).view(np.int64)).to(d) for s, d in zip(seqs, devs)]
    w2 = [torch.from_numpy(oracle.encode_np(s).view(np.int64)).to(d) for s, d in zip(others, devs)]
    counts = [torch.empty(4, dtype=torch.int64, device=d) for d in devs]
    gc = [torch.empty(1, dtype=torch.float64, device=d) for d in devs]
    whole = np.concatenate(seqs)
    exp_counts = [int(x) for x in np.bincount(np.searchsorted(ACGT, whole), minlength=4)]
    exp_gc = float((np.float64(exp_counts[1] + exp_counts[2]) / np.float64(whole.size)) * np.float64(100.0))
    for rep in range(5):   # consecutive epochs of the mailbox (both parities), back to back without a sync in between
        multi.base_counts_dev(w, sizes, counts, gc)
    multi.synchronize()
    for i in range(n):   # every device holds the global value
        assert counts[i].tolist() == exp_counts
        assert gc[i].item() == exp_gc
    ms = multi.last_ms()
    assert len(ms) == n and all(0.0 < t < 1000.0 for t in ms)
    total = [torch.empty(1, dtype=torch.int64, device=d) for d in devs]
    multi.hdist_dev(w, w2, sizes, total)
    multi.synchronize()
    exp_total = sum(oracle.hdist(oracle.encode_np(a), oracle.encode_np(b), s, wide=True) for a, b, s in zip(seqs, others, sizes))
    assert all(t.item() == exp_total for t in total)
    # fixed-length reads (cfg 4 shape: 150 bp -> 5 words per read), per-read outputs stay local, totals are global
    read_len, wpr = 150, 5
    n_reads = [3001 + 17 * i for i in range(n)]
    reads = [rand_seq(rng, r * read_len).reshape(r, read_len) for r in n_reads]
    packed = []
    for rd in reads:
        pw = np.zeros((rd.shape[0], wpr), dtype=np.uint64)
        for r in range(0, rd.shape[0]):
            pw[r] = oracle.encode_np(rd[r])
        packed.append(pw)
    fw = [torch.from_numpy(p.view(np.int64).reshape(-1)).to(d) for p, d in zip(packed, devs)]
    totals = [torch.empty(4, dtype=torch.int64, device=d) for d in devs]
    c4 = [torch.empty(r * 4, dtype=torch.int64, device=d) for r, d in zip(n_reads, devs)]
    gcr = [torch.empty(r, dtype=torch.float64, device=d) for r, d in zip(n_reads, devs)]
    multi.base_counts_fixed_dev(fw, n_reads, read_len, totals, gc, c4, gcr)
    multi.synchronize()
    allr = np.concatenate([r.reshape(-1) for r in reads])
    exp_tot = [int(x) for x in np.bincount(np.searchsorted(ACGT, allr), minlength=4)]
    for i in range(n):
        assert totals[i].tolist() == exp_tot
        assert gc[i].item() == float((np.float64(exp_tot[1] + exp_tot[2]) / np.float64(allr.size)) * np.float64(100.0))
        got4 = c4[i].cpu().numpy().reshape(-1, 4)
        for r in range(0, n_reads[i], 211):
            ps = oracle.PackedSequence(reads[i][r].tobytes())
            assert got4[r].tolist() == ps.base_counts() and gcr[i][r].item() == ps.gc_content()
    # the collective alone, many epochs in a row
    bufs = [torch.tensor([i + 1, 10 * (i + 1), 0, 7], dtype=torch.int64, device=d) for i, d in enumerate(devs)]
    base = [b.clone() for b in bufs]
    for rep in range(40):
        for b, b0 in zip(bufs, base):
            b.copy_(b0)
        for d in devs:   # the copies above ran on torch's streams: order the contexts' streams after them
            torch.cuda.synchronize(d)
        multi.allreduce_u64_dev(bufs, 4)
        multi.synchronize()
        s = n * (n + 1) // 2
        assert all(b.tolist() == [s, 10 * s, 0, 7 * n] for b in bufs)
    with pytest.raises(bn.NucleotideError):
        multi.base_counts_dev([t[:1] for t in w], sizes, counts, gc)


def test_kernel_timing_through_the_abi():
    """bn_ctx_set_timing / bn_last_kernel_ms: device time of the last device-pointer call."""
    import ctypes as C
    import torch
    import bitnuc_b200 as bn
    from bitnuc_b200 import device as dv
    ctx = bn.default_context(0)
    ms = C.c_float(0)
    seq = torch.from_numpy(rand_seq(np.random.default_rng(0), 1 << 24)).cuda()
    dv.encode(seq)[1].check()   # the first launch of a kernel loads its module on the host, between the two events
    try:
        assert ctx.lib.bn_ctx_set_timing(ctx.handle, 1) == 0
        assert ctx.lib.bn_last_kernel_ms(ctx.handle, C.byref(ms)) == -2   # nothing timed yet
        words, status = dv.encode(seq)
        assert ctx.lib.bn_last_kernel_ms(ctx.handle, C.byref(ms)) == 0
        status.check()
        assert 0.0 < ms.value < 50.0
        gbs = 1.25 * seq.numel() / (ms.value * 1e-3) / 1e9
        assert gbs > 500.0, gbs   # a 16 Mbase encode is far from the roofline but nowhere near a host-timed number
    finally:
        ctx.lib.bn_ctx_set_timing(ctx.handle, 0)
    assert ctx.lib.bn_last_kernel_ms(ctx.handle, C.byref(ms)) == -2
