"""Randomised differential tests of the batch entry points against the oracle: many small batches with adversarial
shapes (reads ending exactly on tile / word / vector boundaries, runs of empty reads, one very long read among short
ones, batches that start at an odd byte of the buffer), each compared word for word / byte for byte."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGTacgt", dtype=np.uint8)


@pytest.fixture(scope="module")
def bn():
    import bitnuc_b200
    return bitnuc_b200


def _lens(rng, kind, n):
    if kind == "tile_edges":      # multiples of 32 / 64 / 65536 bases, +-1: word, vector and tile boundaries
        base = rng.choice([32, 64, 2048 * 32, 4096, 16], n)
        return np.maximum(0, base * rng.integers(0, 3, n) + rng.integers(-1, 2, n))
    if kind == "empties":
        return np.where(rng.random(n) < 0.8, 0, rng.integers(1, 100, n))
    if kind == "one_giant":
        l = rng.integers(1, 200, n)
        l[rng.integers(0, n)] = 300_000 + int(rng.integers(0, 64))
        return l
    if kind == "tiny":
        return rng.integers(0, 5, n)
    return (rng.pareto(1.2, n) * 40).astype(np.int64) % 50_000     # heavy tail


@pytest.mark.parametrize("kind", ["tile_edges", "empties", "one_giant", "tiny", "heavy_tail"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_encode_batch_fuzz(bn, kind, seed):
    rng = np.random.default_rng(1000 * seed + len(kind))
    n = int(rng.integers(1, 400))
    lens = _lens(rng, kind, n).astype(np.int64)
    lead = int(rng.integers(0, 40))
    offsets = (lead + np.concatenate([[0], np.cumsum(lens)])).astype(np.uint64)
    data = ACGT[rng.integers(0, 8, int(offsets[-1]) + int(rng.integers(0, 20)))]
    words, wo = bn.encode_batch(data, offsets)
    exp_words, exp_wo = [], [0]
    for r in range(n):
        seq = data[int(offsets[r]) : int(offsets[r + 1])]
        if seq.size:
            exp_words.append(onp.encode(seq))
        exp_wo.append(exp_wo[-1] + (seq.size + 31) // 32)
    exp = np.concatenate(exp_words) if exp_words else np.zeros(0, np.uint64)
    assert np.array_equal(wo, np.array(exp_wo, dtype=np.uint64))
    assert np.array_equal(words, exp)
    # one invalid byte somewhere (or none when the batch is empty): exact record / position / byte
    nz = np.flatnonzero(lens)
    if nz.size:
        r = int(rng.choice(nz))
        pos = int(rng.integers(0, lens[r]))
        bad = data.copy()
        bad[int(offsets[r]) + pos] = int(rng.choice([ord("N"), 0, 255, ord("U"), ord(" ")]))
        with pytest.raises(bn.NucleotideError) as ei:
            bn.encode_batch(bad, offsets)
        assert ei.value.key() == ("InvalidBase", int(bad[int(offsets[r]) + pos])) and (ei.value.record, ei.value.position) == (r, pos)
        _, _, status = bn.encode_batch(bad, offsets, per_read_status=True)
        expect = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
        expect[r] = pos
        assert np.array_equal(status, expect)


@pytest.mark.parametrize("seed", range(4))
def test_split_slice_kmers_fuzz(bn, seed):
    rng = np.random.default_rng(77 + seed)
    n = int(rng.integers(1, 300))
    lens = _lens(rng, ["tiny", "heavy_tail", "empties", "tile_edges"][seed], n).astype(np.int64) % 3000
    seqs = [ACGT[rng.integers(0, 4, int(l))].tobytes() for l in lens]
    packed = [oracle.PackedSequence(s) for s in seqs]
    words = np.array([w for p in packed for w in p.data], dtype=np.uint64)
    wo = np.concatenate([[0], np.cumsum([len(p.data) for p in packed])]).astype(np.uint64)
    idx = (rng.random(n) * (lens + 1)).astype(np.uint64)
    left, lo, right, ro = bn.split_packed_batch(words, wo, lens.astype(np.uint64), idx)
    o = [oracle.split_packed(p.data, int(l), int(i)) for p, l, i in zip(packed, lens, idx)]
    assert [int(x) for x in left] == [w for a, _ in o for w in a] and [int(x) for x in right] == [w for _, b in o for w in b]
    assert np.array_equal(lo, np.concatenate([[0], np.cumsum([len(a) for a, _ in o])]).astype(np.uint64))
    assert np.array_equal(ro, np.concatenate([[0], np.cumsum([len(b) for _, b in o])]).astype(np.uint64))
    nq = int(rng.integers(1, 500))
    qr = rng.integers(0, n, nq)
    a, b = (rng.random(nq) * (lens[qr] + 1)).astype(np.int64), (rng.random(nq) * (lens[qr] + 1)).astype(np.int64)
    qs, qe = np.minimum(a, b), np.maximum(a, b)
    data, oo = bn.slice_batch(words, wo[:-1], lens.astype(np.uint64), qr, qs, qe)
    assert data.tobytes() == b"".join(seqs[r][s:e] for r, s, e in zip(qr, qs, qe))
    long_seq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    for k in (int(rng.integers(1, 33)), 32):
        if long_seq.size >= k:
            got = bn.kmers(long_seq, k)
            pick = rng.integers(0, long_seq.size - k + 1, 50)
            assert [int(got[i]) for i in pick] == [oracle.as_2bit(long_seq[i : i + k]) for i in pick]


@pytest.mark.parametrize("kind", ["tile_edges", "empties", "one_giant", "tiny", "heavy_tail"])
@pytest.mark.parametrize("k", [1, 16, 21, 32])
def test_kmers_batch_fuzz(bn, kind, k):
    import torch
    from bitnuc_b200 import device as dv
    rng = np.random.default_rng(31 * k + len(kind))
    n = int(rng.integers(1, 300))
    lens = _lens(rng, kind, n).astype(np.int64) % 20_000
    lead = int(rng.integers(0, 40))
    offsets = (lead + np.concatenate([[0], np.cumsum(lens)])).astype(np.uint64)
    data = ACGT[rng.integers(0, 8, int(offsets[-1]) + 7)]
    words, oo = bn.kmers_batch(data, offsets, k)
    exp_oo = np.concatenate([[0], np.cumsum(np.maximum(lens - k + 1, 0))]).astype(np.uint64)
    assert np.array_equal(oo, exp_oo) and words.size == int(exp_oo[-1])
    for r in rng.integers(0, n, 25):   # per read: the single-sequence kernel (itself checked against the as_2bit loop) and the oracle
        seq = data[int(offsets[r]) : int(offsets[r + 1])]
        got = words[int(oo[r]) : int(oo[r + 1])]
        assert np.array_equal(got, bn.kmers(seq, k))
        for i in rng.integers(0, max(1, got.size), 5):
            if got.size:
                assert int(got[i]) == oracle.as_2bit(seq[i : i + k])
    d_words, d_oo, st = dv.kmers_batch(torch.from_numpy(data).cuda(), torch.from_numpy(offsets.view(np.int64)).cuda(), k)
    st.check()
    assert np.array_equal(d_oo.cpu().numpy().view(np.uint64), oo)
    assert np.array_equal(d_words[: words.size].cpu().numpy().view(np.uint64), words)
    # an invalid byte in a read that has windows is reported; in a read shorter than k it is not even looked at
    bad = data.copy()
    short = np.flatnonzero((lens > 0) & (lens < k))
    for r in short:
        bad[int(offsets[r])] = ord("N")
    w2, _ = bn.kmers_batch(bad, offsets, k)
    assert np.array_equal(w2, words)
    full = np.flatnonzero(lens >= k)
    if full.size:
        r = int(rng.choice(full))
        pos = int(rng.integers(0, lens[r]))
        bad[int(offsets[r]) + pos] = ord("n")
        with pytest.raises(bn.NucleotideError) as ei:
            bn.kmers_batch(bad, offsets, k)
        assert ei.value.key() == ("InvalidBase", ord("n")) and (ei.value.record, ei.value.position) == (r, pos)
