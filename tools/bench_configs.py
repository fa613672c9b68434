#!/usr/bin/env python
"""Per-kernel measurements of BASELINE.json configs 2-5 on one B200 (device-resident, CUDA events).

Every line reports the kernel's algorithmic bytes over its event-timed duration against
the measured HBM peak, after checking the result against a size-independent property.  Not the
headline benchmark (that is bench.py); this is the evidence behind the per-kernel rooflines in
DESIGN.md.

    python tools/bench_configs.py [--scale 1.0] [--reps 10] [--only cfg3,cfg4,short,...]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from bitnuc_b200 import device as dv
from bitnuc_b200 import synth

SEED = 0x5EEDB17C0DE5
M64 = (1 << 64) - 1


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def timed(fn, reps):
    for _ in range(int(os.environ.get("BN_WARM", "3"))):   # BN_WARM=0: profiling runs (one launch to capture)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, nbytes, units, unit_name, extra=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "ms": round(ms, 5), "algorithmic_bytes": nbytes, "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak(), 4),
            "frac_of_8000": round(gbs / 8000.0, 4), unit_name + "/s": round(units / (ms * 1e-3) / 1e9, 3), "unit": "G" + unit_name}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)
    return line


def cfg2(scale, reps):
    """encode + decode of one contiguous sequence (also a ragged 2^30+17 one)."""
    for n in (int(1_000_000_000 * scale), int(((1 << 30) + 17) * scale)):
        asc = dv.synth_ascii(SEED, 0, 0, n)
        words = torch.empty(dv.words_for(n), dtype=torch.int64, device="cuda")
        back = torch.empty(n, dtype=torch.uint8, device="cuda")
        st = dv.Status("cuda")
        ms_e = timed(lambda: dv.encode(asc, out=words, status=st), reps)
        ms_d = timed(lambda: dv.decode(words, n, out=back), reps)
        st.check()
        expect = dv.synth_words(SEED, 0, 0, dv.words_for(n))
        if n % 32:
            expect[-1] &= (1 << (2 * (n % 32))) - 1
        assert torch.equal(words, expect) and torch.equal(back, asc)
        report(f"cfg2 encode n={n}", ms_e, 1.25 * n, n, "bases")
        report(f"cfg2 decode n={n}", ms_d, 1.25 * n, n, "bases")
        del asc, words, back, expect


def cfg3(scale, reps):
    """batched as_2bit / from_2bit of 2^28 31-mers (tight 31-byte records and 32-byte padded records)."""
    n = int((1 << 28) * scale)
    words = dv.synth_words(SEED, 1, 0, n)
    expect = words & ((1 << 62) - 1)
    for layout, stride in (("tight", 31), ("padded", 32)):
        recs = torch.zeros((n - 1) * stride + 31 + (1 if stride == 32 else 0), dtype=torch.uint8, device="cuda")
        if stride == 32:
            dv.from_2bit_batch(words, 32, 32, out=recs)  # fills all 32 slots (k = 32 fast path)
            ms_f = timed(lambda: dv.from_2bit_batch(words, 32, 32, out=recs), reps)
            kf = 32
        else:
            ms_f = timed(lambda: dv.from_2bit_batch(words, 31, 31, out=recs), reps)
            kf = 31
        packed = torch.empty(n, dtype=torch.int64, device="cuda")
        st = dv.Status("cuda")
        ms_a = timed(lambda: dv.as_2bit_batch(recs, n, 31, stride, out=packed, status=st), reps)
        st.check()
        assert torch.equal(packed, expect), layout
        report(f"cfg3 as_2bit k=31 {layout} n={n}", ms_a, (stride + 8) * n, n, "kmers")
        report(f"cfg3 from_2bit k={kf} {layout} n={n}", ms_f, (stride + 8) * n, n, "kmers")
        del recs, packed
    del words, expect


def cfg4(scale, reps):
    """hdist over 2^30 pairs of packed 32-mers (+ whole-sequence total), base_counts/gc on 10M x 150 bp reads."""
    n = int((1 << 30) * scale)
    u, v = dv.synth_words(SEED, 2, 0, n), dv.synth_words(SEED, 3, 0, n)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    tot = torch.empty(1, dtype=torch.int64, device="cuda")
    ms_p = timed(lambda: dv.hdist_pairs(u, v, 32, out=out), reps)
    ms_t = timed(lambda: dv.hdist(u, v, 32 * n, out=tot), reps)
    assert int(out.sum(dtype=torch.int64).item()) == int(tot.item())  # checksum of checksums
    x = (u[:100000] ^ v[:100000]).cpu().numpy().view(np.uint64)
    m = (x | (x >> np.uint64(1))) & np.uint64(0x5555555555555555)
    ref = np.unpackbits(m.view(np.uint8).reshape(-1, 8), axis=1).sum(axis=1)
    assert np.array_equal(out[:100000].cpu().numpy(), ref.astype(np.int32))
    report(f"cfg4 hdist_pairs len=32 n={n}", ms_p, 20 * n, n, "pairs")
    report(f"cfg4 hdist whole-sequence n_bases={32 * n}", ms_t, 16 * n, 32 * n, "bases", {"total": int(tot.item())})
    ms_c = timed(lambda: dv.base_counts(u, 32 * n), reps)
    counts, gc = dv.base_counts(u, 32 * n)
    assert int(counts.sum().item()) == 32 * n
    report(f"cfg4 base_counts whole-sequence n_bases={32 * n}", ms_c, 8 * n, 32 * n, "bases", {"counts": counts.tolist(), "gc": gc.item()})
    del u, v, out

    reads = int(10_000_000 * scale)
    words = dv.synth_words(SEED, 4, 0, 5 * reads)
    words.view(reads, 5)[:, 4] &= (1 << 44) - 1  # 150 = 4*32 + 22 bases: mask the tail of word 4
    wo = torch.arange(reads, dtype=torch.int64, device="cuda") * 5
    lens = torch.full((reads,), 150, dtype=torch.int64, device="cuda")
    counts4 = torch.empty((reads, 4), dtype=torch.int64, device="cuda")
    gcs = torch.empty(reads, dtype=torch.float64, device="cuda")
    totals = torch.empty(4, dtype=torch.int64, device="cuda")
    ms_b = timed(lambda: dv.base_counts_batch(words, wo, lens, counts4=counts4, gc=gcs, totals=totals), reps)
    assert torch.equal(counts4.sum(dim=0), totals) and int(totals.sum().item()) == 150 * reads
    gc_np = (counts4[:, 1] + counts4[:, 2]).cpu().numpy().astype(np.float64)
    assert np.array_equal(gcs.cpu().numpy(), (gc_np / np.float64(150.0)) * np.float64(100.0))  # reference operation order
    report(f"cfg4 base_counts+gc per read (offset-indexed) reads={reads} x 150bp", ms_b, 80 * reads, reads, "reads",
           {"totals": totals.tolist(), "note": "also fetches 16 B/read of offsets+lengths: 96 B/read of real traffic"})
    c2, g2, t2 = torch.empty_like(counts4), torch.empty_like(gcs), torch.empty_like(totals)
    ms_f = timed(lambda: dv.base_counts_fixed(words, reads, 150, counts4=c2, gc=g2, totals=t2), reps)
    assert torch.equal(c2, counts4) and torch.equal(g2, gcs) and torch.equal(t2, totals)
    report(f"cfg4 base_counts+gc per read (fixed length) reads={reads} x 150bp", ms_f, 80 * reads, reads, "reads")


def cfg5(scale, reps):
    """variable-length read batch (50 bp - 10 kbp, ~32 Gbases at scale 1), offset-indexed, with injected N."""
    target = int(32e9 * scale)
    lens = synth.cfg5_read_lengths(target, SEED)
    n_reads = lens.size
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    total = int(offsets[-1])
    data = dv.synth_ascii(SEED, 5, 0, total)
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    max_words = total // 32 + n_reads
    words = torch.empty(max_words, dtype=torch.int64, device="cuda")
    wo = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    rs = torch.empty(n_reads, dtype=torch.int32, device="cuda")
    ctx = dv.api.default_context(0)
    scratch = torch.empty(ctx.lib.bn_encode_batch_scratch_bytes(n_reads, total), dtype=torch.uint8, device="cuda")
    st = dv.Status("cuda")

    def run():
        dv.raise_for(ctx.lib.bn_encode_batch_dev(ctx.handle, dv._stream(), dv._ptr(data), dv._ptr(d_off), n_reads, total, dv._ptr(words),
                                                 dv._ptr(wo), dv._ptr(rs), dv._ptr(st.word), dv._ptr(scratch)))

    ms = timed(run, max(2, reps // 2))
    st.check()
    n_words = int(wo[-1].item())
    assert n_words == int(((lens + np.uint64(31)) // np.uint64(32)).sum())
    # round trip of a few reads through decode
    for ridx in (0, 1, n_reads // 2, n_reads - 1):
        w0, ln = int(wo[ridx].item()), int(lens[ridx])
        back = dv.decode(words[w0 : w0 + (ln + 31) // 32].contiguous(), ln)
        assert torch.equal(back, data[int(offsets[ridx]) : int(offsets[ridx]) + ln])
    nbytes = total + 8 * n_words + 16 * n_reads
    report(f"cfg5 encode_batch reads={n_reads} bases={total}", ms, nbytes, total, "bases")
    # injected N: read r gets 'N' at h2(r) % len iff h1(r) % 100003 == 0
    victims, pos = synth.cfg5_injected_n(n_reads, lens)
    if victims.size == 0:
        victims = np.array([n_reads // 3])
        pos = np.array([int(lens[n_reads // 3]) // 2], dtype=np.int64)
    idx = torch.from_numpy((offsets[victims].astype(np.int64) + pos)).cuda()
    data[idx] = ord("N")
    run()
    try:
        st.check()
        raise AssertionError("InvalidBase not reported")
    except dv._lib.NucleotideError as e:
        assert e.key() == ("InvalidBase", ord("N")) and e.offset == int(offsets[victims[0]]) + int(pos[0]), (e.key(), e.offset)
    bad = torch.nonzero(rs != -1).flatten().cpu().numpy()
    assert np.array_equal(bad, victims) and np.array_equal(rs[torch.from_numpy(victims).cuda()].cpu().numpy().astype(np.int64), pos)
    print(json.dumps({"cfg5 error parity": "ok", "injected": int(victims.size), "first": [int(victims[0]), int(pos[0])]}), flush=True)


def short_reads(scale, reps):
    """encode_batch on a short-read profile (40 M x 100-151 bp, Illumina-like) and split_packed of the result
    at a barcode|UMI boundary (idx = 26) -- SURVEY.md 8(f) rank 1."""
    n_reads = int(40_000_000 * scale)
    lens = (100 + synth.splitmix64(np.arange(n_reads, dtype=np.uint64) + np.uint64(SEED + 6)) % np.uint64(52)).astype(np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    total = int(offsets[-1])
    data = dv.synth_ascii(SEED, 6, 0, total)
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    words = torch.empty(total // 32 + n_reads, dtype=torch.int64, device="cuda")
    wo = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    ctx = dv.api.default_context(0)
    scratch = torch.empty(ctx.lib.bn_encode_batch_scratch_bytes(n_reads, total), dtype=torch.uint8, device="cuda")
    st = dv.Status("cuda")

    def run():
        dv.raise_for(ctx.lib.bn_encode_batch_dev(ctx.handle, dv._stream(), dv._ptr(data), dv._ptr(d_off), n_reads, total, dv._ptr(words),
                                                 dv._ptr(wo), None, dv._ptr(st.word), dv._ptr(scratch)))

    ms = timed(run, reps)
    st.check()
    n_words = int(wo[-1].item())
    assert n_words == int(((lens + np.uint64(31)) // np.uint64(32)).sum())
    for ridx in (0, 1, n_reads // 2, n_reads - 1):
        w0, ln = int(wo[ridx].item()), int(lens[ridx])
        assert torch.equal(dv.decode(words[w0 : w0 + (ln + 31) // 32].contiguous(), ln), data[int(offsets[ridx]) : int(offsets[ridx]) + ln])
    report(f"short reads encode_batch reads={n_reads} bases={total}", ms, total + 8 * n_words + 16 * n_reads, total, "bases")

    d_lens = torch.from_numpy(lens.view(np.int64)).cuda()
    idx = torch.full((n_reads,), 26, dtype=torch.int64, device="cuda")
    words = words[:n_words]
    left = torch.empty(n_words + n_reads, dtype=torch.int64, device="cuda")
    right = torch.empty(n_words, dtype=torch.int64, device="cuda")
    lo = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    ro = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    sscr = torch.empty(ctx.lib.bn_split_packed_scratch_bytes(n_reads), dtype=torch.uint8, device="cuda")
    sst = dv.SplitStatus("cuda")

    def split():
        dv.raise_for(ctx.lib.bn_split_packed_batch_dev(ctx.handle, dv._stream(), dv._ptr(words), dv._ptr(wo), dv._ptr(d_lens), dv._ptr(idx),
                                                       n_reads, dv._ptr(left), dv._ptr(lo), dv._ptr(right), dv._ptr(ro), dv._ptr(sst.word),
                                                       dv._ptr(sscr)))

    ms = timed(split, reps)
    sst.check(d_lens, idx)
    assert int(lo[-1].item()) == n_reads and int(ro[-1].item()) == n_words
    # the left halves decode to the first 26 bases of each read
    for ridx in (0, 7, n_reads - 1):
        got = dv.decode(left[ridx : ridx + 1].contiguous(), 26)
        o = int(offsets[ridx])
        assert torch.equal(got, data[o : o + 26])
    # in: words + offsets + lens + idx; out: left (1 word) + right (same words) + two offsets
    report(f"split_packed idx=26 reads={n_reads}", ms, 8 * n_words + 24 * n_reads + 8 * n_reads + 8 * n_words + 16 * n_reads, n_reads, "reads")


def next_rows(scale, reps):
    """every 31-mer of one sequence (seq.windows(31) -> as_2bit), and slice gathers
    (10 M windows of 50 bases out of 10 M x 150 bp packed reads)."""
    n = int(500_000_000 * scale)
    asc = dv.synth_ascii(SEED, 7, 0, n)
    out = torch.empty(n - 30, dtype=torch.int64, device="cuda")
    st = dv.Status("cuda")
    ms = timed(lambda: dv.kmers(asc, 31, out=out, status=st), reps)
    st.check()
    words = dv.synth_words(SEED, 7, 0, dv.words_for(n))
    pos = torch.tensor([0, 1, 31, 32, 33, 12345, n - 31], device="cuda")
    for p_ in pos.tolist():  # window p = bits [2p, 2p+62) of the packed stream
        lo = (int(words[p_ // 32].item()) & M64) >> (2 * (p_ % 32))
        hi = ((int(words[p_ // 32 + 1].item()) & M64) << (64 - 2 * (p_ % 32))) & M64 if p_ % 32 and p_ // 32 + 1 < words.numel() else 0
        assert (int(out[p_].item()) & M64) == ((lo | hi) & ((1 << 62) - 1)), p_
    report(f"kmers (all 31-mers of one sequence) n={n}", ms, n + 8 * (n - 30), n - 30, "kmers")
    del asc, out, words

    # per-read windows: 20 M reads of 100-151 bp, k = 31 (windows never cross a read)
    n_reads = int(20_000_000 * scale)
    lens = (100 + synth.splitmix64(np.arange(n_reads, dtype=np.uint64) + np.uint64(SEED + 6)) % np.uint64(52)).astype(np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    total = int(offsets[-1])
    data = dv.synth_ascii(SEED, 6, 0, total)
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    n_win = total - 30 * n_reads
    kout = torch.empty(n_win, dtype=torch.int64, device="cuda")
    koo = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    ctx = dv.api.default_context(0)
    kscr = torch.empty(ctx.lib.bn_kmers_batch_scratch_bytes(n_reads, total), dtype=torch.uint8, device="cuda")
    kst = dv.Status("cuda")

    def run_k():
        dv.raise_for(ctx.lib.bn_kmers_batch_dev(ctx.handle, dv._stream(), dv._ptr(data), dv._ptr(d_off), n_reads, total, 31, dv._ptr(kout),
                                                dv._ptr(koo), dv._ptr(kst.word), dv._ptr(kscr)))

    ms = timed(run_k, reps)
    kst.check()
    assert int(koo[-1].item()) == n_win
    for r in (0, 3, n_reads - 1):   # against the single-sequence kernel
        o, ln, w0 = int(offsets[r]), int(lens[r]), int(koo[r].item())
        ref, st1 = dv.kmers(data[o : o + ln], 31)
        st1.check()
        assert torch.equal(kout[w0 : w0 + ln - 30], ref)
    report(f"kmers per read (k=31) reads={n_reads} x 100-151 bp", ms, total + 8 * n_win + 16 * n_reads, n_win, "kmers")
    del data, kout, koo, kscr

    reads = int(10_000_000 * scale)
    words = dv.synth_words(SEED, 4, 0, 5 * reads)
    wo = torch.arange(reads, dtype=torch.int64, device="cuda") * 5
    lens = torch.full((reads,), 150, dtype=torch.int64, device="cuda")
    qr = torch.arange(reads, dtype=torch.int64, device="cuda")
    qs = (qr * 7919) % 101
    qe = qs + 50
    data = torch.empty(50 * reads, dtype=torch.uint8, device="cuda")
    oo = torch.empty(reads + 1, dtype=torch.int64, device="cuda")
    scratch = torch.empty(ctx.lib.bn_slice_batch_scratch_bytes(reads), dtype=torch.uint8, device="cuda")
    qst = dv.QueryStatus("cuda")

    def run():
        dv.raise_for(ctx.lib.bn_slice_batch_dev(ctx.handle, dv._stream(), dv._ptr(words), dv._ptr(wo), dv._ptr(lens), reads, dv._ptr(qr),
                                                dv._ptr(qs), dv._ptr(qe), reads, dv._ptr(data), dv._ptr(oo), dv._ptr(qst.word), dv._ptr(scratch)))

    ms = timed(run, reps)
    assert qst.first_failing() is None and int(oo[-1].item()) == 50 * reads
    full = dv.decode(words[:5 * 1000].contiguous(), 160 * 1000).view(1000, 160)   # 150 bases + 10 slots of padding per read
    for r in (0, 1, 999):
        s0 = int(qs[r].item())
        assert torch.equal(data[50 * r : 50 * r + 50], full[r, s0 : s0 + 50])
    # in: the touched part of each read (~16 B) + 16 B read index arrays + 24 B query arrays; out: 50 B + 8 B offsets
    report(f"slice gathers: {reads} windows of 50 bases out of 150 bp reads", ms, (16 + 16 + 24 + 50 + 8) * reads, reads, "queries")


def fastq(scale, reps, cpu_port=None):
    """SURVEY.md 8(f) row 3, first half: FASTQ text resident in HBM -> record offsets -> per-read packed words
    (count + index + encode), on 20 M x 150 bp reads (23-byte header lines, so every alignment occurs) and on
    10 kbp reads.  Checked against bn_encode_batch_dev of the same sequences."""
    ctx = dv.api.default_context(0)
    for n_reads, rl, hdr, fasta in ((int(20_000_000 * scale), 150, 23, False), (int(300_000 * scale), 10_000, 37, False),
                                    (int(3_000_000 * scale), 1_000, 37, False),
                                    (int(20_000_000 * scale), 150, 23, True)):
        rec = hdr + rl + 1 if fasta else hdr + rl + 1 + 2 + rl + 1
        seqs = dv.synth_ascii(SEED, 7, 0, n_reads * rl).view(n_reads, rl)
        text2 = torch.full((n_reads, rec), ord("I"), dtype=torch.uint8, device="cuda")
        text2[:, 0] = ord(">") if fasta else ord("@")
        text2[:, 1 : hdr - 1] = ord("h")
        text2[:, hdr - 1] = 10
        text2[:, hdr : hdr + rl] = seqs
        text2[:, hdr + rl] = 10
        if not fasta:
            text2[:, hdr + rl + 1] = ord("+")
            text2[:, hdr + rl + 2] = 10
            text2[:, rec - 1] = 10
        text = text2.view(-1)
        n_bytes = text.numel()
        scratch = torch.empty(ctx.lib.bn_fastq_scratch_bytes(n_bytes), dtype=torch.uint8, device="cuda")
        iscratch = torch.empty(ctx.lib.bn_fastq_index_scratch_bytes(n_reads), dtype=torch.uint8, device="cuda")
        n_lines = torch.zeros(1, dtype=torch.int64, device="cuda")
        so = torch.empty(n_reads, dtype=torch.int64, device="cuda")
        sl = torch.empty(n_reads, dtype=torch.int64, device="cuda")
        wo = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
        wpr = (rl + 31) // 32
        words = torch.empty(n_reads * wpr, dtype=torch.int64, device="cuda")
        st = dv.FastqStatus("cuda")
        P = dv._ptr

        L = ctx.lib
        f_count, f_index, f_encode = (L.bn_fasta_count_dev, L.bn_fasta_index_dev, L.bn_fasta_encode_dev) if fasta else \
            (L.bn_fastq_count_dev, L.bn_fastq_index_dev, L.bn_fastq_encode_dev)

        def count():
            dv.raise_for(f_count(ctx.handle, dv._stream(), P(text), n_bytes, P(scratch), P(n_lines)))

        def index():
            dv.raise_for(f_index(ctx.handle, dv._stream(), P(text), n_bytes, n_reads, P(scratch), P(iscratch), P(so), P(sl), P(wo), P(st.word)))

        def encode():
            dv.raise_for(f_encode(ctx.handle, dv._stream(), P(text), n_bytes, n_reads, P(scratch), P(so), P(sl), P(wo), P(words), P(st.word)))

        def whole():
            count(), index(), encode()

        ms_c, ms_i, ms_e, ms = timed(count, reps), timed(index, reps), timed(encode, reps), timed(whole, reps)
        st.n_lines, st.seq_offsets, st.n_reads, st.fasta = int(n_lines.item()), so, n_reads, fasta
        st.check()
        assert st.n_lines == (2 if fasta else 4) * n_reads and int(wo[-1].item()) == n_reads * wpr
        assert torch.equal(sl, torch.full_like(sl, rl)) and torch.equal(so, torch.arange(n_reads, device="cuda") * rec + hdr)
        offs = torch.arange(n_reads + 1, dtype=torch.int64, device="cuda") * rl
        ref_words, ref_wo, _, bst = dv.encode_batch(seqs.reshape(-1), offs)
        bst.check()
        assert torch.equal(ref_wo, wo) and torch.equal(ref_words[: n_reads * wpr], words)
        del ref_words, ref_wo
        alg = n_bytes + 8 * n_reads * wpr + 24 * n_reads
        extra = {}
        if cpu_port is not None:
            # the CPU form of this row beside it (tests/test_gpu_full_size.py passes the oracle's reader + per-record AVX2
            # encode, the caller's loop of README.md:160-180): one host thread, a bounded sample of the same text
            sample_reads = min(n_reads, max(1, 60_000_000 // rec))
            cpu_s, cpu_w, _ = cpu_port(text[: sample_reads * rec].cpu().numpy(), fasta)
            assert np.array_equal(cpu_w.view(np.int64), words[: sample_reads * wpr].cpu().numpy())
            extra = {"cpu_port_1thread_Gbases_s": round(sample_reads * rl / cpu_s / 1e9, 3),
                     "cpu_sample": f"{sample_reads} reads, {sample_reads * rec} bytes of text"}
        extra |= {"count_ms": round(ms_c, 4), "index_ms": round(ms_i, 4), "encode_ms": round(ms_e, 4), "text_bytes": n_bytes,
                 "text_GB/s": round(n_bytes / (ms * 1e-3) / 1e9, 1)}
        report(f"{'fasta' if fasta else 'fastq'} scan+encode reads={n_reads} x {rl} bp", ms, alg, n_reads * rl, "bases", extra)
        del text, text2, seqs, words, scratch, iscratch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="cfg2,cfg3,cfg4,cfg5")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    print(json.dumps({"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak(), "scale": args.scale}), flush=True)
    for name in args.only.split(","):
        {"cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "short": short_reads, "next": next_rows, "fastq": fastq}[name](args.scale, args.reps)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
