#!/usr/bin/env python
"""End-to-end (host pointers, PCIe inside the timed region) throughput of the pipelined host-pointer calls on one
GPU, pinned buffers: as_2bit / from_2bit batches (cfg 3 shape), hdist_pairs and hdist (cfg 4 shape), base_counts,
kmers.  The yardstick is the PCIe link (tools/pcie_probe.py): a call is link-bound in its dominant direction."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

import bitnuc_b200 as bn
from bitnuc_b200 import device as dv

SEED = 0x5EEDB17C0DE5
ctx = bn.Context(0)


def pinned(t: torch.Tensor, dtype):
    a = ctx.pinned_empty(t.numel(), dtype)
    a[:] = t.cpu().numpy().view(dtype)
    return a


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def line(name, s, h2d, d2h, units, unit):
    print(json.dumps({"call": name, "ms": round(s * 1e3, 3), "h2d_GB": round(h2d / 1e9, 3), "d2h_GB": round(d2h / 1e9, 3),
                      "h2d_GB_s": round(h2d / s / 1e9, 1), "d2h_GB_s": round(d2h / s / 1e9, 1), f"G{unit}_s": round(units / s / 1e9, 3)}), flush=True)


n = 1 << 26
words = dv.synth_words(SEED, 1, 0, n)
recs = pinned(dv.from_2bit_batch(words, 31), np.uint8)
h_words = pinned(words, np.uint64)
out_w = ctx.pinned_empty(n, np.uint64)
s = timed(lambda: ctx.lib.bn_as_2bit_batch(ctx.handle, recs.ctypes.data, n, 31, 31, out_w.ctypes.data, None))
assert np.array_equal(out_w, h_words & np.uint64((1 << 62) - 1))
line("bn_as_2bit_batch k=31 tight", s, 31 * n, 8 * n, n, "kmers")
out_r = ctx.pinned_empty(31 * n, np.uint8)
s = timed(lambda: ctx.lib.bn_from_2bit_batch(ctx.handle, h_words.ctypes.data, n, 31, out_r.ctypes.data, 31, None))
assert np.array_equal(out_r, recs)
line("bn_from_2bit_batch k=31 tight", s, 8 * n, 31 * n, n, "kmers")
del recs, out_r

n = 1 << 27
u, v = pinned(dv.synth_words(SEED, 2, 0, n), np.uint64), pinned(dv.synth_words(SEED, 3, 0, n), np.uint64)
d = ctx.pinned_empty(n, np.uint32)
s = timed(lambda: ctx.lib.bn_hdist_pairs(ctx.handle, u.ctypes.data, v.ctypes.data, n, 32, d.ctypes.data, None))
line("bn_hdist_pairs len=32", s, 16 * n, 4 * n, n, "pairs")
import ctypes as C
tot = C.c_uint64(0)
s = timed(lambda: ctx.lib.bn_hdist(ctx.handle, u.ctypes.data, n, v.ctypes.data, n, 32 * n, C.byref(tot), None))
assert tot.value == int(d.astype(np.uint64).sum())
line("bn_hdist whole sequence", s, 16 * n, 8, 32 * n, "bases")
counts, gc = (C.c_uint64 * 4)(), C.c_double(0)
s = timed(lambda: ctx.lib.bn_base_counts(ctx.handle, u.ctypes.data, n, 32 * n, counts, C.byref(gc), None))
assert sum(counts) == 32 * n
line("bn_base_counts whole sequence", s, 8 * n, 40, 32 * n, "bases")
del u, v, d

n = 1 << 27
seq = pinned(dv.synth_ascii(SEED, 7, 0, n), np.uint8)
out_k = ctx.pinned_empty(n - 30, np.uint64)
n_out = C.c_size_t(0)
s = timed(lambda: ctx.lib.bn_kmers(ctx.handle, seq.ctypes.data, n, 31, out_k.ctypes.data, C.byref(n_out), None))
assert n_out.value == n - 30
line("bn_kmers k=31", s, n, 8 * (n - 30), n - 30, "kmers")

# per-read base counts + gc on 10 M x 150 bp reads (cfg 4 shape): in-order read tables go through the 3-stage pipeline in chunks of whole reads
n_reads = 10_000_000
h_rw = pinned(dv.synth_words(SEED, 4, 0, 5 * n_reads), np.uint64)
h_wo = ctx.pinned_empty(n_reads + 1, np.uint64)
h_wo[:] = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(5)
h_ln = ctx.pinned_empty(n_reads, np.uint64)
h_ln[:] = 150
h_c4, h_gc = ctx.pinned_empty(4 * n_reads, np.uint64), ctx.pinned_empty(n_reads, np.float64)
tot4 = (C.c_uint64 * 4)()


def bcb():
    assert ctx.lib.bn_base_counts_batch(ctx.handle, h_rw.ctypes.data, 5 * n_reads, h_wo.ctypes.data, h_ln.ctypes.data, n_reads,
                                        h_c4.ctypes.data, h_gc.ctypes.data, tot4, None) == 0


for label, chunk in (("pipelined over chunks of whole reads", 0), ("staged whole", 1 << 40)):
    ctx.set_chunk_bytes(chunk)
    h_c4[:8] = 0
    s = timed(bcb)
    assert sum(tot4) == 150 * n_reads and int(h_c4[:4].sum()) == 150 and int(h_c4[-4:].sum()) == 150
    line(f"bn_base_counts_batch 10 M x 150 bp ({label})", s, 56 * n_reads, 40 * n_reads, n_reads, "reads")
ctx.set_chunk_bytes(0)
del h_rw, h_c4, h_gc

# FASTQ text (2 M x 150 bp, 23-byte headers) -> records -> packed reads: scan + encode, text in pinned / pageable memory
n_reads, rl, hdr = 2_000_000, 150, 23
rec = hdr + rl + 1 + 2 + rl + 1
t2 = torch.full((n_reads, rec), ord("I"), dtype=torch.uint8, device="cuda")
t2[:, 0] = ord("@"); t2[:, 1:hdr - 1] = ord("h"); t2[:, hdr - 1] = 10
t2[:, hdr:hdr + rl] = dv.synth_ascii(SEED, 7, 0, n_reads * rl).view(n_reads, rl)
t2[:, hdr + rl] = 10; t2[:, hdr + rl + 1] = ord("+"); t2[:, hdr + rl + 2] = 10; t2[:, rec - 1] = 10
text_pin = pinned(t2.view(-1), np.uint8)
text_page = np.array(text_pin)
nb = text_pin.size
wpr = (rl + 31) // 32
out_w = ctx.pinned_empty(n_reads * wpr, np.uint64)
wo, so, sl = (ctx.pinned_empty(n_reads + 1, np.uint64) for _ in range(3))


def fastq(text):
    nr, nw = C.c_size_t(0), C.c_size_t(0)
    assert ctx.lib.bn_fastq_scan(ctx.handle, text.ctypes.data, nb, C.byref(nr), C.byref(nw), None) == 0
    assert nr.value == n_reads and nw.value == n_reads * wpr
    assert ctx.lib.bn_fastq_encode(ctx.handle, text.ctypes.data, nb, nr.value, nw.value, out_w.ctypes.data, wo.ctypes.data, so.ctypes.data,
                                   sl.ctypes.data, None) == 0


for name, text in (("pinned", text_pin), ("pageable", text_page)):
    s = timed(lambda: fastq(text))
    assert int(wo[n_reads]) == n_reads * wpr and int(sl[7]) == rl and int(so[1]) == rec + hdr
    line(f"bn_fastq_scan + bn_fastq_encode ({name} text)", s, nb, 8 * n_reads * wpr + 24 * n_reads, n_reads * rl, "bases")

# wrapped FASTA, the genome-file form: 64 records of 8 Mbases wrapped at 60 columns (0.52 GB of text) -> joined sequences -> packed
del t2, text_pin, text_page
n_rec, rec_bases, width = 64, 8_000_040, 60
lines_per = rec_bases // width
body = torch.empty((n_rec, lines_per, width + 1), dtype=torch.uint8, device="cuda")
body[:, :, :width] = dv.synth_ascii(SEED, 8, 0, n_rec * rec_bases).view(n_rec, lines_per, width)
body[:, :, width] = 10
hdr_line = torch.tensor(list(b">chromosome_xx description\n"), dtype=torch.uint8, device="cuda")
wtext = torch.cat([torch.cat([hdr_line, body[r].reshape(-1)]) for r in range(n_rec)])
wtext_pin = pinned(wtext, np.uint8)
wnb = wtext_pin.size
w_words = (rec_bases + 31) // 32
w_out = ctx.pinned_empty(n_rec * w_words, np.uint64)
w_wo, w_ho, w_sl = (ctx.pinned_empty(n_rec + 1, np.uint64) for _ in range(3))


def fasta_wrapped():
    nr, nbase, nw = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    assert ctx.lib.bn_fasta_wrapped_scan(ctx.handle, wtext_pin.ctypes.data, wnb, C.byref(nr), C.byref(nbase), C.byref(nw), None) == 0
    assert (nr.value, nbase.value, nw.value) == (n_rec, n_rec * rec_bases, n_rec * w_words)
    assert ctx.lib.bn_fasta_wrapped_encode(ctx.handle, wtext_pin.ctypes.data, wnb, nr.value, nw.value, w_out.ctypes.data, w_wo.ctypes.data,
                                           w_ho.ctypes.data, w_sl.ctypes.data, None) == 0


s = timed(fasta_wrapped)
assert int(w_sl[3]) == rec_bases and int(w_ho[1]) == hdr_line.numel() + lines_per * (width + 1)
expect = dv.synth_words(SEED, 8, 0, 4).cpu().numpy().view(np.uint64)   # record 0 is the generator's stream 8 from base 0
assert np.array_equal(w_out[:4], expect)
line("bn_fasta_wrapped_scan + bn_fasta_wrapped_encode (60-column genome FASTA, pinned text)", s, wnb, 8 * n_rec * w_words, n_rec * rec_bases, "bases")

# split_packed over 10 M x 150 bp packed reads at base 26 (barcode | insert), pinned buffers (the call stages whole buffers)
del wtext, wtext_pin, body, w_out
n_reads = 10_000_000
s_words = pinned(dv.synth_words(SEED, 4, 0, 5 * n_reads), np.uint64)
s_wo = ctx.pinned_empty(n_reads + 1, np.uint64); s_wo[:] = np.arange(n_reads + 1, dtype=np.uint64) * 5
s_len = ctx.pinned_empty(n_reads, np.uint64); s_len[:] = 150
s_idx = ctx.pinned_empty(n_reads, np.uint64); s_idx[:] = 26
s_left, s_right = ctx.pinned_empty(6 * n_reads, np.uint64), ctx.pinned_empty(5 * n_reads, np.uint64)
s_lo, s_ro = ctx.pinned_empty(n_reads + 1, np.uint64), ctx.pinned_empty(n_reads + 1, np.uint64)


def split():
    assert ctx.lib.bn_split_packed_batch(ctx.handle, s_words.ctypes.data, s_words.size, s_wo.ctypes.data, s_len.ctypes.data, s_idx.ctypes.data,
                                         n_reads, s_left.ctypes.data, s_lo.ctypes.data, s_right.ctypes.data, s_ro.ctypes.data, None) == 0


s = timed(split)
assert int(s_lo[n_reads]) == n_reads and int(s_ro[n_reads]) == 5 * n_reads and int(s_left[0]) == int(s_words[0]) & ((1 << 52) - 1)
line("bn_split_packed_batch 10 M x 150 bp at base 26", s, (5 + 3) * 8 * n_reads, (1 + 5 + 2) * 8 * n_reads, n_reads, "reads")
