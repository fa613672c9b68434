"""BASELINE.json configs 2-5 at their FULL sizes on one B200, checked through size-independent properties
(the oracle cannot finish these sizes in seconds, the generator's closed forms can):

  cfg 2  encode(synth_ascii) == synth_words (tail masked) and decode(encode(x)) == x at 10^9 and 2^30+17 bases
  cfg 3  as_2bit(from_2bit(W)) == W & (2^62 - 1) for 2^28 31-mers, tight and padded records
  cfg 4  sum of 2^30 per-pair distances == whole-sequence distance (checksum of checksums); counts sum to n;
         per-read counts sum to the totals, gc == (c+g)/150*100 in the reference's operation order
  cfg 5  ~32 Gbases of 50 bp - 10 kbp reads: word count == sum ceil(len/32), sampled reads round-trip through
         decode, injected N bases reported exactly (first in input order + per-read positions)
  next   every 31-mer of a 0.5 Gbase sequence against the packed stream; slice windows against decode
  fastq  6.5 GB of FASTQ text (2 x 10^7 reads of 150 bp) and 6 GB of 10 kbp reads: line count == 4 x reads, sequence
         offsets / lengths in closed form, packed words == bn_encode_batch of the same sequences

The functions are the ones tools/bench_configs.py times; here they run once each for their assertions.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg():
    import torch
    import bench_configs
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~50 GB of free HBM for the full-size configurations")
    return bench_configs


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg5", "short_reads", "next_rows", "fastq"])
def test_full_size_properties(cfg, name, capsys):
    import torch
    if name == "fastq":   # with the CPU form of the row timed beside it (the oracle's reader + per-record encode, one thread)
        import oracle
        cfg.fastq(1.0, 1, cpu_port=oracle.fastx_encode_timed)
    else:
        getattr(cfg, name)(1.0, 1)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    out = capsys.readouterr().out
    assert '"kernel"' in out
    with capsys.disabled():   # the measured lines (and the CPU form beside the FASTQ row) stay visible in the test log
        print(out, end="")


def test_more_than_2_pow_32_bases(cfg):
    """Index arithmetic past 32 bits: encode / decode / base_counts / hdist on 2^32 + 12345 bases (4.3 GB of ASCII),
    and a read batch holding one read of more than 2^32 bases."""
    import numpy as np
    import torch
    from bitnuc_b200 import device as dv
    n = (1 << 32) + 12345
    asc = dv.synth_ascii(cfg.SEED, 9, 0, n)
    words, st = dv.encode(asc)
    st.check()
    expect = dv.synth_words(cfg.SEED, 9, 0, dv.words_for(n))
    expect[-1] &= (1 << (2 * (n % 32))) - 1
    assert torch.equal(words, expect)
    del expect
    back = dv.decode(words, n)
    assert torch.equal(back, asc)
    del back
    counts, _ = dv.base_counts(words, n)
    assert int(counts.sum().item()) == n
    assert int(dv.hdist(words, words, n).item()) == 0
    asc[n - 7] = ord("N")                                  # an invalid base past offset 2^32
    _, st = dv.encode(asc, out=words)
    with pytest.raises(Exception) as ei:
        st.check()
    assert ei.value.key() == ("InvalidBase", ord("N")) and ei.value.offset == n - 7
    asc[n - 7] = ord("A")
    # the same bytes as a batch of three reads, the middle one longer than 2^32 bases
    offsets = torch.tensor([0, 1000, n - 5, n], dtype=torch.int64, device="cuda")
    bw, bwo, _, bst = dv.encode_batch(asc, offsets)
    bst.check()
    lens = [1000, n - 1005, 5]
    assert bwo.tolist() == [0, 32, 32 + (lens[1] + 31) // 32, 32 + (lens[1] + 31) // 32 + 1]
    mid = dv.decode(bw[32 : 32 + (lens[1] + 31) // 32].contiguous(), lens[1])
    assert torch.equal(mid[:4096], asc[1000:5096]) and torch.equal(mid[-4096:], asc[n - 5 - 4096 : n - 5])
    assert torch.equal(dv.decode(bw[-1:].contiguous(), 5), asc[n - 5 :])


def test_full_size_hdist_pairs_against_the_oracle_on_a_strided_sample(cfg):
    """cfg 4 at its full 2^30 pairs: a strided sample across the WHOLE range (not a prefix) through the oracle's
    hdist_scalar (hamming/scalar.rs:11-48), plus the whole-sequence total of the sampled words through orc_hdist."""
    import numpy as np
    import torch
    import oracle
    from bitnuc_b200 import device as dv
    n = 1 << 30
    u, v = dv.synth_words(cfg.SEED, 2, 0, n), dv.synth_words(cfg.SEED, 3, 0, n)
    out = dv.hdist_pairs(u, v, 32)
    idx = torch.arange(0, n, 4099, device="cuda")          # 262 k pairs, every region of the buffers
    idx = torch.cat([idx, torch.tensor([n - 1], device="cuda")])
    su, sv = (x[idx].cpu().numpy().view(np.uint64) for x in (u, v))
    got = out[idx].cpu().numpy().astype(np.uint32)
    exp = np.zeros(su.size, dtype=np.uint32)
    _, s = oracle.bench_op(oracle.OP_HDIST_PAIRS, su.size, in0=su, in1=sv, out0=exp, k=32, threads=1, reps=1)
    assert np.array_equal(got, exp) and s == int(exp.sum())
    assert oracle.hdist(su, sv, 32 * su.size, wide=True) == s
    for ln in (1, 17, 31):                                   # shorter pair lengths on the same sample
        o2 = dv.hdist_pairs(u[idx].contiguous(), v[idx].contiguous(), ln).cpu().numpy().astype(np.uint32)
        oracle.bench_op(oracle.OP_HDIST_PAIRS, su.size, in0=su, in1=sv, out0=exp, k=ln, threads=1, reps=1)
        assert np.array_equal(o2, exp), ln
