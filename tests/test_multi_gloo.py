"""The N > 1 host logic on CPU: two gloo ranks shard one sequence / one read batch the way bench.py and
the multi-GPU drivers do, run the per-shard work (the oracle stands in for the kernels: there is no GPU
here), and combine with the same collectives -- SUM of the four base counters, SUM of the hdist
partials, MIN of the first invalid offset.  The combined results must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import oracle_np as onp

WORLD = 2
N_BASES = 100_003


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from bitnuc_b200 import sharding as sh
    try:
        seed = oracle.DEFAULT_SEED
        lo, hi = sh.shard_bases(N_BASES, rank, WORLD)
        assert lo % 64 == 0
        # each rank generates only its own shard (counter-based stream) and encodes it
        shard = oracle.synth_ascii(seed, 0, lo, hi - lo)
        words = oracle.encode_np(shard)
        other = oracle.encode_np(oracle.synth_ascii(seed, 3, lo, hi - lo))
        counts = torch.tensor(oracle.base_counts(words, hi - lo), dtype=torch.int64)
        sh.allreduce_counts(counts)
        hd = torch.tensor([oracle.hdist(words, other, hi - lo, wide=True)], dtype=torch.int64)
        sh.allreduce_sum(hd)
        # error parity: rank 1 sees an invalid base early in its shard, rank 0 a later one in its own
        bad_local = {0: (hi - lo - 5, ord("X")), 1: (7, ord("N"))}[rank]
        first = sh.first_error_across_ranks((bad_local[0] << 8) | bad_local[1], lo)
        none = sh.first_error_across_ranks(None, lo)
        only1 = sh.first_error_across_ranks(((3 << 8) | ord("Q")) if rank == 1 else None, lo)
        # gathered words reassemble the unsharded encode (shards are cut on word boundaries)
        gathered = [None] * WORLD
        dist.all_gather_object(gathered, words.tolist())
        if rank == 0:
            results.put({"counts": counts.tolist(), "hdist": int(hd.item()), "first": first, "none": none, "only1": only1,
                         "words": sum(gathered, []), "cut": sh.shard_bases(N_BASES, 1, WORLD)[0]})
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = q.get()
    seed = oracle.DEFAULT_SEED
    full = oracle.synth_ascii(seed, 0, 0, N_BASES)
    words = oracle.encode_np(full)
    other = oracle.encode_np(oracle.synth_ascii(seed, 3, 0, N_BASES))
    assert res["words"] == [int(w) for w in words]
    assert res["counts"] == oracle.base_counts(words, N_BASES) == onp.base_counts(words, N_BASES)
    assert res["hdist"] == oracle.hdist(words, other, N_BASES, wide=True)
    cut = res["cut"]
    assert res["first"] == (cut - 5, ord("X"))      # rank 0's error comes first in sequence order
    assert res["none"] is None
    assert res["only1"] == (cut + 3, ord("Q"))
    from bitnuc_b200 import sharding as sh
    assert sh.gc_from_counts(res["counts"]) == oracle.gc_content(words, N_BASES)


def _fastq_worker(rank, port, text, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from bitnuc_b200 import sharding as sh
    try:
        lo, hi = sh.shard_fastq_text(text, WORLD)[rank]
        words, wo, so, sl = oracle.fastq_encode(text[lo:hi])          # the oracle stands in for bn_fastq_* (no GPU here)
        sizes = torch.tensor([sl.size, words.size], dtype=torch.int64)
        gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(gathered, sizes)
        read_base = sum(int(g[0]) for g in gathered[:rank])
        word_base = sum(int(g[1]) for g in gathered[:rank])
        parts = [None] * WORLD
        dist.all_gather_object(parts, {"words": words.tolist(), "wo": [int(x) + word_base for x in wo[:-1]],
                                       "so": [int(x) + lo for x in so], "sl": sl.tolist(), "read_base": read_base})
        if rank == 0:
            results.put(parts)
    finally:
        dist.destroy_process_group()


def test_two_rank_fastq_sharding_matches_the_whole_text():
    from test_oracle_fastq import make_fastq
    rng = np.random.default_rng(8)
    text = make_fastq(rng, rng.integers(0, 300, 500), alphabet=b"ACGTacgt")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_fastq_worker, args=(r, port, text, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    parts = q.get()
    words, wo, so, sl = oracle.fastq_encode(text)
    assert sum((p["words"] for p in parts), []) == [int(x) for x in words]
    assert sum((p["wo"] for p in parts), []) == [int(x) for x in wo[:-1]]
    assert sum((p["so"] for p in parts), []) == [int(x) for x in so]
    assert sum((p["sl"] for p in parts), []) == [int(x) for x in sl]
    assert parts[1]["read_base"] == len(parts[0]["sl"]) > 0


def test_fastq_cuts_fall_on_record_boundaries():
    from bitnuc_b200 import sharding as sh
    from test_oracle_fastq import make_fastq
    rng = np.random.default_rng(2)
    # quality lines full of '@' and '+': the boundary rule must not be fooled by them
    lens = rng.integers(1, 60, 200)
    recs = []
    for r, n in enumerate(lens):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(n)))
        qual = bytes(rng.choice(np.frombuffer(b"@+I", dtype=np.uint8), int(n)))
        recs.append(b"@r%d\n%s\n+\n%s\n" % (r, seq, qual))
    text = b"".join(recs)
    starts = set(np.cumsum([0] + [len(x) for x in recs]).tolist())
    for world in (1, 2, 3, 8, 64):
        shards = sh.shard_fastq_text(text, world)
        assert shards[0][0] == 0 and shards[-1][1] == len(text)
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        assert all(lo in starts for lo, _ in shards)
    for pos in range(0, len(text), 37):
        assert sh.fastq_record_start(text, pos) == min(s for s in starts if s >= pos)
    assert sh.fastq_record_start(text, len(text)) == len(text)
    # a text cut at these points parses shard by shard to the same records
    total = 0
    for lo, hi in sh.shard_fastq_text(text, 5):
        total += oracle.fastq_scan(text[lo:hi])[0].size
    assert total == 200
    assert make_fastq is not None


def test_fasta_cuts_fall_on_record_boundaries():
    from bitnuc_b200 import sharding as sh
    from test_oracle_fastq import make_fastq
    rng = np.random.default_rng(4)
    text = make_fastq(rng, rng.integers(0, 200, 300), crlf=True, fasta=True)
    starts = {0} | {i + 1 for i in range(len(text) - 1) if text[i] == 10 and text[i + 1] == ord(">")}
    for world in (1, 2, 7, 50):
        shards = sh.shard_fasta_text(text, world)
        assert shards[0][0] == 0 and shards[-1][1] == len(text) and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        assert all(lo in starts or lo == len(text) for lo, _ in shards)
        assert sum(oracle.fasta_scan(text[lo:hi])[0].size for lo, hi in shards) == 300
    for pos in range(0, len(text), 53):
        assert sh.fasta_record_start(text, pos) == min([s for s in starts if s >= pos] + [len(text)])
