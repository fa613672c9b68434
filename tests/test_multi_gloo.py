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
