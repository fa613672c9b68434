"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the host-side mirror reproduces the reference's error vocabulary and PackedSequence bit
pokes, and the product path refuses to run without the CUDA library / a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "bitnuc_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(bn_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    import bitnuc_b200._lib as L
    syms = declared_symbols()
    assert len(syms) >= 35
    assert syms == set(L.PROTOTYPES), syms ^ set(L.PROTOTYPES)
    lib = L.load()
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.bn_abi_version() == 2
    assert lib.bn_device_count() >= 0
    assert lib.bn_encode_batch_scratch_bytes(1, 100) >= 16
    assert C.sizeof(L.BnError) == 56  # layout of bn_error_t in the header


def test_error_strings_match_reference(kats):
    import bitnuc_b200 as bn
    import bitnuc_b200._lib as L
    codes = {v: i + 1 for i, v in enumerate(bn.NucleotideError.VARIANTS)}
    for k in kats["error_display"]:
        assert str(bn.NucleotideError(*k["error"])) == k["text"], k["src"]
        payload = (k["error"][1:] + [0, 0, 0])[:3]
        e = L.BnError(code=codes[k["error"][0]], base=payload[0] & 0xFF, a=payload[0], b=payload[1], c=payload[2])
        assert L.error_string(e) == k["text"]
        with pytest.raises(bn.NucleotideError) as ei:
            L.raise_for(codes[k["error"][0]], e)
        assert ei.value.key() == tuple(k["error"])
    assert bn.NucleotideError.InvalidBase(78) == bn.NucleotideError("InvalidBase", 78)
    assert bn.NucleotideError.InvalidBase(78) != bn.NucleotideError.InvalidLength(78)
    with pytest.raises(bn.ReferencePanic):
        L.raise_for(L.BN_ERR_EMPTY_ENCODE)


def _manual_sequence(words, length):
    import bitnuc_b200 as bn
    ps = object.__new__(bn.PackedSequence)
    ps.data, ps.length, ps._ctx = np.array(words, dtype=np.uint64), length, None
    return ps


def test_packed_sequence_host_side_bit_pokes(kats, to_int):
    import bitnuc_b200 as bn
    import oracle
    p = kats["packed_sequence"]
    for k in p["slice"]:
        ps = _manual_sequence(oracle.encode_alloc(k["seq"].encode()), len(k["seq"]))
        assert ps.slice(k["start"], k["end"]) == k["out"].encode(), k["src"]
    for k in p["slice_errors"]:
        ps = _manual_sequence(oracle.encode_alloc(k["seq"].encode()), len(k["seq"]))
        with pytest.raises(bn.NucleotideError) as ei:
            ps.slice(k["start"], k["end"])
        assert ei.value.key() == tuple(k["error"])
    ps = _manual_sequence(oracle.encode_alloc(p["get"]["seq"].encode()), len(p["get"]["seq"]))
    assert bytes(ps.get(i) for i in range(len(ps))) == p["get"]["values"].encode()
    with pytest.raises(bn.NucleotideError) as ei:
        ps.get(p["get_oob"]["index"])
    assert ei.value.key() == tuple(p["get_oob"]["error"])
    rng = np.random.default_rng(0)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1000)]
    ref = oracle.PackedSequence(seq.tobytes())
    ps = _manual_sequence(ref.data, 1000)
    for s, e in [(0, 1000), (31, 33), (999, 1000), (500, 500), (64, 96)]:
        assert ps.slice(s, e) == ref.slice(s, e)
    assert ps == _manual_sequence(ref.data, 1000) and hash(ps) == hash(_manual_sequence(ref.data, 1000))
    assert ps != _manual_sequence(ref.data, 999)
    assert not ps.is_empty() and ps.len() == 1000


def test_product_path_has_no_cpu_fallback_and_never_touches_the_oracle():
    import torch
    import bitnuc_b200 as bn
    for f in (ROOT / "bitnuc_b200").rglob("*"):
        if f.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = f.read_text()
            for needle in ("import oracle", "from oracle", "libbitnuc_oracle", "orc_", "bitnuc_oracle"):
                assert needle not in text, (f, needle)  # no import, link or call of the checker
    if not torch.cuda.is_available():
        for call in (lambda: bn.as_2bit(b"ACGT"), lambda: bn.encode_alloc(b"ACGT"), lambda: bn.PackedSequence(b"ACGT"),
                     lambda: bn.hdist_scalar(0, 1, 1)):
            with pytest.raises(bn.BitnucCudaError):
                call()


def test_sharding_helpers():
    from bitnuc_b200 import sharding as sh
    for n in [0, 1, 63, 64, 65, 1000, 10**9, (1 << 30) + 17]:
        for world in [1, 2, 3, 4, 8]:
            cuts = [sh.shard_bases(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for r in range(world):
                assert cuts[r][0] <= cuts[r][1]
                if r:
                    assert cuts[r][0] == cuts[r - 1][1]
                    assert cuts[r][0] % 64 == 0 or cuts[r][0] == n
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 128 or n < 64 * world
    lens = np.array([5, 0, 100, 7, 7, 7, 300, 1, 0, 64])
    off = np.concatenate([[10], 10 + np.cumsum(lens)]).astype(np.uint64)
    for world in [1, 2, 3, 8, 16]:
        cuts = sh.shard_reads_by_volume(off, world)
        assert cuts[0][0] == 0 and cuts[-1][1] == len(lens)
        assert all(cuts[g][1] == cuts[g + 1][0] for g in range(world - 1))
    assert sh.gc_from_counts([2, 1, 0, 0]) == (1.0 / 3.0) * 100.0
    assert sh.gc_from_counts([0, 0, 0, 0]) == 0.0
    assert sh.first_error_across_ranks(None, 0) is None
    assert sh.first_error_across_ranks((5 << 8) | 78, 1000) == (1005, 78)
