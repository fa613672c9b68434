"""GPU parity of the FASTQ path through the C ABI: bn_fastq_scan / bn_fastq_encode (host pointers)
and the three *_dev calls (device pointers) against the oracle's reader + per-record encode, on the hand-written
cases, on random texts with adversarial shapes (reads across tile boundaries, CRLF, empties, long reads), and on
every fault / invalid-base position class."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from test_oracle_fastq import make_fastq

pytestmark = pytest.mark.gpu

CASES = json.loads((Path(__file__).parent / "golden" / "fastq_cases.json").read_text())
ENC_TILE = 65536


@pytest.fixture(scope="module")
def bn():
    import bitnuc_b200
    return bitnuc_b200


@pytest.fixture(scope="module")
def dv():
    from bitnuc_b200 import device
    return device


def expected(text: bytes, fasta: bool = False):
    """('ok', words, wo, so, sl) | ('fault', record, kind) | ('base', byte, record, position, offset)"""
    try:
        starts, lens = oracle.fastq_scan(text, fasta)
    except oracle.FastqFault as e:
        return ("fault", e.record, e.fault)
    t = np.frombuffer(text, dtype=np.uint8)
    words, offs = [], [0]
    for r, (s, l) in enumerate(zip(starts.tolist(), lens.tolist())):
        if l:
            try:
                words.append(oracle.encode_np(t[s : s + l]))
            except oracle.OracleError as e:
                seq = text[s : s + l]
                pos = next(i for i, b in enumerate(seq) if b not in b"ACGTacgt")
                assert e.key() == ("InvalidBase", seq[pos])
                return ("base", seq[pos], r, pos, s + pos)
        offs.append(offs[-1] + (l + 31) // 32)
    w = np.concatenate(words) if words else np.zeros(0, dtype=np.uint64)
    return ("ok", w, np.asarray(offs, dtype=np.uint64), starts, lens)


def run_host(bn, text: bytes, fasta: bool = False):
    try:
        w, wo, so, sl = (bn.fasta_encode if fasta else bn.fastq_encode)(np.frombuffer(text, dtype=np.uint8))
    except bn.FastqError as e:
        return ("fault", e.record, e.fault)
    except bn.NucleotideError as e:
        assert e.variant == "InvalidBase"
        return ("base", e.payload[0], e.record, e.position, e.offset)
    return ("ok", w, wo, so, sl)


def run_dev(dv, bn, text: bytes, fasta: bool = False):
    import torch
    t = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).cuda() if text else torch.empty(0, dtype=torch.uint8, device="cuda")
    w, wo, so, sl, st = (dv.fasta_encode if fasta else dv.fastq_encode)(t)
    try:
        st.check()
    except bn.FastqError as e:
        return ("fault", e.record, e.fault)
    except bn.NucleotideError as e:
        return ("base", e.payload[0], e.record, e.position, e.offset)
    u = lambda x: x.cpu().numpy().view(np.uint64)
    return ("ok", u(w), u(wo), u(so), u(sl))


def same(a, b):
    if a[0] != b[0]:
        return False
    if a[0] != "ok":
        return tuple(int(x) for x in a[1:]) == tuple(int(x) for x in b[1:])
    return all(np.array_equal(np.asarray(x, dtype=np.uint64), np.asarray(y, dtype=np.uint64)) for x, y in zip(a[1:], b[1:]))


def check(bn, dv, text: bytes, fasta: bool = False):
    exp = expected(text, fasta)
    got_h, got_d = run_host(bn, text, fasta), run_dev(dv, bn, text, fasta)
    assert same(got_h, exp), (got_h[:1], exp[:1], got_h[1:] if got_h[0] != "ok" else "", exp[1:] if exp[0] != "ok" else "")
    assert same(got_d, exp), (got_d[:1], exp[:1], got_d[1:] if got_d[0] != "ok" else "", exp[1:] if exp[0] != "ok" else "")
    return exp


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_cases(bn, dv, case):
    exp = check(bn, dv, case["text"].encode(), case.get("fasta", False))
    if "fault" in case:
        assert exp == ("fault", *case["fault"])
    elif exp[0] == "ok":
        assert [[int(s), int(l)] for s, l in zip(exp[3], exp[4])] == case["reads"]


@pytest.mark.parametrize("kind", ["short", "tiny", "empties", "long", "mixed", "giant"])
@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_texts(bn, dv, kind, crlf, seed):
    rng = np.random.default_rng(100 * seed + len(kind) + 7 * crlf)
    n = int(rng.integers(1, 600))
    if kind == "short":
        lens = rng.integers(100, 152, n)
    elif kind == "tiny":
        lens = rng.integers(0, 6, n)
    elif kind == "empties":
        lens = np.where(rng.random(n) < 0.7, 0, rng.integers(1, 80, n))
    elif kind == "long":
        lens = rng.integers(1500, 12000, min(n, 40))
    elif kind == "mixed":
        lens = (rng.pareto(1.1, n) * 60).astype(np.int64) % 40_000
    else:
        lens = rng.integers(1, 200, min(n, 50))
        lens[int(rng.integers(0, lens.size))] = 200_000 + int(rng.integers(0, 70))
    text = make_fastq(rng, lens, crlf=crlf, final_newline=bool(rng.integers(0, 2)), alphabet=b"ACGTacgt")
    exp = check(bn, dv, text)
    assert exp[0] == "ok" and [int(x) for x in exp[4]] == [int(x) for x in lens]


@pytest.mark.parametrize("fasta", [False, True])
@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("mean_len", [300, 600, 1000, 1600, 2500, 5000, 70_000])
def test_every_encode_path_by_record_size(bn, dv, mean_len, crlf, fasta):
    """The launcher picks the encode kernel from the average record size: a thread per read, 8 / 16 / 32 lanes per read (chunks of
    16 / 32 / 64 words), the giant-read queue above 2^20 bases.  Reads of mixed lengths around each mean (so chunks end ragged, on
    a chunk boundary, one word past it), then one invalid byte at a time at the positions a chunked walk can get wrong."""
    rng = np.random.default_rng(mean_len + crlf + 2 * fasta)
    n = max(6, 200_000 // mean_len)
    lens = np.maximum(1, (mean_len * (0.5 + rng.random(n))).astype(np.int64))
    for k, extra in enumerate((0, 1, 31, 32, 33, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049)):   # around every chunk size
        lens[k % n] = max(1, (mean_len // 2048) * 2048 + extra)
    if mean_len == 70_000:
        lens[n // 2] = (1 << 20) + 77   # one read for the whole-grid queue
    text = make_fastq(rng, lens, crlf=crlf, final_newline=bool(rng.integers(0, 2)), alphabet=b"ACGTacgt", fasta=fasta)
    exp = check(bn, dv, text, fasta)
    assert exp[0] == "ok" and [int(x) for x in exp[4]] == [int(x) for x in lens]
    starts = [int(x) for x in exp[3]]
    for r in (0, n // 2, n - 1):
        s0, ln = starts[r], int(lens[r])
        for pos in sorted({0, 1, 15, 16, 31, 32, 511, 512, 1023, 1024, 2047, 2048, ln - 33, ln - 32, ln - 2, ln - 1}):
            if 0 <= pos < ln:
                bad = bytearray(text)
                bad[s0 + pos] = ord("N")
                got = check(bn, dv, bytes(bad), fasta)
                assert got == ("base", ord("N"), r, pos, s0 + pos), (r, pos, got[:5])


@pytest.mark.parametrize("mix", ["all_dense", "dense_then_normal", "normal_then_dense"])
def test_dense_lines_take_the_fallback_index(bn, dv, mix):
    """More than 2048 lines in a 16 KiB tile (average line under 8 bytes) overflow the slot rows: the dense index runs."""
    rng = np.random.default_rng(len(mix))
    dense = rng.integers(0, 4, 6000)
    normal = rng.integers(100, 152, 300)
    lens = {"all_dense": dense, "dense_then_normal": np.concatenate([dense, normal]),
            "normal_then_dense": np.concatenate([normal, dense])}[mix]
    text = make_fastq(rng, lens, alphabet=b"ACGTacgt")
    assert text.count(b"\n") * 8 > len(text) or mix != "all_dense"
    exp = check(bn, dv, text)
    assert exp[0] == "ok" and [int(x) for x in exp[4]] == [int(x) for x in lens]
    bad = bytearray(text)
    sep = int(exp[3][-1]) + int(exp[4][-1]) + 1      # the last record's separator line (quality bytes may hold a '+' too)
    assert bad[sep] == ord("+")
    bad[sep] = ord("-")
    assert check(bn, dv, bytes(bad)) == ("fault", len(lens) - 1, 2)


def _text_with_read_at(rng, start: int, length: int, before: int = 3, after: int = 3):
    """A text whose read `before` has its sequence line starting exactly at byte `start`."""
    head = make_fastq(rng, rng.integers(20, 60, before))
    pad = start - len(head) - 1           # header line "@" + pad chars + "\n" puts the sequence at `start`
    assert pad >= 1
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = al[rng.integers(0, 4, length)].tobytes()
    rec = b"@" + b"h" * (pad - 1) + b"\n" + seq + b"\n+\n" + b"I" * length + b"\n"
    text = head + rec + make_fastq(rng, rng.integers(20, 60, after))
    assert text[start : start + length] == seq
    return text


@pytest.mark.parametrize("delta", [-40, -33, -32, -17, -16, -15, -1, 0, 1, 15, 16, 17])
@pytest.mark.parametrize("length", [1, 31, 32, 33, 47, 48, 49, 150, 5000, 70_000])
def test_reads_around_a_tile_boundary(bn, dv, delta, length):
    rng = np.random.default_rng(abs(delta) * 131 + length)
    text = _text_with_read_at(rng, ENC_TILE + delta, length)
    exp = check(bn, dv, text)
    assert exp[0] == "ok"


@pytest.mark.parametrize("where", ["first", "head_partial", "interior", "tail_partial", "last", "spill", "spill_last"])
@pytest.mark.parametrize("byte", [ord("N"), 0, 255, ord("\r") + 1, ord("@")])
def test_invalid_base_positions(bn, dv, where, byte):
    rng = np.random.default_rng(len(where) * 17 + byte)
    length = 400 if not where.startswith("spill") else 3000
    start = ENC_TILE - (200 if where.startswith("spill") else 5000) + 5      # misaligned on purpose
    text = bytearray(_text_with_read_at(rng, start, length))
    pos = {"first": 0, "head_partial": 3, "interior": 201, "tail_partial": length - 2, "last": length - 1,
           "spill": 1500, "spill_last": length - 1}[where]
    text[start + pos] = byte
    exp = check(bn, dv, bytes(text))
    assert exp[0] == "base" and exp[1:] == (byte, 3, pos, start + pos)


def test_first_invalid_base_in_file_order_wins(bn, dv):
    rng = np.random.default_rng(9)
    lens = rng.integers(50, 300, 2000)
    text = bytearray(make_fastq(rng, lens))
    starts, _ = oracle.fastq_scan(bytes(text))
    hits = sorted(int(x) for x in rng.choice(2000, 30, replace=False))
    for r in hits:
        text[int(starts[r]) + int(rng.integers(0, lens[r]))] = ord("N")
    exp = check(bn, dv, bytes(text))
    assert exp[0] == "base" and exp[2] == hits[0]


@pytest.mark.parametrize("fault", [1, 2, 3])
def test_faults_deep_inside_a_text(bn, dv, fault):
    rng = np.random.default_rng(fault)
    lens = rng.integers(80, 200, 3000)
    text = bytearray(make_fastq(rng, lens))
    starts, _ = oracle.fastq_scan(bytes(text))
    for r in (2500, 1700):      # two faulty records: the earlier one is reported
        s, l = int(starts[r]), int(lens[r])
        if fault == 1:
            text[text.rfind(b"@", 0, s)] = ord("x")
        elif fault == 2:
            text[s + l + 1] = ord("-")
        else:                    # drop one quality byte (the later record first, so the earlier offsets stay valid)
            del text[s + l + 3]
    # an invalid base earlier in the file does not mask the format fault
    text[int(starts[10]) + 5] = ord("N")
    exp = check(bn, dv, bytes(text))
    assert exp == ("fault", 1700, fault)


def test_truncated_text_and_large_text(bn, dv):
    rng = np.random.default_rng(5)
    lens = rng.integers(100, 152, 60_000)          # ~19 MB of text: several hundred tiles
    text = make_fastq(rng, lens)
    exp = check(bn, dv, text)
    assert exp[0] == "ok" and exp[2][-1] == sum((int(l) + 31) // 32 for l in lens)
    cut = text[: len(text) - 40]                   # ends inside the last record's quality line: lengths differ
    assert check(bn, dv, cut)[0] == "fault"
    cut = text[: text.rfind(b"+")]                 # ends before the last separator
    assert check(bn, dv, cut) == ("fault", 59_999, 4)


def test_fastq_output_feeds_the_packed_domain_calls(bn):
    """The layout bn_fastq_encode produces (words, word offsets, lengths) is the one the packed-domain batch calls take:
    per-read base counts / GC, split at a barcode boundary and slices run on it directly, and agree with the oracle's
    PackedSequence of every record."""
    rng = np.random.default_rng(21)
    lens = rng.integers(30, 200, 400)
    text = make_fastq(rng, lens, alphabet=b"ACGTacgt")
    w, wo, so, sl = bn.fastq_encode(np.frombuffer(text, dtype=np.uint8))
    seqs = [text[int(s) : int(s) + int(l)] for s, l in zip(so, sl)]
    counts, gc, totals = bn.base_counts_batch(w, wo, sl)
    for r in (0, 1, 57, 399):
        ps = oracle.PackedSequence(seqs[r])
        assert [int(x) for x in counts[r]] == ps.base_counts() and gc[r] == ps.gc_content()
    assert sum(totals) == int(lens.sum())
    idx = np.full(400, 16, dtype=np.uint64)
    left, lo, right, ro = bn.split_packed_batch(w, wo, sl, idx)
    for r in (0, 3, 399):
        ol, orr = oracle.split_packed(oracle.encode_alloc(seqs[r]), int(sl[r]), 16)
        assert [int(x) for x in left[int(lo[r]) : int(lo[r + 1])]] == ol and [int(x) for x in right[int(ro[r]) : int(ro[r + 1])]] == orr
    q = np.arange(400, dtype=np.uint64)
    data, oo = bn.slice_batch(w, wo[:-1], sl, q, np.full(400, 5, dtype=np.uint64), np.full(400, 25, dtype=np.uint64))
    assert data.tobytes() == b"".join(s[5:25].upper() for s in seqs)


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_text_reassembles_to_the_whole(bn, world):
    """Multi-GPU shape on one device: the text is cut on record boundaries (sharding.shard_fastq_text), every shard goes
    through bn_fastq_* on its own, offsets are rebased by the totals of the shards before it."""
    from bitnuc_b200 import sharding as sh
    rng = np.random.default_rng(world)
    text = make_fastq(rng, rng.integers(0, 400, 3000), alphabet=b"ACGTacgt")
    w, wo, so, sl = bn.fastq_encode(np.frombuffer(text, dtype=np.uint8))
    parts_w, parts_wo, parts_so, parts_sl, word_base = [], [], [], [], 0
    for lo, hi in sh.shard_fastq_text(text, world):
        pw, pwo, pso, psl = bn.fastq_encode(np.frombuffer(text[lo:hi], dtype=np.uint8))
        parts_w.append(pw.copy())
        parts_wo.append(pwo[:-1] + np.uint64(word_base))
        parts_so.append(pso + np.uint64(lo))
        parts_sl.append(psl.copy())
        word_base += int(pwo[-1])
    assert np.array_equal(np.concatenate(parts_w), w) and np.array_equal(np.concatenate(parts_wo), wo[:-1])
    assert np.array_equal(np.concatenate(parts_so), so) and np.array_equal(np.concatenate(parts_sl), sl)


@pytest.mark.parametrize("size", [16384, 65536, 2 * 65536])
@pytest.mark.parametrize("delta", [-1, 0, 1])
@pytest.mark.parametrize("final_newline", [False, True])
def test_text_lengths_around_tile_multiples(bn, dv, size, delta, final_newline):
    """The last line may lack its newline: the kernels read a virtual one at byte n, which sits in a tile of its own
    when n is a multiple of the tile size."""
    rng = np.random.default_rng(size + delta)
    body = b""
    while len(body) < size - 1500:                                       # whole records, well short of the target
        body += make_fastq(rng, rng.integers(50, 150, 1))
    tail_len = 100
    fixed = len(body) + 1 + 1 + tail_len + 3 + tail_len + (1 if final_newline else 0)   # '@' + header pad + '\n' ...
    pad = size + delta - fixed
    assert pad >= 0
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), tail_len))
    text = body + b"@" + b"p" * pad + b"\n" + seq + b"\n+\n" + b"I" * tail_len + (b"\n" if final_newline else b"")
    assert len(text) == size + delta
    exp = check(bn, dv, text)
    assert exp[0] == "ok" and int(exp[4][-1]) == tail_len


@pytest.mark.parametrize("seed", range(8))
def test_mutated_texts_fuzz(bn, dv, seed):
    """Random damage to valid texts -- a byte turned into a newline / '@' / '+' / '\\r' / junk, a byte deleted, the text cut,
    a line duplicated -- must give the oracle's answer whatever it is: the same records, the same first fault, or the same
    first invalid base."""
    rng = np.random.default_rng(500 + seed)
    outcomes = set()
    for _ in range(25):
        kind = int(rng.integers(0, 3))
        lens = [rng.integers(0, 40, int(rng.integers(1, 30))), rng.integers(100, 152, int(rng.integers(1, 400))),
                rng.integers(1, 6000, int(rng.integers(1, 20)))][kind]
        text = bytearray(make_fastq(rng, lens, crlf=bool(rng.integers(0, 2)), final_newline=bool(rng.integers(0, 2)), alphabet=b"ACGTacgt"))
        for _ in range(int(rng.integers(1, 4))):
            if not text:
                break
            i = int(rng.integers(0, len(text)))
            k = int(rng.integers(0, 8))
            if k == 0:
                text[i] = 10
            elif k == 1:
                text[i] = ord("@")
            elif k == 2:
                text[i] = ord("+")
            elif k == 3:
                text[i] = 13
            elif k == 4:
                text[i] = int(rng.integers(0, 256))
            elif k == 5:
                del text[i]
            elif k == 6:
                del text[i:]
            else:
                e = text.find(b"\n", i)
                if e >= 0:
                    text[e + 1 : e + 1] = text[text.rfind(b"\n", 0, i) + 1 : e + 1]
        outcomes.add(check(bn, dv, bytes(text))[0])
    assert outcomes   # usually all three of ok / fault / base


# ---- FASTA with one sequence line per record: the same kernels with two-line records -------------------------------

@pytest.mark.parametrize("kind", ["short", "tiny", "long", "mixed"])
@pytest.mark.parametrize("crlf", [False, True])
def test_fasta_random_texts(bn, dv, kind, crlf):
    rng = np.random.default_rng(len(kind) + 3 * crlf)
    n = int(rng.integers(1, 800))
    lens = {"short": rng.integers(100, 152, n), "tiny": rng.integers(0, 6, 8 * n), "long": rng.integers(1500, 40000, min(n, 30)),
            "mixed": (rng.pareto(1.1, n) * 60).astype(np.int64) % 90_000}[kind]
    text = make_fastq(rng, lens, crlf=crlf, final_newline=bool(rng.integers(0, 2)), alphabet=b"ACGTacgt", fasta=True)
    exp = check(bn, dv, text, fasta=True)
    assert exp[0] == "ok" and [int(x) for x in exp[4]] == [int(x) for x in lens]


def test_fasta_faults_and_invalid_bases(bn, dv):
    rng = np.random.default_rng(12)
    lens = rng.integers(50, 3000, 400)
    text = bytearray(make_fastq(rng, lens, alphabet=b"ACGT", fasta=True))
    starts, _ = oracle.fastq_scan(bytes(text), fasta=True)
    bad = bytearray(text)
    bad[int(starts[300]) + 17] = ord("N")
    bad[int(starts[120]) + int(lens[120]) - 1] = ord("n")
    assert check(bn, dv, bytes(bad), fasta=True) == ("base", ord("n"), 120, int(lens[120]) - 1, int(starts[120]) + int(lens[120]) - 1)
    wrapped = bytearray(text)
    wrapped[int(starts[200]) + 20] = 10                    # a sequence broken over two lines: the next "record" has no '>'
    assert check(bn, dv, bytes(wrapped), fasta=True) == ("fault", 201, 1)
    assert check(bn, dv, bytes(text[: int(starts[399])]), fasta=True) == ("fault", 399, 4)
    assert check(bn, dv, bytes(text), fasta=False)[0] == "fault"      # FASTA is not FASTQ


@pytest.mark.parametrize("seed", range(4))
def test_fasta_mutated_texts_fuzz(bn, dv, seed):
    rng = np.random.default_rng(900 + seed)
    for _ in range(25):
        lens = [rng.integers(0, 40, int(rng.integers(1, 60))), rng.integers(100, 152, int(rng.integers(1, 400))),
                rng.integers(1, 30000, int(rng.integers(1, 12)))][int(rng.integers(0, 3))]
        text = bytearray(make_fastq(rng, lens, crlf=bool(rng.integers(0, 2)), final_newline=bool(rng.integers(0, 2)), alphabet=b"ACGTacgt",
                                    fasta=True))
        for _ in range(int(rng.integers(0, 3))):
            if not text:
                break
            i = int(rng.integers(0, len(text)))
            k = int(rng.integers(0, 5))
            if k == 0:
                text[i] = 10
            elif k == 1:
                text[i] = ord(">")
            elif k == 2:
                text[i] = int(rng.integers(0, 256))
            elif k == 3:
                del text[i]
            else:
                del text[i:]
        check(bn, dv, bytes(text), fasta=True)


def test_device_text_that_is_16_but_not_32_byte_aligned(bn, dv):
    """The thread-per-read encode uses 256-bit loads when the text itself is 32-byte aligned and falls back to 128-bit ones when it
    is only 16-byte aligned (the documented requirement): both against the oracle, on reads that end at the end of the text."""
    import torch
    rng = np.random.default_rng(77)
    recs = []
    for r in range(3000):
        n = int(rng.integers(1, 300))
        seq = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, n)].tobytes()
        recs.append(b"@r%d\n%s\n+\n%s\n" % (r, seq, b"I" * n))
    text = b"".join(recs)[:-1]                                   # no final newline: the last quality line ends the text
    exp = expected(text)
    for shift in (0, 16):
        buf = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
        view = buf[shift: shift + len(text)]
        view.copy_(torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()))
        assert view.data_ptr() % 32 == shift
        w, wo, so, sl, st = dv.fastq_encode(view)
        st.check()
        u = lambda x: x.cpu().numpy().view(np.uint64)
        assert same(("ok", u(w), u(wo), u(so), u(sl)), exp), shift
