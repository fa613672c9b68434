"""Static checks on the compiled sm_100a code (cuobjdump works without a GPU): the hot kernels use 128-bit global
accesses, keep everything in registers (no local-memory spills), and the library carries sm_100a SASS only --
the hardware mapping DESIGN.md describes, verified on the artefact that ships."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "bitnuc_b200" / "libbitnuc_cuda.so"


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists() or not LIB.exists():
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and "/*" in line:
            funcs[name].append(line)
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    return funcs, arch


def _ops(lines):
    ops = []
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append(m.group(1))
    return ops


def _find(funcs, needle):
    hits = [k for k in funcs if needle in k]
    assert hits, f"no kernel matching {needle}"
    return hits


def test_library_targets_sm_100a_only(sass):
    _, arch = sass
    assert arch == {"sm_100a"}, arch


@pytest.mark.parametrize("needle,loads,stores", [
    ("encode_kernelILi4ELi512", r"LDG\.E(\.NA)?\.128", r"STG\.E"),        # 16 bases per 128-bit load, one 32-bit code out
    ("decode_kernelILi4ELi512", r"LDG\.E", r"STG\.E\.EF\.128"),            # one 32-bit code in, 16 bases per 128-bit streaming store
    ("hdist_pairs_kernel", r"LDG\.E(\.NA)?\.128", r"STG"),
    ("base_counts_kernelEPK5uint4", r"LDG\.E(\.NA)?\.128", None),
    ("hdist_sum_kernel", r"LDG\.E(\.NA)?\.128", None),
    ("encode_batch_kernel", r"LDG\.E(\.NA)?\.128", r"STG\.E\.64"),
    ("kmer_windows_kernel", r"LDG\.E(\.NA)?\.128", r"STG\.E\.EF\.64"),
    ("as_2bit_tight_kernel", r"LDG\.E(\.NA)?\.128", r"STG\.E\.EF\.64"),
    ("from_2bit_tight31_kernel", r"LDG\.E", r"STG\.E\.EF\.128"),            # one or two 64-bit words in, 16 bases per 128-bit streaming store
    ("fastq_lines_kernel", r"LDG\.E(\.NA)?\.128", r"STG\.E"),              # the text in 128-bit loads, 32-bit line entries out
    ("fastq_encode_kernelILi49152ELi128ELi9ELb1", r"LDG\.E(\.NA)?\.128", r"STG\.E\.64"),
    ("fastq_encode_long_kernelILi32", r"LDG\.E(\.NA)?\.128", r"STG\.E\.64"),   # a warp per read: the sequence line in 128-bit loads
    ("fastq_encode_long_kernelILi16", r"LDG\.E(\.NA)?\.128", r"STG\.E\.64"),
    ("fastq_encode_long_kernelILi8", r"LDG\.E(\.NA)?\.128", r"STG\.E\.64"),
    ("split_packed_fused_kernel", r"LDG\.E\.64", r"STS\.64"),                   # the staged destination is written with STS, not generic ST
    ("slice_short_kernel", r"LDG\.E\.64", r"STS"),
])
def test_hot_kernels_use_wide_accesses_and_no_local_memory(sass, needle, loads, stores):
    funcs, _ = sass
    for name in _find(funcs, needle):
        ops = _ops(funcs[name])
        assert any(re.match(loads, o) for o in ops), (name, loads)
        if stores:
            assert any(re.match(stores, o) for o in ops), (name, stores)
        assert not any(o.startswith(("LDL", "STL")) for o in ops), f"{name} spills to local memory"


def test_encode_stays_within_the_instruction_budget(sass):
    """SURVEY.md 7 (hard part 1): encode has a budget of a few integer instructions per base.  The tile loop of
    the production encode kernel handles 4 x 16 bases per thread; count the SASS between its first 128-bit load and
    its last 32-bit streaming store."""
    funcs, _ = sass
    name = _find(funcs, "encode_kernelILi4ELi512")[0]
    ops = _ops(funcs[name])
    first = next(i for i, o in enumerate(ops) if re.match(r"LDG\.E(\.NA)?\.128", o))
    last = max(i for i, o in enumerate(ops) if o.startswith("STG.E.EF"))
    per_base = (last - first + 1) / 64.0
    assert per_base < 3.5, per_base
    assert sum(o == "POPC" for o in ops) == 0 and not any(o.startswith(("LDS", "STS")) for o in ops)   # pure register arithmetic


def test_decode_is_register_only(sass):
    funcs, _ = sass
    name = _find(funcs, "decode_kernelILi4ELi512")[0]
    ops = _ops(funcs[name])
    assert not any(o.startswith(("LDS", "STS", "LDL", "STL")) for o in ops)
    assert sum(o == "PRMT" for o in ops) >= 24          # PRMT as the 4-entry byte LUT: 6 per 16 bases, 4 words per thread


def test_scan_kernels_have_no_subroutine_calls(sass):
    """The scans run once per read: a 64-bit division by a run-time value (a CALL to the division subroutine) in
    the count functors or the per-item hooks would cost more than the scan itself."""
    funcs, _ = sass
    for name in _find(funcs, "scan_offsets_kernel") + _find(funcs, "scan_block_sums_kernel"):
        assert not any(o.startswith("CALL") for o in _ops(funcs[name])), name
