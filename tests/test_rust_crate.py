"""CPU-only guards for the Rust crate (rust/bitnuc-cuda), which cannot be compiled in this image (no cargo / rustc):
its build script must compile exactly the sources the exercised build compiles, and its `extern "C"` block must
declare exactly the header's functions with the header's argument types, so that the one artefact nobody can build
here cannot drift from the one that is tested."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "bitnuc_cuda.h"
CRATE = ROOT / "rust" / "bitnuc-cuda"

BASE = {"int": "c_int", "size_t": "usize", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "double": "f64",
        "float": "f32", "char": "c_char", "void": "c_void", "bn_ctx": "bn_ctx", "bn_multi": "bn_multi", "bn_error_t": "bn_error_t"}


def c_type_to_rust(decl: str) -> str:
    """'const uint64_t *const *d_words' -> '*const *const u64'; 'uint64_t counts[4]' -> '*mut u64'; 'size_t n' -> 'usize'."""
    decl = decl.strip()
    array = decl.endswith("]")
    if array:
        decl = decl[: decl.index("[")].strip()
    m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)$", decl)   # strip the parameter name
    body = m.group(1).strip() if m and m.group(1).strip() and m.group(2) not in BASE else decl
    parts = [p.strip() for p in body.split("*")]
    base = parts[0].split()
    is_const = "const" in base
    ty = BASE[[t for t in base if t != "const"][0]]
    for qual in parts[1:]:
        ty = f"*{'const' if is_const else 'mut'} {ty}"
        is_const = qual == "const"
    if array:
        ty = f"*{'const' if is_const else 'mut'} {ty}"
    return ty


def header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"^\s*([A-Za-z_][A-Za-z0-9_ ]*?\s*\**)\s*\b(bn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ret = ret.strip()
        rust_ret = None if ret == "void" else c_type_to_rust(ret + " x").replace(" x", "") if "*" not in ret else c_type_to_rust(ret + "x")
        arglist = [] if args.strip() in ("", "void") else [c_type_to_rust(a) for a in args.split(",")]
        protos[name] = (rust_ret, arglist)
    return protos


def rust_prototypes():
    text = (CRATE / "src" / "ffi.rs").read_text()
    block = text[text.index('extern "C" {'):]
    protos = {}
    for name, args, ret in re.findall(r"pub fn (bn_[a-z0-9_]+)\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        arglist = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()]
        protos[name] = (ret.strip() if ret else None, arglist)
    return protos


def test_type_mapping_examples():
    assert c_type_to_rust("const uint64_t *const *d_words") == "*const *const u64"
    assert c_type_to_rust("uint64_t *const *d_counts") == "*const *mut u64"
    assert c_type_to_rust("uint64_t counts[4]") == "*mut u64"
    assert c_type_to_rust("bn_ctx **out") == "*mut *mut bn_ctx"
    assert c_type_to_rust("const bn_ctx *ctx") == "*const bn_ctx"
    assert c_type_to_rust("size_t n") == "usize"
    assert c_type_to_rust("void *stream") == "*mut c_void"


def test_ffi_rs_declares_exactly_the_header():
    c, r = header_prototypes(), rust_prototypes()
    assert len(c) >= 90
    assert set(c) == set(r), sorted(set(c) ^ set(r))
    for name in sorted(c):
        assert c[name] == r[name], (name, c[name], r[name])


def test_bn_error_layout_matches_header():
    text = (CRATE / "src" / "ffi.rs").read_text()
    rust_fields = re.findall(r"pub (\w+): ([^,]+),", text[text.index("pub struct bn_error_t"): text.index("}", text.index("pub struct bn_error_t"))])
    header = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    body = header[header.index("typedef struct bn_error {") + len("typedef struct bn_error {"): header.index("} bn_error_t;")]
    c_fields = []
    for line in body.split(";"):
        line = line.strip()
        if not line:
            continue
        ty, names = line.split(None, 1)
        for nm in names.split(","):
            nm = nm.strip()
            if nm.endswith("]"):
                c_fields.append((nm[: nm.index("[")], f"[{BASE[ty]}; {nm[nm.index('[') + 1:-1]}]"))
            else:
                c_fields.append((nm, BASE[ty]))
    assert [(n, t.replace("i32", "i32")) for n, t in rust_fields] == c_fields
    for const, value in re.findall(r"pub const (BN_[A-Z_]+): c_int = (-?\d+);", text):
        m = re.search(rf"\b{const}\s*=\s*(-?\d+)", header)
        assert m and int(m.group(1)) == int(value), const


def test_build_rs_compiles_the_same_sources_as_build_py():
    from bitnuc_b200 import build
    rs = (CRATE / "build.rs").read_text()
    m = re.search(r"let sources = \[(.*?)\];", rs, flags=re.S)
    sources = re.findall(r'"([^"]+)"', m.group(1))
    assert sources == build.SOURCES
    for s in sources:
        assert (ROOT / "bitnuc_b200" / "csrc" / s).exists()
    # same architecture flags: sm_100a only
    assert '"arch=compute_100a,code=sm_100a"' in rs and "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
    # every csrc translation unit is in the list (a new .cu must be added to both builds)
    assert sorted(p.name for p in (ROOT / "bitnuc_b200" / "csrc").glob("*.cu")) == sorted(sources)


def test_lib_rs_binds_every_host_entry_point_it_mirrors():
    """The safe layer calls the FFI names it documents; a renamed C function must not leave a dangling call."""
    protos = rust_prototypes()
    lib = (CRATE / "src" / "lib.rs").read_text() + (CRATE / "src" / "multi.rs").read_text()
    used = set(re.findall(r"\b(bn_[a-z0-9_]+)\(", lib))
    assert used <= set(protos), sorted(used - set(protos))
    for must in ("bn_encode", "bn_decode", "bn_as_2bit_batch", "bn_from_2bit_batch", "bn_hdist", "bn_hdist_pairs", "bn_base_counts",
                 "bn_multi_create", "bn_multi_encode", "bn_multi_decode", "bn_multi_base_counts", "bn_multi_encode_batch"):
        assert must in used, must
