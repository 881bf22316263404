"""C-ABI boundary checks that need no GPU: the library loads, exports exactly the
symbols include/perceive_cuda.h declares, fails loudly without a device, and its
host-side codec/distance helpers match the hand-derived known answers."""
import ctypes as C
import json
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())


def _header_symbols():
    text = (ROOT / "include" / "perceive_cuda.h").read_text()
    return sorted(set(re.findall(r"PCV_API\s+[\w\s\*]+?\b(pcv_\w+)\s*\(", text)))


def test_header_binding_and_library_agree(pcv_lib):
    from perceive_b200 import _ffi
    declared = _header_symbols()
    assert declared == sorted(_ffi.SYMBOLS), "ctypes binding out of sync with include/perceive_cuda.h"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_ffi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if " T " in ln)
    assert exported == declared, "the .so must export exactly the declared C ABI"
    for name in declared:
        assert getattr(pcv_lib, name) is not None
    assert pcv_lib.pcv_abi_version() == 2


def test_library_has_no_torch_or_python_dependency(pcv_lib):
    from perceive_b200 import _ffi
    out = subprocess.run(["readelf", "-d", str(_ffi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any("torch" in n or "python" in n or "c10" in n for n in needed), needed
    assert any(n.startswith("libcudart") for n in needed)
    # NCCL is bound lazily with dlopen so the host process keeps control of which build is loaded
    assert not any(n.startswith("libnccl") for n in needed), needed


def test_sm100a_code_is_embedded(pcv_lib):
    from perceive_b200 import _ffi
    out = subprocess.run(["cuobjdump", "-lelf", str(_ffi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]
    assert not re.search(r"sm_(8|9)\d", out), "only sm_100a code may be embedded"


def test_scan_kernel_uses_bulk_async_copy(pcv_lib):
    """The K1 scan stages rows through shared memory with the TMA engine:
    cp.async.bulk shows up as UBLKCP in SASS (B200_PROFILING.md)."""
    from perceive_b200 import _ffi
    out = subprocess.run(f"cuobjdump -sass {_ffi.LIB_PATH} | grep -c UBLKCP", shell=True, capture_output=True, text=True).stdout
    assert int(out.strip() or 0) > 0


def test_no_gpu_is_a_loud_error_not_a_fallback(pcv_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import perceive_b200 as pb
    with pytest.raises(pb.PcvError) as e:
        pb.Index(384)
    assert e.value.code == 2 and "no CPU fallback" in e.value.message
    n = C.c_int32(-1)
    assert pcv_lib.pcv_device_count(C.byref(n)) == 2 and n.value == 0


def test_argument_validation_without_device(pcv_lib):
    h = C.c_void_p()
    assert pcv_lib.pcv_index_create(0, 0, 0, 0, 0, C.byref(h)) == 1  # dim 0
    assert b"dim" in pcv_lib.pcv_last_error()
    assert pcv_lib.pcv_index_create(0, 384, 7, 0, 0, C.byref(h)) == 1  # bad dtype
    assert pcv_lib.pcv_index_create(0, 384, 0, 0, 0xF0, C.byref(h)) == 1  # unknown flags
    assert pcv_lib.pcv_index_create(0, 384, 0, 0, 0, None) == 1
    assert pcv_lib.pcv_search(None, None, 1, 10, None, 0, None, None, None, None) == 1
    assert pcv_lib.pcv_index_destroy(None) == 0
    assert pcv_lib.pcv_index_set_hidden(None, None, 0) == 1
    row = C.c_uint64(7)
    assert pcv_lib.pcv_index_find_id(None, 5, C.byref(row)) == 1 and row.value == 7


def test_codec_known_answers_through_the_abi(pcv_lib):
    import perceive_b200 as pb
    for c in GOLDEN["codec"]:
        blob = bytes.fromhex(c["hex"])
        assert pb.serialize_embedding(c["floats"]) == blob
        assert pb.deserialize_embedding(blob).tolist() == c["floats"]
    for n in GOLDEN["codec_bad_lengths"]:
        with pytest.raises(pb.PcvError) as e:
            pb.deserialize_embedding(b"\0" * n)
        assert e.value.code == 1
    v = np.random.default_rng(0).standard_normal(768).astype(np.float32)
    assert pb.serialize_embedding(v) == v.astype("<f4").tobytes()
    assert np.array_equal(pb.deserialize_embedding(pb.serialize_embedding(v)), v)


def test_distance_known_answers_through_the_abi(pcv_lib, orc):
    for c in GOLDEN["distance"]:
        assert pcv_lib.pcv_distance_from_dot(c["dot"], c["len"]) == c["want"]
    rng = np.random.default_rng(1)
    for d in rng.standard_normal(200).astype(np.float32) * 3:
        assert pcv_lib.pcv_distance_from_dot(float(d), 384) == orc.distance_from_dot(float(d), 384)


def test_host_generator_matches_oracle(pcv_lib, orc):
    from perceive_b200 import _ffi
    for dist, dim, first in ((0, 384, 0), (1, 768, 12345), (0, 100, 7)):
        out = np.empty((16, dim), dtype=np.float32)
        _ffi.check(pcv_lib.pcv_synthetic_rows_host(9, dist, first, 16, dim, out.ctypes.data))
        assert np.array_equal(out, orc.synth_rows(9, dist, first, 16, dim))


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (checked textually)."""
    for p in list((ROOT / "perceive_b200").rglob("*.py")) + list((ROOT / "perceive_b200" / "csrc").glob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".hpp", ".h"}:
            text = p.read_text()
            assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, p


def test_header_is_valid_c_and_host_entry_points_work_from_c(pcv_lib, tmp_path):
    """include/perceive_cuda.h compiles as plain C99 and the device-free entry points behave as
    documented when called from C (the caller a Rust/C integrator would be)."""
    from perceive_b200 import _ffi
    exe = tmp_path / "abi_check"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
           str(Path(__file__).parent / "abi_header_check.c"), "-o", str(exe),
           "-L", str(_ffi.LIB_PATH.parent), "-lperceive_cuda", "-lm", f"-Wl,-rpath,{_ffi.LIB_PATH.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "abi ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_rust_binding_names_every_declared_function():
    """rust/perceive-cuda/src/lib.rs (unbuilt here: no Rust toolchain) must stay in step with the header."""
    rust = (ROOT / "rust" / "perceive-cuda" / "src" / "lib.rs").read_text()
    bound = set(re.findall(r"pub fn (pcv_\w+)\s*\(", rust))
    assert bound == set(_header_symbols()), sorted(set(_header_symbols()) ^ bound)


def test_flag_and_limit_constants_agree_across_the_bindings():
    """Every `#define PCV_FLAG_* / PCV_MAX_*` of the header has the same value in the ctypes binding and in the
    Rust crate (a flag one binding does not know is a feature its callers cannot reach)."""
    from perceive_b200 import _ffi
    text = (ROOT / "include" / "perceive_cuda.h").read_text()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(PCV_(?:FLAG|MAX)_\w+)\s+(\d+)u?\b", text)}
    assert {"PCV_FLAG_PRENORMALISE", "PCV_FLAG_NO_TIMING", "PCV_MAX_K"} <= set(defines)
    rust = (ROOT / "rust" / "perceive-cuda" / "src" / "lib.rs").read_text()
    rust_consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (PCV_\w+): u32 = (\d+);", rust)}
    for name, value in defines.items():
        if name.startswith(("PCV_FLAG_", "PCV_MAX_")):
            assert getattr(_ffi, name) == value, name
            assert rust_consts.get(name) == value, (name, rust_consts.get(name))
    flags = [v for n, v in defines.items() if n.startswith("PCV_FLAG_")]
    assert len(set(flags)) == len(flags) and all(v & (v - 1) == 0 for v in flags), "flags are distinct single bits"


def _header_prototypes():
    """name -> number of parameters, parsed from include/perceive_cuda.h."""
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "perceive_cuda.h").read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"PCV_API\s+[\w\s\*]+?\b(pcv_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return protos


def test_bindings_agree_with_the_header_on_arity(pcv_lib):
    """Every binding of the C ABI — ctypes (executed), Rust extern block (unbuilt), the copy printed in
    INTEGRATION.md — declares as many parameters per function as the header does."""
    protos = _header_prototypes()
    assert sorted(protos) == _header_symbols()
    for name, n in protos.items():
        fn = getattr(pcv_lib, name)
        assert fn.argtypes is not None and len(fn.argtypes) == n, (name, n, fn.argtypes)

    def rust_arity(text):
        out = {}
        for m in re.finditer(r"pub fn (pcv_\w+)\s*\(([^;]*?)\)\s*(?:->\s*[\w\*\s:]+)?;", text, flags=re.S):
            params = re.sub(r"//[^\n]*", "", m.group(2)).strip()
            out[m.group(1)] = 0 if not params else params.rstrip(",").count(",") + 1
        return out

    for path in (ROOT / "rust" / "perceive-cuda" / "src" / "lib.rs", ROOT / "INTEGRATION.md"):
        got = rust_arity(path.read_text())
        for name, n in protos.items():
            assert got.get(name) == n, (path.name, name, n, got.get(name))


def _c_params(params: str):
    """Parameter list of a C prototype -> list of normalised C types (names dropped)."""
    out = []
    if params.strip() in ("", "void"):
        return out
    for prm in params.split(","):
        prm = re.sub(r"/\*.*?\*/", "", prm).strip()
        arr = re.search(r"\[\d*\]$", prm)  # `uint8_t out_id[128]` decays to a pointer
        prm = re.sub(r"\[\d*\]$", "", prm).strip()
        m = re.match(r"^(.*?)(\w+)$", prm)
        ctype = m.group(1).strip() if m and m.group(1).strip() else prm  # drop the parameter name
        ctype = re.sub(r"\s*\*\s*", "*", ctype).strip()
        out.append(ctype + ("*" if arr else ""))
    return out


def _c_to_rust(ctype: str) -> str:
    scalars = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "float": "f32", "size_t": "usize",
               "uint8_t": "u8", "char": "c_char", "void": "c_void", "pcv_dtype": "i32", "pcv_metric": "i32", "pcv_dist": "i32",
               "pcv_index": "pcv_index", "pcv_rowset": "pcv_rowset", "pcv_stats": "pcv_stats"}
    m = re.match(r"^(const\s+)?(\w+)(\**)$", ctype)
    assert m, ctype
    const, base, stars = bool(m.group(1)), scalars[m.group(2)], len(m.group(3))
    if stars == 0:
        return base
    t = ("*const " if const else "*mut ") + base  # the innermost pointer carries the C const
    for _ in range(stars - 1):
        t = "*mut " + t
    return t


def test_rust_extern_block_matches_the_header_types():
    """Argument TYPES, not just names and arity: every parameter of every function in the Rust extern block
    (rust/perceive-cuda/src/lib.rs, and the copy printed in INTEGRATION.md) is the Rust spelling of the
    header's C type, and so is the return type."""
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "perceive_cuda.h").read_text(), flags=re.S)
    want = {}
    for m in re.finditer(r"PCV_API\s+([\w\s\*]+?)\b(pcv_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret = re.sub(r"\s*\*\s*", "*", m.group(1).strip())
        want[m.group(2)] = ([_c_to_rust(t) for t in _c_params(m.group(3))], _c_to_rust(ret))
    assert sorted(want) == _header_symbols()
    for path in (ROOT / "rust" / "perceive-cuda" / "src" / "lib.rs", ROOT / "INTEGRATION.md"):
        src = re.sub(r"//[^\n]*", "", path.read_text())
        got = {}
        for m in re.finditer(r"pub fn (pcv_\w+)\s*\(([^;{]*?)\)\s*(?:->\s*([\w\*\s:]+?))?;", src, flags=re.S):
            params = [p.split(":", 1)[1].strip() for p in m.group(2).split(",") if p.strip()]
            params = [re.sub(r"\s+", " ", p).replace("std::os::raw::", "") for p in params]
            got[m.group(1)] = (params, (m.group(3) or "()").strip())
        for name, (params, ret) in want.items():
            assert name in got, (path.name, name)
            assert got[name][0] == params, (path.name, name, got[name][0], params)
            assert got[name][1] == ret, (path.name, name, got[name][1], ret)


def test_rust_sources_are_complete_files():
    """rust/perceive-core/search.rs and rust/perceive-cli/cmd/bench.rs are whole files, not sketches: no elided
    bodies, balanced delimiters, and every `pub` item of the reference's search.rs is defined."""
    for rel in ("perceive-core/search.rs", "perceive-cli/cmd/bench.rs", "perceive-cuda/src/lib.rs"):
        src = (ROOT / "rust" / rel).read_text()
        code = re.sub(r"//[^\n]*", "", src)
        code = re.sub(r'"(?:\\[\s\S]|[^"\\])*"', '""', code)  # string literals (a backslash may escape a newline)
        assert "/*" not in code and "todo!" not in code and "unimplemented!" not in code, rel
        for o, c in ("{}", "()", "[]"):
            assert code.count(o) == code.count(c), (rel, o, code.count(o), code.count(c))
    search = (ROOT / "rust" / "perceive-core" / "search.rs").read_text()
    for item in ("pub struct SearchItem", "pub struct Searcher", "pub hidden: HashSet<i64>", "pub fn build(", "pub fn rebuild_source(",
                 "fn load_rows(", "pub fn search_vector(", "pub fn search_vectors(", "pub fn search(", "pub fn search_vector_and_retrieve(",
                 "pub fn search_and_retrieve(", "pub fn encode_query(", "pub struct NdArrayDistance", "pub fn deserialize_embedding(",
                 "pub fn serialize_embedding("):
        assert item in search, item
    bench = (ROOT / "rust" / "perceive-cli" / "cmd" / "bench.rs").read_text()
    for item in ("pub struct BenchArgs", "config: String", "gpus: usize", "steps: usize", "warmup: usize", "seed: u64",
                 "pub fn handle_bench_command("):
        assert item in bench, item
