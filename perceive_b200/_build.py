"""Build recipe for libperceive_cuda.so (sm_100a only, in-tree).

`python -m perceive_b200._build` compiles every translation unit under
`perceive_b200/csrc/` with nvcc (cross-compiles without a GPU) in parallel and
links `perceive_b200/libperceive_cuda.so`.  The .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libperceive_cuda.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + [
    "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall",
    "-I", str(ROOT / "include"),
] + os.environ.get("PCV_BUILD_DEFINES", "").split()  # A/B experiments only: extra -D flags (part of the build digest)

# (source, object tag, extra defines)
UNITS = [
    ("pcv_api.cu", "api", []),
    ("pcv_gemm.cu", "gemm", []),
    ("pcv_sqlite.cu", "sqlite", []),
    ("pcv_scan_inst.cu", "scan_f32_dot", ["-DPCV_T=float", "-DPCV_COS=false", "-DPCV_TAG=f32_dot", "-DPCV_GROUPED"]),
    ("pcv_scan_inst.cu", "scan_split_dot", ["-DPCV_T=SplitF32", "-DPCV_COS=false", "-DPCV_TAG=split_dot", "-DPCV_GROUPED"]),
    ("pcv_scan_inst.cu", "scan_f32_cos", ["-DPCV_T=float", "-DPCV_COS=true", "-DPCV_TAG=f32_cos"]),
    ("pcv_scan_inst.cu", "scan_bf16_dot", ["-DPCV_T=uint16_t", "-DPCV_COS=false", "-DPCV_TAG=bf16_dot"]),
    ("pcv_scan_inst.cu", "scan_bf16_cos", ["-DPCV_T=uint16_t", "-DPCV_COS=true", "-DPCV_TAG=bf16_cos"]),
]


def _digest(paths, extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _sources():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + [ROOT / "include" / "perceive_cuda.h"]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile if sources changed since the last build; returns the .so path."""
    OBJ.mkdir(exist_ok=True)
    stamp = OBJ / "stamp"
    # the digest must not depend on where the repo is checked out (the GPU box uses another path)
    flags = " ".join(f for f in COMMON if not f.startswith(str(ROOT)))
    digest = _digest(_sources(), flags + repr(UNITS))

    def current() -> bool:
        return LIB.exists() and stamp.exists() and stamp.read_text() == digest

    if not force and current():
        return LIB
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build libperceive_cuda.so")
    # one builder at a time (torchrun starts one process per GPU): the others wait, then reuse
    lock = open(OBJ / ".lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and current():
            return LIB
        return _build_locked(stamp, digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(stamp: Path, digest: str, verbose: bool) -> Path:

    def compile_one(unit):
        src, tag, defs = unit
        obj = OBJ / f"{tag}.o"
        cmd = [NVCC] + COMMON + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} [{tag}]:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(UNITS))) as ex:
        objs = list(ex.map(compile_one, UNITS))
    tmp = LIB.with_suffix(".so.tmp")
    link = [NVCC] + ARCH + ["-shared", "-o", str(tmp)] + [str(o) for o in objs] + [
        "-Xlinker", "--no-undefined", "-lcudart", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)  # atomic: a process that already mapped the old file keeps it
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
