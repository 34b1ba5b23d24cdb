"""In-tree build of the CUDA library (sm_100a only) and of the CPU oracle used by the tests.

    python -m encrypt_zkvm_b200.build          # builds encrypt_zkvm_b200/libezkvm.so and oracle/liborc.so

nvcc cross-compiles without a GPU; the resulting .so files travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libezkvm.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liborc.so"

SOURCES = [
    "capi.cu",
    "prover.cu",
    "verifier.cu",
    "ntt/ntt.cu",
    "merkle/merkle.cu",
    "air/constraints.cu",
    "compose/compose.cu",
    "fri/fri.cu",
    "host/transcript.cc",
    "host/vm.cc",
    "dist/comm.cc",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


def _host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(str(p).encode())
        h.update(Path(p).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _all_inputs():
    files = [p for p in CSRC.rglob("*") if p.suffix in (".cu", ".cuh", ".cc", ".h")]
    files += list((ROOT / "include").glob("*.h"))
    return files


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    stamp = BUILD / "stamp"
    digest = _digest(_all_inputs())
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    BUILD.mkdir(exist_ok=True)
    nvcc, cxx = _nvcc(), _host_cxx()

    def compile_one(rel: str):
        src = CSRC / rel
        obj = BUILD / (rel.replace("/", "_") + ".o")
        cmd = [nvcc, "-ccbin", cxx, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (rel.replace("/", "_") + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {rel}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-ccbin", cxx, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


# Experimental build variants: the default objects with some translation units recompiled under extra defines, linked
# into libezkvm_<name>.so next to the default library.  A variant is never loaded unless EZKVM_LIB points at it
# (encrypt_zkvm_b200/_lib.py); tools/variant_probe.py compares its proofs and stage times with the default's.
VARIANTS = {
    # in-tile NTT twiddles in precomputed form (fe_mul_pre): both passes / the final pass only (ntt/ntt.cu)
    "pretw": {"defines": ["-DEZK_NTT_PRE_TWIDDLES=1"], "sources": ["ntt/ntt.cu"]},
    "pretwf": {"defines": ["-DEZK_NTT_PRE_TWIDDLES=2"], "sources": ["ntt/ntt.cu"]},
    # inter-pass twiddle tables of the strided passes in precomputed form (ntt/ntt.cu)
    "prepass": {"defines": ["-DEZK_NTT_PRE_PASS_TABLE=1"], "sources": ["ntt/ntt.cu"]},
}


def build_variant(name: str, force: bool = False) -> Path:
    spec = VARIANTS[name]
    build_cuda()
    out = PKG / f"libezkvm_{name}.so"
    vdir = BUILD / name
    stamp = vdir / "stamp"
    digest = _digest(_all_inputs()) + " ".join(spec["defines"])
    if not force and out.exists() and stamp.exists() and stamp.read_text() == digest:
        return out
    vdir.mkdir(parents=True, exist_ok=True)
    nvcc, cxx = _nvcc(), _host_cxx()
    objs = []
    for rel in SOURCES:
        if rel in spec["sources"]:
            obj = vdir / (rel.replace("/", "_") + ".o")
            cmd = [nvcc, "-ccbin", cxx, *NVCC_FLAGS, *spec["defines"], "-x", "cu", "-c", str(CSRC / rel), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            (vdir / (rel.replace("/", "_") + ".log")).write_text(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {rel} ({name}):\n{r.stdout}\n{r.stderr}")
        else:
            obj = BUILD / (rel.replace("/", "_") + ".o")
        objs.append(obj)
    cmd = [nvcc, "-ccbin", cxx, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *map(str, objs), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed ({name}):\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return out


def build_oracle(force: bool = False) -> Path:
    """Test infrastructure only: the CPU restatement in oracle/ (never loaded by the product path)."""
    srcs = [ORACLE_DIR / f for f in ("capi.cpp", "stark.hpp", "air.hpp", "ntt.hpp", "f128.hpp", "blake3.hpp")]
    srcs.append(ROOT / "include" / "ezkvm_rescue_constants.h")
    stamp = ORACLE_DIR / ".stamp"
    digest = _digest(srcs)
    if not force and ORACLE_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return ORACLE_LIB
    base = [_host_cxx(), "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
            "-o", str(ORACLE_LIB), str(ORACLE_DIR / "capi.cpp")]
    r = subprocess.run(base + ["-fopenmp"], capture_output=True, text=True)
    if r.returncode != 0:  # OpenMP runtime missing: build single-threaded
        r = subprocess.run(base, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return ORACLE_LIB


def build_ubench(force: bool = False):
    """Micro-benchmarks of tools/ubench (stand-alone sm_100a binaries next to their sources; not part of the library)."""
    out = []
    for src in sorted((ROOT / "tools" / "ubench").glob("*.cu")):
        exe = src.with_suffix("")
        if force or not exe.exists() or exe.stat().st_mtime < src.stat().st_mtime:
            cmd = [_nvcc(), "-ccbin", _host_cxx(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "-I", str(CSRC), "-o", str(exe), str(src)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        out.append(exe)
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_cuda(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
    for exe in build_ubench(force=force):
        print(exe)
    for name in VARIANTS:
        print(build_variant(name, force=force))
