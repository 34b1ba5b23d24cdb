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
    "trace/expand.cu",
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


# EZK_EXTRA_NVCC_FLAGS: extra compile-time defines for A/B measurements (e.g. -DEZK_NTT_PRE_TWIDDLES=0); part of the stamp
def _extra_flags():
    return os.environ.get("EZK_EXTRA_NVCC_FLAGS", "").split()


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(str(p).encode())
        h.update(Path(p).read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
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
        cmd = [nvcc, "-ccbin", cxx, *NVCC_FLAGS, *_extra_flags(), "-x", "cu", "-c", str(src), "-o", str(obj)]
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


def build_oracle(force: bool = False) -> Path:
    """Test infrastructure only: the CPU restatement in oracle/ (never loaded by the product path)."""
    srcs = [ORACLE_DIR / f for f in ("capi.cpp", "tracegen.cpp", "stark.hpp", "air.hpp", "ntt.hpp", "f128.hpp", "blake3.hpp")]
    srcs.append(ROOT / "include" / "ezkvm_rescue_constants.h")
    srcs += [CSRC / "host" / "vm.cc", CSRC / "host" / "vm.h", CSRC / "field" / "f128_host.h"]  # tracegen.cpp's input generator
    stamp = ORACLE_DIR / ".stamp"
    digest = _digest(srcs)
    if not force and ORACLE_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return ORACLE_LIB
    base = [_host_cxx(), "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
            "-o", str(ORACLE_LIB), str(ORACLE_DIR / "capi.cpp"), str(ORACLE_DIR / "tracegen.cpp")]
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
