"""In-tree nvcc build of libgonova_hift.so for sm_100a (B200).

The library is the product: there is no CPU fallback and nothing else computes the decoder.  The
built .so lives in gonova_tts_b200/lib/ (git-ignored, but it travels with the working tree)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libgonova_hift.so"
INCLUDE = PKG.parent / "include"
C_HOST_SRC = PKG.parent / "examples" / "c_host.c"
C_HOST = LIBDIR / "hift_c_host"          # plain-C host over the C ABI (tests/test_gpu_c_host.py)

SOURCES = ["api.cu", "conv_tc.cu", "conv_tc2.cu", "conv_pair.cu", "conv_chain.cu", "aux_kernels.cu", "flow_kernels.cu", "flow_blk.cu", "flow_attn.cu", "flow_enc_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgonova_hift.so cannot be built (set NVCC=...)")


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = (list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + list(INCLUDE.glob("*.h")) +
            [C_HOST_SRC])
    if not C_HOST.exists():
        return True
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str):
        obj = objdir / (src + ".o")
        cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], "-I", str(INCLUDE), "-c",
               str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        (objdir / (src + ".ptxas.log")).write_text(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIBDIR / "libgonova_hift.so.tmp"
    r = subprocess.run([nvcc, "-shared", "-o", str(tmp), *map(str, objs)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    build_c_host()
    return LIB


def build_c_host() -> Path:
    """examples/c_host.c -> lib/hift_c_host: gcc, the C ABI header, libcudart and the library; no C++ and no torch."""
    cuda_home = Path(_nvcc()).resolve().parent.parent
    cc = shutil.which("gcc") or "cc"
    cmd = [cc, "-O2", "-std=c11", "-Wall", "-Wextra", "-I", str(INCLUDE), "-I", str(cuda_home / "include"),
           str(C_HOST_SRC), "-o", str(C_HOST), "-L", str(LIBDIR), "-lgonova_hift", "-L", str(cuda_home / "lib64"),
           "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + str(cuda_home / "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building {C_HOST_SRC.name} failed:\n{r.stdout}\n{r.stderr}")
    return C_HOST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
