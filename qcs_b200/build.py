"""Builds the native libraries in-tree (qcs_b200/lib/).

  libqcs_cuda.so  hand-written sm_100a CUDA engine + extern "C" ABI (include/qcs_cuda.h)
  libqcs.so       C89 host layer implementing include/qcs.h on top of that ABI

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU
container; the built .so files travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CUDA_SRC = os.path.join(PKG, "csrc", "cuda")
HOST_SRC = os.path.join(PKG, "csrc", "host")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-fmad=false",  # never contract a*b+c: every op rounds like the reference's (-std=c89)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
]
CUDA_UNITS = ["engine.cu", "dist.cu", "kernels_fused.cu", "kernels_simple.cu",
              "kernels_reduce.cu", "planner.cpp"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _generate() -> None:
    """fused_lists.inc (register lists + the per-gate dispatch switch) from its generator."""
    gen = os.path.join(CUDA_SRC, "gen_fused_lists.py")
    for what, name in (("lists", "fused_lists.inc"), ("ops", "fused_ops.h")):
        inc = os.path.join(CUDA_SRC, name)
        if _newer(inc, [gen]):
            text = subprocess.check_output([sys.executable, gen, what], text=True)
            with open(inc, "w") as f:
                f.write(text)


def _headers() -> list[str]:
    hs = [os.path.join(CUDA_SRC, f) for f in os.listdir(CUDA_SRC) if f.endswith((".h", ".cuh", ".inc"))]
    hs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return hs


def _compile(unit: str, force: bool, verbose: bool) -> str:
    src = os.path.join(CUDA_SRC, unit)
    obj = os.path.join(OBJ_DIR, unit + ".o")
    if force or _newer(obj, [src] + _headers()):
        cmd = [NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(-4, "-Xptxas")
            cmd.insert(-4, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def build_all(force: bool = False, verbose: bool = False) -> dict:
    os.makedirs(OBJ_DIR, exist_ok=True)
    _generate()
    with cf.ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(lambda u: _compile(u, force, verbose), CUDA_UNITS))
    cuda_so = os.path.join(LIB_DIR, "libqcs_cuda.so")
    if force or _newer(cuda_so, objs):
        cmd = [NVCC, "-shared", "-o", cuda_so] + objs + ["-ldl", "-Xlinker", "-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    host_so = os.path.join(LIB_DIR, "libqcs.so")
    host_c = os.path.join(HOST_SRC, "qcs_host.c")
    if force or _newer(host_so, [host_c, cuda_so] + _headers()):
        cmd = ["gcc", "-std=c89", "-pedantic", "-Wall", "-Wextra", "-O2", "-ffp-contract=off",
               "-fPIC", "-shared", "-o", host_so, host_c, "-L" + LIB_DIR, "-lqcs_cuda", "-lm",
               "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
        if r.stderr.strip():
            sys.stderr.write(r.stderr)
    return {"cuda": cuda_so, "host": host_so}


if __name__ == "__main__":
    out = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(out)
