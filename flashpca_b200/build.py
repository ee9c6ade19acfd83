"""Build recipe for the in-tree native artefacts (no JIT cache: the built .so
files travel to the GPU box with the repository snapshot).

  libflashpca_b200.so   CUDA kernels + C ABI (include/flashpca_b200.h), sm_100a only
  flashpca              the command-line front end (host C++ over the C ABI)
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libflashpca_b200.so")
CLI = os.path.join(HERE, "flashpca")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: flashpca_b200 needs the CUDA toolkit to build")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    sources = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))]
    sources.append(os.path.join(ROOT, "include", "flashpca_b200.h"))
    if not force and _newer(LIB, sources):
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, os.path.join(csrc, "fpb_capi.cu"), "-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False) -> str:
    host = os.path.join(HERE, "host")
    sources = [os.path.join(host, f) for f in sorted(os.listdir(host))
               if f.endswith((".cpp", ".hpp"))]
    if not sources:
        return ""
    sources.append(os.path.join(ROOT, "include", "flashpca_b200.h"))
    if not force and _newer(CLI, sources + [LIB]):
        return CLI
    cpps = [s for s in sources if s.endswith(".cpp")]
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-o", CLI, *cpps, "-L", HERE, "-lflashpca_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return CLI


RAPI = os.path.join(HERE, "rapi_check")


def build_rapi_check(force: bool = False) -> str:
    """Test driver for the flashpcaR entry points (host/flashpcar.hpp)."""
    host = os.path.join(HERE, "host")
    cpps = [os.path.join(host, f) for f in sorted(os.listdir(host))
            if f.endswith(".cpp") and f != "flashpca.cpp"]
    main = os.path.join(host, "rapi", "rapi_check.cpp")
    hdrs = [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".hpp")]
    if not force and _newer(RAPI, cpps + hdrs + [main, LIB]):
        return RAPI
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-o", RAPI, main, *cpps, "-L", HERE, "-lflashpca_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return RAPI


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_cli(force)
    build_rapi_check(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print(LIB)
