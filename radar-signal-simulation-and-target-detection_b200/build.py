"""Build librsp.so (sm_100a CUDA kernels + C ABI) in-tree with nvcc; no GPU needed to build.

    python radar-signal-simulation-and-target-detection_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "librsp.so")
EMUL = os.path.join(LIB_DIR, "librsp_emul.so")

CUDA_SOURCES = ["rsp_api.cu"]
CXX_SOURCES = ["rsp_cluster.cpp"]
HEADERS = ["rsp_math.cuh", "rsp_phases.cuh", "rsp_kernels.cuh", "rsp_fused.cuh", "rsp_dbf_tc.cuh", "rsp_cfar1d.cuh", "rsp_plan.hpp", "rsp_dft_big.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; librsp.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in CUDA_SOURCES + CXX_SOURCES + HEADERS]
    deps.append(os.path.join(ROOT, "include", "rsp.h"))
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", LIB]
        cmd += [os.path.join(CSRC, f) for f in CUDA_SOURCES + CXX_SOURCES]
        subprocess.check_call(cmd)
    emul_deps = [os.path.join(CSRC, f) for f in ("host_emul.cpp", "rsp_math.cuh", "rsp_phases.cuh", "rsp_plan.hpp", "rsp_dft_big.cuh")]
    if force or _stale(EMUL, emul_deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               os.path.join(CSRC, "host_emul.cpp"), "-I", CSRC, "-o", EMUL])
    return LIB


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B build of the same ABI with extra -D flags: lib/variants/librsp_<name>.so (select with RSP_LIBRARY)."""
    out_dir = os.path.join(LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"librsp_{name}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + [
        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", out]
    cmd += [os.path.join(CSRC, f) for f in CUDA_SOURCES + CXX_SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")], verbose="--verbose" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
