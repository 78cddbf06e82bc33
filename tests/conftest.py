"""Shared test plumbing.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200."""
import ctypes
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "radar-signal-simulation-and-target-detection_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _load_oracle():
    spec = importlib.util.spec_from_file_location("rsp_oracle", os.path.join(ROOT, "oracle", "rsp_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rsp_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


oracle = _load_oracle()          # tests are one of the few places allowed to import oracle/


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def emul_lib():
    """Host-compiled kernel phases (csrc/host_emul.cpp); built on demand with g++ (no GPU)."""
    so = os.path.join(PKG_DIR, "lib", "librsp_emul.so")
    src = os.path.join(PKG_DIR, "csrc", "host_emul.cpp")
    deps = [src] + [os.path.join(PKG_DIR, "csrc", f) for f in ("rsp_math.cuh", "rsp_phases.cuh", "rsp_plan.hpp", "rsp_dft_big.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src,
                               "-I", os.path.join(PKG_DIR, "csrc"), "-o", so])
    return ctypes.CDLL(so)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
