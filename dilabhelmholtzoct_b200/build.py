"""Builds libtopoloss.so in-tree with nvcc for sm_100a (the only target; no fallbacks)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libtopoloss.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(_HERE, "..", "include", "topoloss.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB_PATH, os.path.join(CSRC, "topoloss_api.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


def build_variant(name: str, defines) -> str:
    """Experiment build: libtopoloss_<name>.so with extra -D flags (select it with TL_LIB_PATH)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(_HERE, f"libtopoloss_{name}.so")
    subprocess.check_call([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out, os.path.join(CSRC, "topoloss_api.cu")], cwd=CSRC)
    return out


def build_stats() -> str:
    """Debug build with merge event counters (-DTL_STATS); used by scripts/stats_probe.py only."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(_HERE, "libtopoloss_stats.so")
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DTL_STATS", "-o", out, os.path.join(CSRC, "topoloss_api.cu")], cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
