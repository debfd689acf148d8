"""In-tree build of the native pieces (sm_100a only).

    python -m spotify_recommender_b200.build

produces spotify_recommender_b200/libsr_engine.so (C ABI of include/sr_engine.h:
hand-written CUDA kernels + host engine) and libsr_recommender.so (the C++
`Recommender` class mirror, Recommender.h:28-82 of the reference, on top of it).
Built files are git-ignored but travel to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
ENGINE_SO = os.path.join(PKG, "libsr_engine.so")
RECOMMENDER_SO = os.path.join(PKG, "libsr_recommender.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources() -> list[str]:
    src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h"))]
    return src + [os.path.join(ROOT, "include", "sr_engine.h")]


def build_engine(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(ENGINE_SO, engine_sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("SR_NVCC_EXTRA", "").split(),  # e.g. -DSR_SCAN_TIMING (development)
               "-o", ENGINE_SO, os.path.join(CSRC, "sr_engine.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return ENGINE_SO


def build_recommender(force: bool = False) -> str:
    src = os.path.join(CSRC, "recommender_host.cpp")
    if not os.path.exists(src):
        return ""
    if force or _stale(RECOMMENDER_SO, engine_sources()):
        build_engine(force)
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "-o", RECOMMENDER_SO, src, "-L", PKG, "-lsr_engine", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return RECOMMENDER_SO


CLI_BIN = os.path.join(PKG, "sr_recommend")


def build_cli(force: bool = False) -> str:
    """The batch CLI (csrc/sr_cli.cpp) on the C++ Recommender class."""
    src = os.path.join(CSRC, "sr_cli.cpp")
    if force or _stale(CLI_BIN, engine_sources()):
        build_recommender(force)
        cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", CLI_BIN, src,
               "-L", PKG, "-lsr_recommender", "-lsr_engine", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return CLI_BIN


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_engine(force, verbose)
    build_recommender(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(ENGINE_SO)
