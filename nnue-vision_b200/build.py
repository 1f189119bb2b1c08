"""Build recipe for libnnue_b200.so: every CUDA source compiled for sm_100a only, in-tree.

    python nnue-vision_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libnnue_b200.so"
CLI = PKG / "nnue_inference_b200"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(CSRC.glob("*.cu"))


def stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    if not CLI.exists():
        return True
    deps = list(CSRC.glob("*")) + list((PKG / "cli").glob("*")) + [ROOT / "include" / "nnue_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC, "-O3", "-std=c++17", "-t", "8", *ARCH_FLAGS, "-lineinfo", "-Xcompiler", "-fPIC,-Wall", "-shared",
           "-I", str(ROOT / "include"), "-I", str(CSRC), "-o", str(LIB), *map(str, sources())]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    # the command-line twin of the reference's nnue_inference (engine/nnue_inference.cpp), host C++ over the C ABI
    cli = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-I", str(ROOT / "include"),
           str(PKG / "cli" / "nnue_inference.cpp"), "-o", str(CLI), "-L", str(PKG), "-lnnue_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cli, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
