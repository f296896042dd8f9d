"""Build the CUDA engine in-tree with nvcc for sm_100a (B200).

The shared library is written next to this file (``_libclusttraj_b200.so``) so that it travels to the
GPU box with the repository snapshot.  There is no other backend: if the library is missing and cannot
be built, importing the engine raises.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_libclusttraj_b200.so")
SOURCES = ["engine.cu", "pack.cu", "gram.cu", "ozaki.cu", "consumers.cu", "linkage.cu", "xyzio.cu", "pair.cu", "stepwise.cu"]
HEADERS = ["common.cuh", "solve3x3.cuh", os.path.join("..", "..", "include", "clusttraj_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: clusttraj_b200 needs the CUDA toolkit to build its engine")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
