"""Build the CUDA engine in-tree with nvcc for sm_100a (B200).

The shared library is written next to this file (``_libclusttraj_b200.so``) so that it travels to the
GPU box with the repository snapshot.  There is no other backend: if the library is missing and cannot
be built, importing the engine raises.
"""
import fcntl
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CLUSTTRAJ_B200_LIB / CLUSTTRAJ_B200_NVCC_FLAGS: development hooks (tools/ab_variants.sh builds kernel variants side by side)
LIB = os.environ.get("CLUSTTRAJ_B200_LIB") or os.path.join(HERE, "_libclusttraj_b200.so")
SOURCES = ["engine.cu", "pack.cu", "gram.cu", "ozaki.cu", "consumers.cu", "linkage.cu", "xyzio.cu", "pair.cu", "stepwise.cu"]
HEADERS = ["common.cuh", "solve3x3.cuh", os.path.join("..", "..", "include", "clusttraj_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: clusttraj_b200 needs the CUDA toolkit to build its engine")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    if os.environ.get("CLUSTTRAJ_B200_LIB"):
        return False   # an explicitly chosen library (a kernel variant under test) is used as it is
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into one shared library; returns its path.  Safe under several processes
    (torchrun ranks, pytest-xdist workers): one builds under a file lock into its own temporary file and publishes it
    with an atomic rename, the others wait for the lock and find the library fresh."""
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():   # another process built it while we waited
                return LIB
            fd, tmp = tempfile.mkstemp(prefix="_libclusttraj_b200.", suffix=".so.tmp", dir=HERE)
            os.close(fd)
            cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("CLUSTTRAJ_B200_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
                  [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
            os.chmod(tmp, 0o755)
            os.replace(tmp, LIB)
            if verbose:
                print(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
