"""Build librn_b200.so (the C-ABI library, include/rn_b200.h) in-tree with nvcc for sm_100a.

    python retinanet-for-table-detection_b200/build.py [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file, is git-ignored and
travels to the GPU box with the snapshot.  -fmad=false keeps every fp32/fp64 expression in the
reference's operation order (no FMA contraction), which the bit-exactness of K1/K3/K5 relies on.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librn_b200.so")
STAMP = LIB + ".stamp"
# -cudart shared: the CUDA runtime is NOT linked into the library (the process' libcudart.so.12 -- torch's -- is used; the
# rpath covers a plain ctypes load without torch), so the artefact carries none of the runtime's own entry points
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
              "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "rn_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.isfile(LIB) or not os.path.isfile(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def build(force=False, verbose=False):
    """Compile if sources changed.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + _sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%s):\n%s" % (" ".join(cmd), proc.stdout))
    if verbose:
        print(proc.stdout)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


def build_variant(name, defines):
    """A/B builds for the tuning sweeps under profiles/: librn_b200.<name>.so compiled with extra -D flags;
    select it at run time with RN_B200_LIB=<path> (see _lib.py)."""
    out = os.path.join(HERE, "librn_b200.%s.so" % name)
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + _sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%s):\n%s" % (" ".join(cmd), proc.stdout))
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
