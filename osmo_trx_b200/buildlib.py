"""Build libtrxb200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and the tests."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtrxb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # decision-bearing arithmetic must not be contracted into FMAs; kernels
    # call fmaf() explicitly where fusing is allowed
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-shared",
]


def sources():
    return [os.path.join(CSRC, "capi.cu"), os.path.join(CSRC, "tables.cpp")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(os.path.dirname(HERE), "include", "trxb200.h")
    return os.path.getmtime(inc) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libtrxb200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
