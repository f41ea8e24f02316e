"""Builds libtaub200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m taufactor_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libtaub200.so")
SOURCES = ["taub_core.cu", "taub_init.cu", "taub_sweep.cu", "taub_fused.cu", "taub_resident.cu", "taub_metrics.cu", "taub_percolate.cu",
           "taub_tiff.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",              # never contract a*b+c: every fp32 op rounds like the reference's
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "taub200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("TAUB_NVCC_EXTRA", "").split()     # extra nvcc flags (-D...) for an A/B build
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
