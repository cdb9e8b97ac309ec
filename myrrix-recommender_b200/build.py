"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU).

    python myrrix-recommender_b200/build.py [--force]

Produces myrrix-recommender_b200/libmyrrix_als.so.  The .so is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmyrrix_als.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))  # (csrc_host/ builds separately)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "myrrix_als.h")
    return any(os.path.getmtime(s) > t for s in _sources() + [hdr, os.path.abspath(__file__)])


INGEST_SRC = os.path.join(HERE, "csrc_host", "ingest.cpp")
INGEST_LIB = os.path.join(HERE, "libmyrrix_ingest.so")
CXX = os.environ.get("CXX", "g++")


def build_ingest(force=False):
    """Host-only library (input canonicalisation, include/myrrix_ingest.h): plain g++."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "myrrix_ingest.h")
    if (not force and os.path.exists(INGEST_LIB) and
            all(os.path.getmtime(f) <= os.path.getmtime(INGEST_LIB) for f in (INGEST_SRC, hdr))):
        return INGEST_LIB
    subprocess.check_call([CXX, "-O3", "-std=c++17", "-pthread", "-fPIC", "-shared", "-Wall",
                           "-o", INGEST_LIB, INGEST_SRC])
    return INGEST_LIB


FOLDIN_SRC = os.path.join(HERE, "csrc_host", "foldin.cpp")
FOLDIN_LIB = os.path.join(HERE, "libmyrrix_foldin.so")


def build_foldin(force=False):
    """Host-only library (fold-in math, include/myrrix_foldin.h): plain g++."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "myrrix_foldin.h")
    if (not force and os.path.exists(FOLDIN_LIB) and
            all(os.path.getmtime(f) <= os.path.getmtime(FOLDIN_LIB) for f in (FOLDIN_SRC, hdr))):
        return FOLDIN_LIB
    subprocess.check_call([CXX, "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", FOLDIN_LIB, FOLDIN_SRC])
    return FOLDIN_LIB


MODEL_IO_SRC = os.path.join(HERE, "csrc_host", "model_io.cpp")
MODEL_IO_LIB = os.path.join(HERE, "libmyrrix_model_io.so")


def build_model_io(force=False):
    """Host-only library (model file codec, include/myrrix_model_io.h): plain g++."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "myrrix_model_io.h")
    if (not force and os.path.exists(MODEL_IO_LIB) and
            all(os.path.getmtime(f) <= os.path.getmtime(MODEL_IO_LIB) for f in (MODEL_IO_SRC, hdr))):
        return MODEL_IO_LIB
    subprocess.check_call([CXX, "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", MODEL_IO_LIB,
                           MODEL_IO_SRC])
    return MODEL_IO_LIB


INIT_SRC = os.path.join(HERE, "csrc_host", "initial_y.cpp")
INIT_LIB = os.path.join(HERE, "libmyrrix_init.so")


def build_init(force=False):
    """Host-only library (cold start: MersenneTwister stream, constructInitialY; include/myrrix_init.h).
    -ffp-contract=off: the reference's float products are rounded before they are widened."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "myrrix_init.h")
    if (not force and os.path.exists(INIT_LIB) and
            all(os.path.getmtime(f) <= os.path.getmtime(INIT_LIB) for f in (INIT_SRC, hdr))):
        return INIT_LIB
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", INIT_LIB,
                           INIT_SRC])
    return INIT_LIB


def build(force=False, verbose=False):
    build_init(force)
    build_ingest(force)
    build_foldin(force)
    build_model_io(force)
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB, os.path.join(CSRC, "als_abi.cu"), "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
