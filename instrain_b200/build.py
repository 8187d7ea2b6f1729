"""Build libinstrain_b200.so in-tree with nvcc for sm_100a (no other architecture, no fallback).

    python -m instrain_b200.build [--force]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libinstrain_b200.so")
CSRC_SYNTH = os.path.join(HERE, "csrc_synth")
LIB_SYNTH = os.path.join(LIB_DIR, "libisb_synth.so")     # bench/test support: device-side synthetic data generator

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # K2/K3 floating point must not be contracted (bit-exact with CPython doubles)
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _stale(lib, srcs):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    synth_src = sorted(glob.glob(os.path.join(CSRC_SYNTH, "*.cu")))
    for lib, srcs in ((LIB, sources()), (LIB_SYNTH, synth_src)):
        if force or _stale(lib, srcs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", lib] + (["-lz"] if lib == LIB else [])
            subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines):
    """A/B build of the same sources with extra -D defines -> lib/libisbv_<name>.so (select it with ISB_LIB_PATH)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libisbv_%s.so" % name)   # "isbv": never matches libisb_synth.so in a cleanup glob
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + sources() + ["-o", out, "-lz"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
