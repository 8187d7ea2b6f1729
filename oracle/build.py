"""Build the oracle's C restatement (test infrastructure only): gcc -> oracle/_build/liboracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
                           "-std=c99", src, "-o", LIB, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
