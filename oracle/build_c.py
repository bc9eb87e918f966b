"""Compile oracle/c/orx_oracle.c -> oracle/_build/liborx_oracle.so (gcc -O3 -fopenmp).
Test infrastructure / CPU baseline; `__graft_entry__.build()` calls this."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "orx_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liborx_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"],
                   check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
