"""Builds oracle/_build/libjt_ref.so (C restatement of the reference's CPU algorithm; test/bench infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libjt_ref.so")


def build():
    src = os.path.join(HERE, "jt_ref.c")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=native", "-std=gnu11", "-shared", "-fPIC", "-pthread", src, "-o", LIB, "-lm"],
                   check=True)
    return LIB


if __name__ == "__main__":
    print(build())
