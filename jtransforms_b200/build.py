"""Builds jtransforms_b200/libjtb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python -m jtransforms_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libjtb200.so")
UNITS = ["jtb_ctx", "tile_f64", "tile_f32", "jtb_capi", "jtb_fast", "jtb_fast2", "jtb_mixed", "jtb_r2r_inv", "jtb_stage", "jtb_tma", "jtb_slab"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", CSRC,
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(HERE, "..", "include", "jtb200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources() if os.path.exists(s))


def _compile(unit: str, verbose: bool, objdir: str = OBJ, defines=()) -> str:
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(objdir, unit + ".o")
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(objdir, unit + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stderr[-4000:]))
    if verbose:
        print(r.stderr)
    return obj


def build_variant(name: str, defines) -> str:
    """Tuning builds (A/B of compile-time options): jtransforms_b200/libjtb200_<name>.so, never loaded by default."""
    objdir = OBJ + "_" + name
    out = os.path.join(HERE, "libjtb200_%s.so" % name)
    os.makedirs(objdir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, False, objdir, defines), UNITS))
    r = subprocess.run([NVCC, "-shared", "-o", out] + objs + ["-lcudart", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u + ".cu"))]
    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        objs = list(ex.map(lambda u: _compile(u, verbose), units))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
