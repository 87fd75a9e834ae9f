"""Compile the reference's OWN C kernels into oracle/_ref/ (test infrastructure, see oracle/__init__.py).

alfi/bubble.py holds the five per-cell kernels of its BubbleTransfer — `split`, `splitadj`, `combine`,
`combineadj`, `count` (bubble.py:57-185) — as C source strings handed to PyOP2.  They are self-contained C, so
they can be compiled as they are: this recipe reads the string literals out of the file *where it lies under
/root/reference* (ast, no import of alfi or Firedrake), writes them to oracle/_ref/bubble_kernels.c and builds
oracle/_ref/libalfi_bubble_ref.so with gcc.  oracle/_ref/ is git-ignored (no reference source enters the
repository) but travels to the GPU box with the snapshot like our own built libraries.

    python -m oracle.build_ref            # build if the reference tree is present
"""
from __future__ import annotations

import ast
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libalfi_bubble_ref.so")
REFERENCE = os.environ.get("ALFI_REFERENCE", "/root/reference")
SOURCE = os.path.join(REFERENCE, "alfi", "bubble.py")
KERNELS = ("split", "splitadj", "combine", "combineadj", "count")


def extract_kernels(path=SOURCE):
    """{kernel name: C source} for every `op2.Kernel("<C source>", "<name>")` call in the file."""
    tree = ast.parse(open(path).read(), path)
    found = {}
    for node in ast.walk(tree):
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "Kernel"
                and len(node.args) >= 2 and all(isinstance(a, ast.Constant) and isinstance(a.value, str) for a in node.args[:2])):
            found[node.args[1].value] = node.args[0].value
    return found


def build(force=False):
    """Build oracle/_ref/libalfi_bubble_ref.so; returns its path, or None when the reference tree is absent
    and no earlier build exists."""
    if not os.path.exists(SOURCE):
        return LIB if os.path.exists(LIB) else None
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SOURCE):
        return LIB
    kernels = extract_kernels()
    missing = [k for k in KERNELS if k not in kernels]
    if missing:
        raise RuntimeError("kernels %s not found in %s" % (missing, SOURCE))
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(OUT, "bubble_kernels.c")
    with open(csrc, "w") as fh:
        fh.write("/* GENERATED from %s by oracle/build_ref.py - reference code, never committed */\n" % SOURCE)
        for name in KERNELS:
            fh.write("\n" + kernels[name].strip() + "\n")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", csrc, "-o", LIB])
    return LIB


def load():
    """ctypes handle with the kernels' signatures (double arrays in the order of the C prototypes), or None."""
    path = build()
    if path is None:
        return None
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    for name in KERNELS:
        getattr(lib, name).argtypes = [dp, dp, dp]
        getattr(lib, name).restype = None
    return lib


if __name__ == "__main__":
    print(build(force=True))
