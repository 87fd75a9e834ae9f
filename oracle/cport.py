"""CPU ORACLE, C/OpenMP kernels — test infrastructure and CPU baseline only.

Builds oracle/c/alfi_oracle.c into oracle/_build/liboracle.so with gcc and swaps the C kernels
into :mod:`oracle.hotpath` levels (`accelerate`), so the F-cycle of the CPU baseline uses every
host core for its memory-bound parts (patch apply, BSR SpMV, P_H apply).  Dense patch inverses
still come from LAPACK (setup is not timed).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "alfi_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        # idle OpenMP workers must sleep, not spin, or they fight numpy's BLAS threads
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_num_threads(n: int):
    """Fix the OpenMP team size explicitly (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def num_threads():
    try:
        omp = C.CDLL("libgomp.so.1")
        return int(omp.omp_get_max_threads())
    except OSError:
        return os.cpu_count() or 1


class CLevel:
    """C-kernel view of one oracle level (inverse mode)."""

    def __init__(self, lv, colours):
        self.lv = lv
        A = lv.A.tobsr(blocksize=(lv.bs, lv.bs))
        A.sort_indices()
        self.rowptr = np.ascontiguousarray(A.indptr, np.int32)
        self.colidx = np.ascontiguousarray(A.indices, np.int32)
        self.vals = np.ascontiguousarray(A.data, np.float64)
        self.nbrows = A.shape[0] // lv.bs
        if lv.offsets is not None:
            self.off = np.ascontiguousarray(lv.offsets, np.int64)
            self.dofs = np.ascontiguousarray(lv.dofs, np.int32)
            self.order = np.ascontiguousarray(lv.order, np.int32)
            self.colours = np.ascontiguousarray(colours, np.int32)
            self.ncolour = int(self.colours.max()) + 1
            n = np.diff(self.off)
            self.inv_off = np.concatenate(([0], np.cumsum(n * n))).astype(np.int64)
            self.inv = np.empty(self.inv_off[-1])
            for p, (kind, f) in enumerate(lv.factors):
                assert kind == "inverse"
                self.inv[self.inv_off[p]:self.inv_off[p + 1]] = np.asarray(f).ravel()

    def spmv(self, x):
        y = np.empty_like(x)
        lib().oracle_bsr_spmv(self.nbrows, self.lv.bs, _p(self.rowptr), _p(self.colidx), _p(self.vals), _p(x), _p(y))
        return y

    def smoother_apply(self, x):
        x = np.ascontiguousarray(x)
        y = np.zeros_like(x)
        lib().oracle_patch_apply(len(self.off) - 1, _p(self.off), _p(self.dofs), self.order.size, _p(self.order),
                                 _p(self.colours), self.ncolour, _p(self.inv_off), _p(self.inv), _p(x), _p(y))
        y[self.lv.bc_dofs] = x[self.lv.bc_dofs]
        return y


def accelerate(levels, colours_by_level):
    """Attach C kernels to oracle levels; hotpath.smooth/vcycle pick them up."""
    for lv, col in zip(levels, colours_by_level):
        lv.ckernels = CLevel(lv, col)
    return levels
