"""Run the reference's `BubbleTransfer.prolong` / `.restrict` (alfi/bubble.py:204-265) VERBATIM.

Test infrastructure (see oracle/__init__.py).  The two methods are sequences of PyOP2 `par_loop`s over the
reference's C kernels, divisions by multiplicity counts, a facet-wise rescaling and two standard transfers.
Here `op2.par_loop` executes the reference's own kernels — compiled from the source strings in alfi/bubble.py by
oracle/build_ref.py — cell by cell with PyOP2's access semantics (READ = gather, INC = scatter-add of a zeroed
local buffer); `Function.dat` is a numpy array; `assemble_rhs` + `pointwiseMult(b, ainv)` is the diagonal facet
"solve" of bubble.py:25-39 (normal component / 0.625, tangential kept); `prolong` / `restrict` of the P1 and
FacetBubble parts are the point-evaluation transfers of oracle/bubble.py.  The object the methods run on carries
the attributes `BubbleTransfer.__init__` would have created (its Firedrake-heavy constructor is not executed; the
two `count` par_loops of bubble.py:192-201 are issued here with the same arguments).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import types

import numpy as np

from . import build_ref, refshim
from .bubble import LiteralBubbleTransfer

READ, INC = "READ", "INC"


class _Vec:
    def __init__(self, data):
        self.data = data

    def zeroEntries(self):
        self.data[...] = 0.0

    def pointwiseMult(self, a, b):
        self.data[...] = a.data * b.data


class _Dat:
    def __init__(self, n):
        self.data = np.zeros((n, 3))

    def __call__(self, access, cmap):
        return (self, access, cmap)

    @property
    @contextlib.contextmanager
    def vec_wo(self):
        yield _Vec(self.data)

    vec_ro = vec_wo


class Fn:
    """A vector Function on the nodes `nodes` of the P1FB space (all of them, its vertices or its faces)."""

    def __init__(self, V, nodes, local):
        self.V, self.nodes, self.local = V, nodes, local        # local: columns of V.cell_nodes this part uses
        self.index = np.full(V.nnodes, -1, dtype=np.int64)
        self.index[nodes] = np.arange(nodes.size)
        self.dat = _Dat(nodes.size)

    def cell_node_map(self):
        return self.index[self.V.cell_nodes[:, self.local]]

    def ufl_domain(self):
        return types.SimpleNamespace(cell_set=self.V.mesh.nc)

    def function_space(self):
        return self


def par_loop(kernel, cell_set, *args):
    lib = build_ref.load()
    fn = getattr(lib, kernel)
    dp = C.POINTER(C.c_double)
    for c in range(cell_set):
        bufs = []
        for dat, access, cmap in args:
            rows = cmap[c]
            bufs.append(np.ascontiguousarray(dat.data[rows]) if access == READ else np.zeros((rows.size, 3)))
        fn(*[b.ctypes.data_as(dp) for b in bufs])
        for (dat, access, cmap), b in zip(args, bufs):
            if access == INC:
                np.add.at(dat.data, cmap[c], b)


class Harness:
    def __init__(self, Vc, Vf, c2f):
        self.lit = LiteralBubbleTransfer(Vc, Vf, c2f)
        self.Vc, self.Vf = Vc, Vf
        me = types.SimpleNamespace(Vc=Vc, Vf=Vf)
        for tag, V in (("c", Vc), ("f", Vf)):
            allnodes = np.arange(V.nnodes)
            setattr(me, "p1" + tag, Fn(V, V.vertex_nodes[:, 0], slice(0, 4)))
            setattr(me, "fb" + tag, Fn(V, V.face_nodes[:, 0], slice(4, 8)))
            setattr(me, "countp1" + tag, Fn(V, V.vertex_nodes[:, 0], slice(0, 4)))
            setattr(me, "countfb" + tag, Fn(V, V.face_nodes[:, 0], slice(4, 8)))
            setattr(me, "countv" + tag, Fn(V, allnodes, slice(0, 8)))
        me.rhs = Fn(Vc, Vc.face_nodes[:, 0], slice(4, 8))
        # facet mass diagonal (any positive diagonal gives the same product rhs * ainv)
        X = Vc.mesh.coords[Vc.mesh.faces]
        area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
        me.ainv = _Vec(np.repeat((1.0 / area)[:, None], 3, axis=1))
        self._area = area

        def assemble_rhs():                       # OneFormAssembler(L, tensor=self.rhs).assemble, bubble.py:36-39
            full = np.zeros((Vc.nnodes, 3))
            full[Vc.face_nodes[:, 0]] = me.fbc.dat.data
            scaled = self.lit.scale_normal(Vc, full)[Vc.face_nodes[:, 0]]
            me.rhs.dat.data[...] = scaled * area[:, None]
        me.assemble_rhs = assemble_rhs
        me.split_kernel, me.split_kernel_adj = "split", "splitadj"
        me.combine_kernel, me.combine_kernel_adj = "combine", "combineadj"
        me.count_kernel = "count"
        # bubble.py:192-201, same arguments
        par_loop(me.count_kernel, me.countp1f.ufl_domain().cell_set, me.countvf.dat(INC, me.countvf.cell_node_map()),
                 me.countfbf.dat(INC, me.countfbf.cell_node_map()), me.countp1f.dat(INC, me.countp1f.cell_node_map()))
        par_loop(me.count_kernel, me.countp1c.ufl_domain().cell_set, me.countvc.dat(INC, me.countvc.cell_node_map()),
                 me.countfbc.dat(INC, me.countfbc.cell_node_map()), me.countp1c.dat(INC, me.countp1c.cell_node_map()))
        self.me = me

    def _std(self, which, adjoint):
        lit, Vc, Vf = self.lit, self.Vc, self.Vf
        cn = Vc.vertex_nodes[:, 0] if which == "p1" else Vc.face_nodes[:, 0]
        fn = Vf.vertex_nodes[:, 0] if which == "p1" else Vf.face_nodes[:, 0]
        if not hasattr(self, "_P" + which):            # matrix of the point-evaluation transfer, column by column
            P = np.zeros((fn.size, cn.size))
            for j in range(cn.size):
                e = np.zeros((Vc.nnodes, 1))
                e[cn[j]] = 1.0
                P[:, j] = lit.point_prolong(which, e)[fn, 0]
            setattr(self, "_P" + which, P)
        P = getattr(self, "_P" + which)
        return P.T if adjoint else P

    def namespace(self):
        def prolong(src, dst):
            which = "p1" if src is self.me.p1c else "fb"
            dst.dat.data[...] = self._std(which, False) @ src.dat.data

        def restrict(src, dst):
            which = "p1" if src is self.me.p1f else "fb"
            dst.dat.data[...] = self._std(which, True) @ src.dat.data
        op2 = types.SimpleNamespace(Kernel=lambda src, name: name, par_loop=par_loop, READ=READ, INC=INC)
        return dict(op2=op2, prolong=prolong, restrict=restrict)

    @contextlib.contextmanager
    def reference_class(self):
        """The reference's BubbleTransfer class, loaded from alfi/bubble.py with the stand-ins above."""
        import importlib.util
        import os
        import sys
        names = self.namespace()
        extra = {"firedrake.assemble": refshim._module("firedrake.assemble", OneFormAssembler=None)}
        with refshim.reference_modules(extra_firedrake=names, extra_modules=extra):
            path = os.path.join(refshim.REFERENCE, "alfi", "bubble.py")
            spec = importlib.util.spec_from_file_location("_alfi_reference_bubble", path)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            yield module.BubbleTransfer
            sys.modules.pop("_alfi_reference_bubble", None)

    def full(self, V, values=None):
        f = Fn(V, np.arange(V.nnodes), slice(0, 8))
        if values is not None:
            f.dat.data[...] = np.asarray(values, dtype=np.float64).reshape(V.nnodes, 3)
        return f
