"""Run the reference's `AutoSchoeberlTransfer.restrict_or_prolong` (alfi/transfer.py:194-275) VERBATIM.

Test infrastructure (see oracle/__init__.py).  The method is the composition of the robust transfer — which
operator acts on what, where boundary rows are zeroed, that restriction applies the same patch solve and not a
transpose, when the patch operators are rebuilt (rows T3, T4, T5 of SURVEY §8a).  Its building blocks are
Firedrake calls; here each of them is a small stand-in over the synthetic problem's own data:

    FunctionSpace / Function / .dat.data / .dat.vec_ro   numpy arrays per level
    TrialFunction, TestFunction, inner, sym, grad, div, cell_avg, dx, action
                                                          a three-node symbolic algebra that recognises the two
                                                          bilinear forms of transfer.py:295-332 and returns
                                                          {"visc": coefficient, "div": coefficient}
    assemble(bilinear, bcs=, tensor=)                    coefficient-weighted sum of the level's unit-coefficient
                                                          parts (alfi_b200.synth.fem.assemble_parts), no bcs inside
    assemble(action(a, rhs), bcs=, tensor=)              matrix-vector product, bc rows zeroed when bcs are given
                                                          (Firedrake zeroes Dirichlet rows of an assembled 1-form)
    LinearSolver(A, solver_parameters=patchparams).ksp.pc PCPATCH semantics (oracle/pcpatch.py): the patch
                                                          constructor NAMED IN THE REFERENCE'S patchparams is
                                                          instantiated from the reference module and called; patch
                                                          dofs = dofs on its points minus the bc nodes; additive
                                                          solve; y[bc] = x[bc]
    prolong / restrict (firedrake.mg)                     the level's standard prolongation matrix P_H and P_H^T
    dmhooks.add_hooks, solver.inserted_options            null context managers

What this pins is the sequence in the reference's source; what stays ours is what each stand-in computes.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np
import scipy.sparse as sp

from . import pcpatch, refshim


# ------------------------------------------------------------------------------------------ symbolic forms
class Expr:
    def __init__(self, op, *args):
        self.op, self.args = op, args

    def __rmul__(self, k):
        return Expr("scale", float(k), self)

    __mul__ = __rmul__


class Form:
    """Sum of recognised bilinear integrals: {"visc": c1, "div": c2}."""

    def __init__(self, terms, V):
        self.terms, self.V = dict(terms), V

    def __add__(self, other):
        out = dict(self.terms)
        for k, v in other.terms.items():
            out[k] = out.get(k, 0.0) + v
        return Form(out, self.V)

    def __rmul__(self, k):
        return Form({key: float(k) * v for key, v in self.terms.items()}, self.V)

    __mul__ = __rmul__


class Integrand:
    def __init__(self, kind, V, coef=1.0):
        self.kind, self.V, self.coef = kind, V, coef

    def __rmul__(self, k):
        return Integrand(self.kind, self.V, self.coef * float(k))

    def __mul__(self, other):
        if isinstance(other, Measure):
            return Form({self.kind: self.coef}, self.V)
        return Integrand(self.kind, self.V, self.coef * float(other))


class Measure:
    def __call__(self, **kw):                       # dx(metadata={"mode": "vanilla"})  transfer.py:322,330
        return self


class Constant:
    """firedrake.Constant: float()-able, multiplies expressions (transfer.py:181, 296-300)."""

    def __init__(self, value):
        self.value = float(value)

    def assign(self, value):
        self.value = float(value)

    def __float__(self):
        return self.value

    def __mul__(self, other):
        return other.__rmul__(self.value)

    __rmul__ = __mul__


def _inner(a, b):
    def strip(e):                                   # (scale, expression without leading scale factors)
        k = 1.0
        while isinstance(e, Expr) and e.op == "scale":
            k *= e.args[0]
            e = e.args[1]
        return k, e
    ka, ea = strip(a)
    kb, eb = strip(b)
    sig = (ea.op, ea.args[0].op if isinstance(ea.args[0], Expr) else None, eb.op)
    V = _space_of(ea)
    if sig == ("sym", "grad", "grad"):              # inner(2*sym(grad(u)), grad(v))
        assert ka * kb == 2.0
        return Integrand("visc", V)
    if sig[0] in ("div", "cell_avg") and eb.op == "div":   # inner(div(u), div(v)), inner(cell_avg(div(u)), div(v))
        return Integrand("div", V, ka * kb)
    raise NotImplementedError("integrand %r is not one of transfer.py:295-332" % (sig,))


def _space_of(e):
    while isinstance(e, Expr):
        if e.op in ("trial", "test"):
            return e.args[0]
        e = e.args[-1]
    raise ValueError("no argument in expression")


class OneForm:
    def __init__(self, a, fn):
        self.a, self.fn = a, fn


# ------------------------------------------------------------------------------------------ data stand-ins
class _Vec:
    def __init__(self, arr):
        self.array = arr                            # view on the Function's data


class _Dat:
    def __init__(self, n, bs):
        self.data = np.zeros((n, bs))

    @property
    def data_ro(self):
        return self.data

    @property
    @contextlib.contextmanager
    def vec_ro(self):
        yield _Vec(self.data.reshape(-1))

    vec_wo = vec_ro


class Space:
    def __init__(self, harness, level):
        self.h, self.level = harness, level
        ld = harness.prob.levels[level]
        self.ld, self.V = ld, ld.V
        self.section = refshim._Section(ld.level.plex, ld.V)
        self.dm = types.SimpleNamespace(getDefaultSection=lambda: self.section)

    def mesh(self):
        return self.h.hier[self.level]

    def ufl_element(self):
        return types.SimpleNamespace(value_shape=lambda: (self.V.bs,), level=self.level)

    def dim(self):
        return self.V.ndofs


class Function:
    def __init__(self, V):
        self.V = V
        self.dat = _Dat(V.V.nnodes, V.V.bs)

    def function_space(self):
        return self.V

    def ufl_domain(self):
        return self.V.mesh()

    def ufl_element(self):
        return self.V.ufl_element()

    @property
    def ufl_shape(self):
        return (self.V.V.bs,)


class Matrix:
    def __init__(self, form, bcs):
        self.bcs, self.version = bcs, 0
        self.fill(form)

    def fill(self, form):
        ld = form.V.ld
        from alfi_b200.synth.fem import BSR
        parts = form.V.h.parts(ld)
        vals = sum(c * parts[k] for k, c in form.terms.items())
        self.csr = BSR(ld.V.nnodes, ld.V.bs, ld.pattern.rowptr, ld.pattern.colidx, vals).to_csr()
        self.terms = dict(form.terms)
        self.version += 1


class PatchPC:
    """firedrake.PatchPC + PCPATCH for the transfer's LinearSolver (transfer.py:100-113)."""

    def __init__(self, solver):
        self.solver, self.version, self.sets = solver, -1, None

    def _setup(self):
        s = self.solver
        params, A = s.parameters, s.A
        assert params["pc_python_type"] == "firedrake.PatchPC" and params["patch_pc_patch_construct_type"] == "python"
        assert params["patch_pc_patch_partition_of_unity"] is False and params["patch_sub_pc_type"] == "lu"
        prob = s._ctx._problem
        fine = prob.u
        V = fine.function_space()
        plex = V.ld.level.plex
        if self.sets is None:
            cls = params["patch_pc_patch_construct_python_type"].rpartition(".")[2]
            maker = getattr(s.harness.transfer_module, cls)()             # the reference's own class
            ctx = types.SimpleNamespace(_x=fine)
            patches, iset = maker(refshim.FakePC(refshim.coord_plex(plex), ctx=ctx))
            self.sets = [p.getIndices() for p in patches]
            self.order = iset.getIndices()
            self.bc = np.asarray(prob.bcs.nodes, dtype=np.int64)
            self.off, self.dofs = pcpatch.patch_dofs(plex, V.V, self.sets, self.bc)
        M = A.csr.tocsr()
        self.inv = []
        for p in range(len(self.sets)):
            I = self.dofs[self.off[p]:self.off[p + 1]]
            self.inv.append(np.linalg.inv(M[I][:, I].toarray()) if I.size else np.empty((0, 0)))
        self.version = A.version
        s.harness.patch_setups += 1

    def apply(self, x, y):
        if self.version != self.solver.A.version:
            self._setup()
        xa = x.array
        out = np.zeros_like(xa)
        for p in self.order:
            I = self.dofs[self.off[p]:self.off[p + 1]]
            out[I] += self.inv[p] @ xa[I]
        bs = self.solver._ctx._problem.u.function_space().V.bs
        bcd = (self.bc[:, None] * bs + np.arange(bs)[None, :]).ravel()
        out[bcd] = xa[bcd]                                                 # PCApply_PATCH: y[bc] = x[bc]
        y.array[:] = out


class Harness:
    """Everything `restrict_or_prolong` needs for one synthetic problem; `transfer(kind)` returns an instance
    of the REFERENCE'S transfer class wired to it."""

    def __init__(self, prob):
        self.prob = prob
        self.hier = refshim.FakeHierarchy([l.level for l in prob.levels])
        self.spaces = {}
        self.assemblies = {"matrix": 0, "vector": 0}
        self.patch_setups = 0
        self.transfer_module = None
        self._parts = {}

    def parts(self, ld):
        if ld.index not in self._parts:
            from alfi_b200.synth.fem import assemble_parts
            self._parts[ld.index] = assemble_parts(ld.V, ld.pattern, None, self.prob.config.discretisation,
                                                   want=("visc", "div"))
        return self._parts[ld.index]

    def space(self, level):
        if level not in self.spaces:
            self.spaces[level] = Space(self, level)
        return self.spaces[level]

    # ---- the names transfer.py pulls out of `from firedrake import *`
    def namespace(self):
        h = self

        def FunctionSpace(mesh, element):
            return h.space(mesh._level)

        def assemble(x, bcs=None, mat_type=None, tensor=None):
            if isinstance(x, Form):
                h.assemblies["matrix"] += 1
                if tensor is None:
                    return Matrix(x, bcs)
                tensor.fill(x)
                return tensor
            h.assemblies["vector"] += 1
            out = tensor if tensor is not None else Function(x.fn.function_space())
            M = Matrix(x.a, None).csr
            v = M @ x.fn.dat.data.reshape(-1)
            if bcs is not None:
                bs = x.fn.function_space().V.bs
                nodes = np.asarray(bcs.nodes, dtype=np.int64)
                v[(nodes[:, None] * bs + np.arange(bs)[None, :]).ravel()] = 0.0
            out.dat.data[:] = v.reshape(out.dat.data.shape)
            return out

        class LinearSolver:
            def __init__(self, A, solver_parameters=None, options_prefix=None):
                self.A, self.parameters, self.harness = A, dict(solver_parameters), h
                self.ksp = types.SimpleNamespace(pc=PatchPC(self), dm=None)
                self._ctx = None

            def inserted_options(self):
                return contextlib.nullcontext()

        def LinearVariationalProblem(a=None, L=None, u=None, bcs=None):
            return types.SimpleNamespace(a=a, L=L, u=u, bcs=bcs)

        def _SNESContext(problem, mat_type=None, pmat_type=None, appctx=None, options_prefix=None):
            return types.SimpleNamespace(_problem=problem)

        def P_of(fn_fine):
            ld = fn_fine.function_space().ld
            bs = ld.V.bs
            return ld.P.tocsr() if ld.P_dof_level else sp.kron(ld.P, sp.identity(bs), format="csr")

        def prolong(source, target):
            target.dat.data[:] = (P_of(target) @ source.dat.data.reshape(-1)).reshape(target.dat.data.shape)

        def restrict(source, target):
            target.dat.data[:] = (P_of(source).T @ source.dat.data.reshape(-1)).reshape(target.dat.data.shape)

        dmhooks = types.SimpleNamespace(add_hooks=lambda dm, solver, appctx=None: contextlib.nullcontext(),
                                        get_appctx=lambda dm: None)
        names = dict(
            FunctionSpace=FunctionSpace, Function=Function, assemble=assemble, LinearSolver=LinearSolver,
            LinearVariationalProblem=LinearVariationalProblem, prolong=prolong, restrict=restrict,
            TrialFunction=lambda V: Expr("trial", V), TestFunction=lambda V: Expr("test", V),
            grad=lambda e: Expr("grad", e), sym=lambda e: Expr("sym", e), div=lambda e: Expr("div", e),
            cell_avg=lambda e: Expr("cell_avg", e), inner=_inner, dx=Measure(), action=OneForm,
            dmhooks=dmhooks, warning=lambda msg: None, RED="%s", Constant=Constant)
        extra = {"firedrake.solving_utils": _module("firedrake.solving_utils", _SNESContext=_SNESContext)}
        return names, extra

    @contextlib.contextmanager
    def transfer(self, kind="SVSchoeberlTransfer", hierarchy="bary"):
        names, extra = self.namespace()
        with refshim.reference_modules(extra_firedrake=names, extra_modules=extra) as (_, tr):
            self.transfer_module = tr
            nu, gamma = Constant(self.prob.nu), Constant(self.prob.gamma)
            obj = getattr(tr, kind)((nu, gamma), self.prob.config.dim, hierarchy)
            yield obj, nu, gamma

    def function(self, level, values=None):
        f = Function(self.space(level))
        if values is not None:
            f.dat.data[:] = np.asarray(values, dtype=np.float64).reshape(f.dat.data.shape)
        return f


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m
