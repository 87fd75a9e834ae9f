"""Evaluate the reference's residual forms NUMERICALLY (TEST INFRASTRUCTURE, see oracle/__init__.py).

`ScottVogeliusSolver.residual` / `ConstantPressureSolver.residual` (alfi/solver.py:613-623, 562-572) are UFL
expressions.  This module is a small interpreter for exactly the operators they use — split, TestFunctions,
grad, div, sym, dot, inner, cell_avg, `*dx`, +, -, scalar * — acting on arrays of values at quadrature points of
the synthetic mesh and element (alfi_b200.synth.fem).  Running the reference's method with these names in place of
Firedrake's yields the discrete residual vector the reference's form defines; the tests compare it with what
alfi_b200.synth.fem assembles (the operator handed to the hot path): the viscous and grad-div
parts directly, the Newton-linearised advection through the exact identity
N(u+d) - N(u) - N(d) = ((grad d) u + (grad u) d, v).

Shapes: a field is an array (ncells, nquad, nbasis, *tensor) with nbasis = 1 unless it involves a test function.
"""
from __future__ import annotations

import numpy as np

from alfi_b200.synth.fem import simplex_quadrature


class F:
    """A tensor field at the quadrature points; `space` is None or the test space ('v' / 'q') it is linear in."""

    def __init__(self, val, space=None, gradval=None):
        self.val, self.space, self.gradval = val, space, gradval

    def _bin(self, other, op):
        if isinstance(other, F):
            assert not (self.space and other.space), "product of two test functions"
            a, b = self.val, other.val
            # scalar (no tensor dims) times tensor broadcasts over the tensor dims
            while a.ndim < b.ndim:
                a = a[..., None]
            while b.ndim < a.ndim:
                b = b[..., None]
            return F(op(a, b), self.space or other.space)
        return F(op(self.val, float(other)), self.space)

    def __mul__(self, o):
        if isinstance(o, Measure):
            return o.integrate(self)
        return self._bin(o, np.multiply)

    def __rmul__(self, o):
        return self._bin(o, np.multiply)

    def __add__(self, o):
        return self._bin(o, np.add)

    def __sub__(self, o):
        return self._bin(o, np.subtract)

    def __neg__(self):
        return F(-self.val, self.space)


class Form:
    """Integrated terms: {space: (ncells, nbasis) array}."""

    def __init__(self, parts):
        self.parts = parts

    def _comb(self, other, sign):
        out = {k: v.copy() for k, v in self.parts.items()}
        for k, v in other.parts.items():
            out[k] = out.get(k, 0.0) + sign * v
        return Form(out)

    def __add__(self, o):
        return self._comb(o, 1.0)

    def __sub__(self, o):
        return self._comb(o, -1.0)

    def __neg__(self):
        return Form({k: -v for k, v in self.parts.items()})

    def __rmul__(self, k):
        return Form({key: float(k) * v for key, v in self.parts.items()})

    __mul__ = __rmul__


class Measure:
    def __init__(self, weights):
        self.w = weights                         # (ncells, nquad): quadrature weight * |det J|

    def __call__(self, **kw):                    # dx(metadata={"mode": "vanilla"})
        return self

    def integrate(self, f):
        assert f.val.ndim == 3, "only scalar integrands can be integrated"
        return Form({f.space: np.einsum("cq,cqb->cb", self.w, f.val)})

    __rmul__ = integrate


class Evaluator:
    """Fields of one (velocity space V, pressure values) pair on the synthetic mesh."""

    def __init__(self, V, degree=None):
        self.V = V
        m, el = V.mesh, V.element
        d = m.dim
        x, w = simplex_quadrature(d, degree or 3 * el.degree + 1)
        X = m.coords[m.cells]                                        # (nc, d+1, d)
        J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))       # dx/dxi
        Jinv = np.linalg.inv(J)
        self.detw = np.abs(np.linalg.det(J))[:, None] * w[None, :]
        self.phi = el.tabulate(x)                                    # (q, n)
        dphi = el.tabulate_grad(x)                                   # (q, n, d) reference
        self.gphi = np.einsum("qna,cab->cqnb", dphi, Jinv)           # physical gradients (nc, q, n, d)
        self.d, self.n = d, el.nnodes
        self.dx = Measure(self.detw)

    def trial(self, U):
        """u_h and grad u_h at the quadrature points from nodal values U (nnodes, d)."""
        Uc = U[self.V.cell_nodes]                                    # (nc, n, d)
        val = np.einsum("qn,cni->cqi", self.phi, Uc)[:, :, None, :]
        g = np.einsum("cqnb,cni->cqib", self.gphi, Uc)[:, :, None, :, :]     # d u_i / d x_b
        return F(val, None, g)

    def test(self):
        """All velocity test functions v = phi_a e_r at once: basis index (a, r) -> a*d + r."""
        nc, nq, n, d = self.gphi.shape
        val = np.zeros((nc, nq, n, d, d))
        g = np.zeros((nc, nq, n, d, d, d))
        for r in range(d):
            val[:, :, :, r, r] = self.phi[None, :, :]
            g[:, :, :, r, r, :] = self.gphi
        return F(val.reshape(nc, nq, n * d, d), "v", g.reshape(nc, nq, n * d, d, d))

    def scatter_v(self, form):
        """(ncells, n*d) element vectors of the velocity test space -> global dof vector."""
        V, d = self.V, self.d
        out = np.zeros(V.ndofs)
        idx = (V.cell_nodes[:, :, None] * d + np.arange(d)[None, None, :]).reshape(V.mesh.nc, -1)
        np.add.at(out, idx.ravel(), form.parts["v"].ravel())
        return out

    # ---- the UFL names
    def namespace(self):
        def grad(f):
            assert f.gradval is not None, "second derivatives are not needed by the reference's forms"
            return F(f.gradval, f.space)

        def div(f):
            return F(np.trace(f.gradval, axis1=-2, axis2=-1), f.space)

        def sym(f):
            return F(0.5 * (f.val + np.swapaxes(f.val, -1, -2)), f.space)

        def dot(A, b):                                               # (grad u) u : A_ij b_j
            assert not (A.space and b.space)
            av, bv = _bcast(A.val, b.val, 2, 1)
            return F(np.einsum("cqbij,cqbj->cqbi", av, bv), A.space or b.space)

        def inner(a, b):
            assert not (a.space and b.space)
            nt = a.val.ndim - 3
            av, bv = _bcast(a.val, b.val, nt, nt)
            return F((av * bv).reshape(av.shape[:3] + (-1,)).sum(axis=-1), a.space or b.space)

        def cell_avg(f):
            assert f.val.ndim == 3
            avg = np.einsum("cq,cqb->cb", self.detw, f.val) / self.detw.sum(axis=1)[:, None]
            return F(np.broadcast_to(avg[:, None, :], f.val.shape).copy(), f.space)
        return dict(grad=grad, div=div, sym=sym, dot=dot, inner=inner, cell_avg=cell_avg, dx=self.dx)


def _bcast(a, b, ta, tb):
    """Broadcast the basis axis (axis 2) of two fields."""
    nb = max(a.shape[2], b.shape[2])
    if a.shape[2] != nb:
        a = np.broadcast_to(a, a.shape[:2] + (nb,) + a.shape[3:])
    if b.shape[2] != nb:
        b = np.broadcast_to(b, b.shape[:2] + (nb,) + b.shape[3:])
    return a, b


def reference_velocity_residual(solver_name, V, U, nu, gamma, advect):
    """F_u of the reference's residual form for velocity nodal values U (nnodes, d) and p = 0, as a dof vector."""
    import types

    from . import refshim
    ev = Evaluator(V)
    names = ev.namespace()
    u, v = ev.trial(U), ev.test()
    zero_p = F(np.zeros(ev.detw.shape + (1,)), None)
    qdummy = F(np.zeros(ev.detw.shape + (1,)), "q")
    names.update(split=lambda z: (u, zero_p), TestFunctions=lambda Z: (v, qdummy))
    with refshim.reference_modules(with_solver=True, extra_firedrake=names) as (_, _, sol):
        me = types.SimpleNamespace(z=None, Z=None, nu=nu, gamma=gamma, advect=advect)
        form = getattr(sol, solver_name).residual(me)
    return ev.scatter_v(form)


def reference_residual(solver_name, V, kq, U, P, nu, gamma, advect):
    """(F_u, F_p) of the reference's residual for velocity nodal values U (nnodes, d) and pressure dofs P of the
    discontinuous P_kq space, numbered cell by cell like alfi_b200.synth.fem.assemble_divergence."""
    import types

    from alfi_b200.synth.fem import LagrangeElement

    from . import refshim
    ev = Evaluator(V)
    d = ev.d
    x, _ = simplex_quadrature(d, 3 * V.element.degree + 1)
    psi = np.ones((x.shape[0], 1)) if kq == 0 else LagrangeElement(d, kq).tabulate(x)        # (q, nq)
    nc, nqb = V.mesh.nc, psi.shape[1]
    p = F(np.einsum("qj,cj->cq", psi, np.asarray(P).reshape(nc, nqb))[:, :, None], None)
    q = F(np.broadcast_to(psi[None, :, :], (nc,) + psi.shape).copy(), "q")
    names = ev.namespace()
    u, v = ev.trial(U), ev.test()
    names.update(split=lambda z: (u, p), TestFunctions=lambda Z: (v, q))
    with refshim.reference_modules(with_solver=True, extra_firedrake=names) as (_, _, sol):
        me = types.SimpleNamespace(z=None, Z=None, nu=nu, gamma=gamma, advect=advect)
        form = getattr(sol, solver_name).residual(me)
    return ev.scatter_v(form), form.parts["q"].reshape(-1)
