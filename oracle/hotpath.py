"""CPU ORACLE — test infrastructure only, never a product path.

Plain numpy/scipy restatement of the velocity-block multigrid hot path that alfi configures
(SURVEY.md §8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.

**PARITY UNPINNED.**  The arithmetic of this path lives in PETSc (PCPATCH, PCMG, KSPFGMRES,
MatMult_SeqBAIJ, LAPACK/UMFPACK patch LU) and Firedrake (PatchPC, prolong/restrict) — none of
which is vendored under /root/reference, no version is pinned (setup.py:1-6), neither can be
imported or built in this image, and the reference ships no tests, golden vectors or fixtures
for the path.  What is restated here follows the reference's *call sites* —
the option dictionaries of alfi/solver.py:305-514 and the transfer code of
alfi/transfer.py:186-275 — plus the published PETSc/Firedrake semantics written down in
SURVEY.md Appendix A.  The restatement is cross-checked by first-principles property tests
(tests/test_oracle_properties.py): dense Σ RᵀA⁻¹R formula, restrict == prolongᵀ, kernel of the
divergence preserved by the Schöberl prolongation, FGMRES against a dense least-squares
solve, F-cycle against explicit recursion.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

__all__ = ["OracleLevel", "patch_matrices", "factor_patches", "smoother_apply", "spmv", "residual",
           "prolong", "restrict", "fgmres", "fcycle", "level_from_host", "backward_error"]


# --------------------------------------------------------------------------- S1: patch setup
def patch_matrices(A_csr, offsets, dofs, corr=None):
    """A_i = A[I_i, I_i] (+ C_i) for every patch.

    PCPATCH with ``save_operators`` + ``precompute_element_tensors`` (alfi/solver.py:320,325)
    sums the element tensors of the patch cells restricted to the kept dofs (Appendix A.2).
    Without interior-facet integrals (stabilisation none/supg/gls) every cell containing two
    kept dofs is a patch cell, so this equals the sub-matrix of the assembled operator.
    With Burman's interior-facet term (stabilisation.py:156-162) PCPATCH integrates only over the
    facets whose both cells are patch cells, and the difference to the sub-matrix is handed over as
    ``corr = (off, rows, cols, vals)``: COO entries in patch-local indices (SURVEY H4;
    tests/test_burman.py checks them against a patch-by-patch assembly).
    """
    out = []
    for i in range(len(offsets) - 1):
        I = dofs[offsets[i]:offsets[i + 1]]
        M = A_csr[I][:, I].toarray() if I.size else np.zeros((0, 0))
        if corr is not None:
            off, rows, cols, vals = corr
            e = slice(off[i], off[i + 1])
            np.add.at(M, (rows[e], cols[e]), vals[e])
        out.append(M)
    return out


def factor_patches(mats, mode="inverse"):
    """``inverse``: LAPACK getrf+getri, as `patch_pc_patch_dense_inverse` (solver.py:602);
    ``lu``: LAPACK getrf, as `patch_sub_pc_type lu` on a dense sub-matrix (solver.py:327,600)."""
    if mode == "inverse":
        return [("inverse", np.linalg.inv(M) if M.size else M) for M in mats]
    if mode == "lu":
        return [("lu", sla.lu_factor(M) if M.size else None) for M in mats]
    raise ValueError(mode)


def _solve(fac, r):
    kind, f = fac
    if r.size == 0:
        return r
    return f @ r if kind == "inverse" else sla.lu_solve(f, r)


# --------------------------------------------------------------------------- S2: smoother apply
def smoother_apply(x, offsets, dofs, order, factors, bc_dofs):
    """PCApply_PATCH, additive, no partition of unity (solver.py:321-322; Appendix A.3):
    y = Σ_{j in iteration order} R_jᵀ A_j⁻¹ R_j x, then y[bc] = x[bc]."""
    y = np.zeros_like(x)
    for j in order:
        I = dofs[offsets[j]:offsets[j + 1]]
        if I.size:
            y[I] += _solve(factors[j], x[I])
    y[bc_dofs] = x[bc_dofs]
    return y


def smoother_apply_multiplicative(A, x, offsets, dofs, order, factors, bc_dofs, symmetric=False):
    """PCApply_PATCH with `pc_patch_local_type multiplicative` (solver.py:322): the patches of the iteration set one
    after the other, each solving for the CURRENT residual on its dofs, y += R_j^T A_j^-1 R_j (x - A y)
    (PETSc: the residual update with the patch operator including its artificial-boundary columns — every row of a
    patch dof only couples to dofs of the patch's cells, so this is the global residual on those rows);
    `symmetrise_sweep` (solver.py:324) adds the same sweep backwards.  Then y[bc] = x[bc]."""
    y = np.zeros_like(x)
    A = A.tocsr()
    sweeps = [list(order)] + ([list(order)[::-1]] if symmetric else [])
    for sw in sweeps:
        for j in sw:
            I = dofs[offsets[j]:offsets[j + 1]]
            if I.size:
                y[I] += _solve(factors[j], x[I] - A[I] @ y)
    y[bc_dofs] = x[bc_dofs]
    return y


def backward_error(mats, offsets, dofs, x, u_by_patch):
    """max_i ||A_i u_i - r_i|| / (||A_i|| ||u_i|| + ||r_i||): the conditioning-free check (H3)."""
    worst = 0.0
    for i, M in enumerate(mats):
        if M.size == 0:
            continue
        r = x[dofs[offsets[i]:offsets[i + 1]]]
        u = u_by_patch[i]
        worst = max(worst, np.linalg.norm(M @ u - r) /
                    (np.linalg.norm(M, 2) * np.linalg.norm(u) + np.linalg.norm(r) + 1e-300))
    return worst


# --------------------------------------------------------------------------- M1: SpMV / residual
def spmv(A, x):
    """MatMult on the level's BAIJ operator (solver.py:512)."""
    return A @ x


def residual(A, b, x):
    return b - A @ x


# --------------------------------------------------------------------------- level container
@dataclass
class OracleLevel:
    A: sp.csr_matrix
    bc_dofs: np.ndarray
    bs: int
    # smoother
    offsets: np.ndarray | None = None
    dofs: np.ndarray | None = None
    order: np.ndarray | None = None
    factors: list | None = None
    # transfer from the next coarser level
    P: sp.csr_matrix | None = None          # dof-level prolongation (P_H ⊗ I_bs)
    D: sp.csr_matrix | None = None          # gamma * div-div form
    cb_dofs: np.ndarray | None = None
    c_offsets: np.ndarray | None = None
    c_dofs: np.ndarray | None = None
    c_factors: list | None = None
    # coarse solve
    coarse_lu: object | None = None
    ckernels: object | None = None          # oracle/cport.py C kernels, if attached

    @property
    def n(self):
        return self.A.shape[0]


def level_from_host(ld, mode="inverse", with_transfer=True, transfer_mode="lu"):
    """Build an OracleLevel from alfi_b200.synth.problem.LevelData (host hand-over data).

    ``mode`` is the smoother's sub-solver: "inverse" = `dense_inverse` (pkp0, solver.py:599-602),
    "lu" = `sub_pc_type lu` (SV, solver.py:655-659).  The transfer solver is always LU
    (transfer.py:100-113)."""
    A = ld.A.to_csr()
    bs = ld.V.bs
    lv = OracleLevel(A=A, bc_dofs=ld.bc_dofs, bs=bs)
    if ld.patches is not None:
        ps = ld.patches
        lv.offsets, lv.dofs, lv.order = ps.offsets, ps.dofs, ps.order
        corr = None
        if getattr(ps, "corrections", None) is not None:
            corr = (ps.corrections.off, ps.corrections.rows, ps.corrections.cols, ps.corr_vals)
        lv.factors = factor_patches(patch_matrices(A, ps.offsets, ps.dofs, corr), mode)
    if ld.P is not None:
        lv.P = ld.P.tocsr() if getattr(ld, "P_dof_level", False) else sp.kron(ld.P, sp.identity(bs), format="csr")
        if with_transfer and ld.cell_patches is not None:
            cp = ld.cell_patches
            lv.D = ld.D.to_csr()
            lv.cb_dofs = ld.cb_dofs
            lv.c_offsets, lv.c_dofs = cp.offsets, cp.dofs
            lv.c_factors = factor_patches(patch_matrices(ld.A0.to_csr(), cp.offsets, cp.dofs), transfer_mode)
    if ld.index == 0:
        lv.coarse_lu = sla.lu_factor(A.toarray())
    return lv


# --------------------------------------------------------------------------- T3/T4: transfers
def _block_solve(lv: OracleLevel, b):
    """PatchPC apply of the transfer solver: disjoint cell patches, bcs = coarse boundaries
    (transfer.py:100-113, 254-257): y = blockdiag(A0)^-1 b on patch dofs, y[cb] = b[cb]."""
    order = np.arange(len(lv.c_offsets) - 1)
    return smoother_apply(b, lv.c_offsets, lv.c_dofs, order, lv.c_factors, lv.cb_dofs)


def prolong(lv: OracleLevel, coarse, robust=True):
    """transfer.py:246-259 then the fine Dirichlet rows zeroed (Appendix A.6):
    rhs = P_H c ; b = gamma D rhs with coarse-boundary rows zeroed ; t = A0^-1 b ; f = rhs - t."""
    rhs = lv.P @ coarse
    if robust and lv.D is not None:
        b = lv.D @ rhs
        b[lv.cb_dofs] = 0.0                       # assemble(bform, bcs=bcs)   transfer.py:249
        t = _block_solve(lv, b)
        fine = rhs - t
    else:
        fine = rhs
    fine[lv.bc_dofs] = 0.0
    return fine


def restrict(lv: OracleLevel, fine, coarse_bc_dofs, robust=True):
    """transfer.py:261-275 then the coarse Dirichlet rows zeroed (Appendix A.6):
    t = f with coarse-boundary rows zeroed ; r = A0^-1 t ; b = gamma D r ; c = P_H^T (f - b)."""
    if robust and lv.D is not None:
        t = fine.copy()
        t[lv.cb_dofs] = 0.0                       # bcs.apply(tildeu)          transfer.py:266
        r = _block_solve(lv, t)
        b = lv.D @ r                              # assemble(bform) without bcs transfer.py:272
        r2 = fine - b
    else:
        r2 = fine
    coarse = lv.P.T @ r2
    coarse[coarse_bc_dofs] = 0.0
    return coarse


# --------------------------------------------------------------------------- K1: level smoother
def fgmres(Aop, Mop, b, x0, m):
    """KSPFGMRES as configured at solver.py:313-317 (Appendix A.4): right preconditioned,
    classical Gram–Schmidt without refinement, exactly m iterations, x = x0 + Z y."""
    n = b.size
    r0 = b - Aop(x0)
    beta = np.linalg.norm(r0)
    V = np.zeros((m + 1, n))
    Z = np.zeros((m, n))
    H = np.zeros((m + 1, m))
    if beta == 0.0:
        return x0.copy()
    V[0] = r0 / beta
    k_done = 0
    for k in range(m):
        Z[k] = Mop(V[k])
        w = Aop(Z[k])
        h = V[:k + 1] @ w                         # classical GS: all dots against the *same* w
        w = w - V[:k + 1].T @ h
        H[:k + 1, k] = h
        H[k + 1, k] = np.linalg.norm(w)
        k_done = k + 1
        if H[k + 1, k] == 0.0:                    # happy breakdown
            break
        V[k + 1] = w / H[k + 1, k]
    g = np.zeros(k_done + 1)
    g[0] = beta
    y = _hessenberg_lsq(H[:k_done + 1, :k_done], g)
    return x0 + Z[:k_done].T @ y


def _hessenberg_lsq(H, g):
    """min ||g - H y|| by Givens rotations (the update PETSc's FGMRES performs)."""
    H = H.copy()
    g = g.copy()
    k = H.shape[1]
    for j in range(k):
        a, b = H[j, j], H[j + 1, j]
        r = np.hypot(a, b)
        c, s = (1.0, 0.0) if r == 0.0 else (a / r, b / r)
        for col in range(j, k):
            t0, t1 = H[j, col], H[j + 1, col]
            H[j, col] = c * t0 + s * t1
            H[j + 1, col] = -s * t0 + c * t1
        g[j], g[j + 1] = c * g[j] + s * g[j + 1], -s * g[j] + c * g[j + 1]
    y = np.zeros(k)
    for j in range(k - 1, -1, -1):
        y[j] = (g[j] - H[j, j + 1:k] @ y[j + 1:]) / H[j, j]
    return y


# --------------------------------------------------------------------------- C1: PCMG full
def _A(lv):
    ck = getattr(lv, "ckernels", None)          # optional C/OpenMP kernels (oracle/cport.py)
    return ck.spmv if ck is not None else (lambda v: lv.A @ v)


def smooth(lv: OracleLevel, b, x, m):
    ck = getattr(lv, "ckernels", None)
    Mop = ck.smoother_apply if ck is not None else (
        lambda v: smoother_apply(v, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs))
    return fgmres(_A(lv), Mop, b, x, m)


def coarse_solve(lv: OracleLevel, b):
    return sla.lu_solve(lv.coarse_lu, b)


def vcycle(levels, l, b, x, m, robust_restrict=True):
    """One V visit on level l (Appendix A.5)."""
    lv = levels[l]
    if l == 0:
        return coarse_solve(lv, b)
    x = smooth(lv, b, x, m)
    r = b - _A(lv)(x)
    bc = restrict(lv, r, levels[l - 1].bc_dofs, robust=robust_restrict)
    xc = vcycle(levels, l - 1, bc, np.zeros_like(bc), m, robust_restrict)
    x = x + prolong(lv, xc)
    return smooth(lv, b, x, m)


def fcycle(levels, b, m, robust_restrict=True):
    """`fieldsplit_0`: richardson(1) + PCMG full with V inner cycles (solver.py:359-379; A.5)."""
    L = len(levels)
    bs = [None] * L
    bs[L - 1] = b
    for l in range(L - 1, 0, -1):
        bs[l - 1] = restrict(levels[l], bs[l], levels[l - 1].bc_dofs, robust=robust_restrict)
    x = np.zeros_like(bs[0])
    for l in range(L - 1):
        x = vcycle(levels, l, bs[l], x, m, robust_restrict)
        x = prolong(levels[l + 1], x)
    return vcycle(levels, L - 1, bs[L - 1], x, m, robust_restrict)
