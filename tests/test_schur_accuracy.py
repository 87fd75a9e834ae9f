"""The numerical fact behind scripts/r2_prep/0005, pinned on the CPU: X_SS = (A^-1)[S,S] of an
augmented-Lagrangian macro-star patch can be formed from the Schur complement A_SS - sum_k A_Sk A_kk^-1 A_kS as
accurately as by cutting it out of the pivoted inverse of the whole patch, PROVIDED A_kk^-1 A_kN comes from a solve
(A_kN carried through the pivoted elimination of A_kk); multiplying by the explicit inverse of A_kk instead loses
several digits.  Reference: LU solve refined with long-double residuals.  (scripts/micro/schur_accuracy.py is the
wider survey; profiles/schur_accuracy_r1.txt its output.)"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def gj_solve(M, R):
    """M^-1 R by Gauss-Jordan with partial pivoting, the row operations applied to R as they happen"""
    M, R = M.copy(), R.copy()
    for j in range(M.shape[0]):
        p = j + int(np.argmax(np.abs(M[j:, j])))
        if p != j:
            M[[j, p]] = M[[p, j]]
            R[[j, p]] = R[[p, j]]
        piv = 1.0 / M[j, j]
        M[j] *= piv
        R[j] *= piv
        f = M[:, j].copy()
        f[j] = 0.0
        M -= np.outer(f, M[j])
        R -= np.outer(f, R[j])
    return R


def test_schur_complement_with_solves_matches_the_full_inverse_cut(problems):
    prob = problems("ldc3d-sv-k3-small")                       # gamma = 1e4, Re = 5000
    ld = prob.levels[-1]
    ps = ld.patches
    A = sp.bsr_matrix((ld.A.vals, ld.A.colidx, ld.A.rowptr), shape=(ld.V.ndofs, ld.V.ndofs)).tocsr()
    p = int(np.flatnonzero(ps.sizes == 609)[0])                # a boundary macro star: 12 blocks of 45, |S| = 69
    I = ps.patch(p)
    blk = ps.blocks[ps.offsets[p]:ps.offsets[p + 1]]
    Ap = A[I][:, I].toarray()
    S = np.flatnonzero(blk < 0)
    B = [np.flatnonzero(blk == k) for k in np.unique(blk[blk >= 0])]
    assert S.size == 69 and len(B) == 12 and all(b.size == 45 for b in B)
    # reference
    E = np.zeros((Ap.shape[0], S.size))
    E[S, np.arange(S.size)] = 1.0
    lu = sla.lu_factor(Ap)
    Al = Ap.astype(np.longdouble)
    Xl = sla.lu_solve(lu, E).astype(np.longdouble)
    for _ in range(3):
        R = E.astype(np.longdouble) - Al @ Xl
        Xl = Xl + sla.lu_solve(lu, np.asarray(R, dtype=np.float64)).astype(np.longdouble)
    ref = np.asarray(Xl[S], dtype=np.float64)

    def err(X):
        return np.linalg.norm(X - ref) / np.linalg.norm(ref)

    cut = err(np.linalg.inv(Ap)[np.ix_(S, S)])
    Sc_solve, Sc_explicit = Ap[np.ix_(S, S)].copy(), Ap[np.ix_(S, S)].copy()
    for b in B:
        Sc_solve -= Ap[np.ix_(S, b)] @ gj_solve(Ap[np.ix_(b, b)], Ap[np.ix_(b, S)])
        Sc_explicit -= Ap[np.ix_(S, b)] @ (np.linalg.inv(Ap[np.ix_(b, b)]) @ Ap[np.ix_(b, S)])
    solve = err(gj_solve(Sc_solve, np.eye(S.size)))
    explicit = err(np.linalg.inv(Sc_explicit))
    kappa = np.linalg.cond(Ap)
    print("cond %.1e: full-inverse cut %.1e, Schur with solves %.1e, Schur with explicit block inverses %.1e" % (kappa, cut, solve, explicit))
    assert kappa > 1e8
    assert solve <= 10 * cut and solve < 1e-7
    assert explicit > 100 * solve
