"""Accuracy of X_SS = (A^-1)[S,S] from the Schur complement, in numpy (no GPU):  python scripts/micro/schur_accuracy.py

Question: the condensed patch inverses need X_SS per patch and Newton step.  Round 1 cuts it out
of the pivoted inverse of the WHOLE patch (2 n^3 flops, n = 1275) because a first numpy attempt at the Schur complement
A_SS - sum_k A_Sk D_k A_kS with explicit D_k = A_kk^-1 was wrong in the third digit.  This script shows that the loss is
caused by the explicit inverse, not by the Schur complement: with W_k = A_kk^-1 A_kN obtained by carrying A_kN through
the pivoted Gauss-Jordan elimination of A_kk (a solve), the result is as accurate as the full-inverse cut at ~200x fewer
flops.  Reference = LU + iterative refinement with long-double residuals.  Output of this container (Re = 5000,
gamma = 1e4) is kept in profiles/schur_accuracy_r1.txt.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.linalg as sla, scipy.sparse as sp
from alfi_b200.synth.problem import build_problem
def gj_aug(M, R):
    M=M.copy(); R=R.copy(); n=M.shape[0]
    for j in range(n):
        p=j+np.argmax(np.abs(M[j:,j]))
        if p!=j: M[[j,p]]=M[[p,j]]; R[[j,p]]=R[[p,j]]
        piv=1.0/M[j,j]; M[j]*=piv; R[j]*=piv
        f=M[:,j].copy(); f[j]=0
        M-=np.outer(f,M[j]); R-=np.outer(f,R[j])
    return R
def study(A, blk, tag):
    n=A.shape[0]; S=np.flatnonzero(blk<0); labels=np.unique(blk[blk>=0]); B=[np.flatnonzero(blk==k) for k in labels]
    if S.size==0 or not B: return
    E=np.zeros((n,S.size)); E[S,np.arange(S.size)]=1
    lu=sla.lu_factor(A); X=sla.lu_solve(lu,E); Al=A.astype(np.longdouble); Xl=X.astype(np.longdouble)
    for it in range(3):
        R=E.astype(np.longdouble)-Al@Xl; Xl=Xl+sla.lu_solve(lu,np.asarray(R,dtype=np.float64)).astype(np.longdouble)
    Xref=np.asarray(Xl[S],dtype=np.float64)
    err=lambda Z: np.linalg.norm(Z-Xref)/np.linalg.norm(Xref)
    inv=np.linalg.inv(A)
    Sc=A[np.ix_(S,S)].copy()
    for b in B:
        Nk=np.flatnonzero(np.abs(A[np.ix_(b,S)]).sum(0)+np.abs(A[np.ix_(S,b)]).sum(1)>0)
        W=gj_aug(A[np.ix_(b,b)],A[np.ix_(b,S[Nk])]); Sc[np.ix_(Nk,Nk)]-=A[np.ix_(S[Nk],b)]@W
    Xs=gj_aug(Sc,np.eye(S.size))
    print("%-28s n=%4d |S|=%3d blocks=%2d cond(A)=%.1e  full-inverse cut %.1e   Schur/GJ-solve %.1e" % (tag,n,S.size,len(B),np.linalg.cond(A),err(inv[np.ix_(S,S)]),err(Xs)),flush=True)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for name,kw in [("ldc3d-sv-k3-small",{}),("ldc2d-sv-k2",{}),("bfs2d-sv-k2-small",{})]:
    prob=build_problem(name,**kw)
    ld=prob.levels[-1]; ps=ld.patches
    A=sp.bsr_matrix((ld.A.vals,ld.A.colidx,ld.A.rowptr),shape=(ld.V.ndofs,ld.V.ndofs)).tocsr()
    sizes=np.unique(ps.sizes)
    for sz in sizes[-4:]:
        p=int(np.flatnonzero(ps.sizes==sz)[len(np.flatnonzero(ps.sizes==sz))//2]); I=ps.dofs[ps.offsets[p]:ps.offsets[p+1]]
        study(A[I][:,I].toarray(), ps.blocks[ps.offsets[p]:ps.offsets[p+1]], "%s %s patch %d"%(name,kw or "",p))
    # transfer cell patches (A0)
    cp=ld.cell_patches
    if cp is not None and cp.blocks is not None:
        A0=sp.bsr_matrix((ld.A0.vals,ld.A.colidx,ld.A.rowptr),shape=A.shape).tocsr()
        q=int(np.argmax(cp.sizes)); I=cp.dofs[cp.offsets[q]:cp.offsets[q+1]]
        study(A0[I][:,I].toarray(), cp.blocks[cp.offsets[q]:cp.offsets[q+1]], "%s cell patch %d"%(name,q))


def explicit_inverse_variant():
    """the variant that fails: D_k = inv(A_kk) explicitly, then A_Sk (D_k A_kS)"""
    prob = build_problem("ldc3d-sv-k3-small")
    ld = prob.levels[-1]; ps = ld.patches
    A = sp.bsr_matrix((ld.A.vals, ld.A.colidx, ld.A.rowptr), shape=(ld.V.ndofs, ld.V.ndofs)).tocsr()
    p = int(np.argmax(ps.sizes)); I = ps.dofs[ps.offsets[p]:ps.offsets[p + 1]]; blk = ps.blocks[ps.offsets[p]:ps.offsets[p + 1]]
    Ap = A[I][:, I].toarray()
    S = np.flatnonzero(blk < 0); B = [np.flatnonzero(blk == k) for k in np.unique(blk[blk >= 0])]
    Sc = Ap[np.ix_(S, S)].copy()
    for b in B:
        Sc -= Ap[np.ix_(S, b)] @ (np.linalg.inv(Ap[np.ix_(b, b)]) @ Ap[np.ix_(b, S)])
    X = np.linalg.inv(Sc); ref = np.linalg.inv(Ap)[np.ix_(S, S)]
    print("explicit D_k in the Schur complement, 1275-dof patch: rel diff to the full-inverse cut %.1e (cond(S_c) %.1e, cond(A_kk) %.1e)"
          % (np.linalg.norm(X - ref) / np.linalg.norm(ref), np.linalg.cond(Sc), np.linalg.cond(Ap[np.ix_(B[0], B[0])])))


explicit_inverse_variant()
