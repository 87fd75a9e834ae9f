"""Burman's interior-facet stabilisation (alfi/stabilisation.py:139-162; --stabilisation-type burman, what the
reference's Scott-Vogelius jobs run with: examples/Makefile:12-16): PCPATCH's patch operators are then NOT sub-matrices
of the assembled operator — PCPATCH integrates over the patch cells and the facets whose both cells are patch
cells, the assembled operator also holds the inside-inside part of the facets on the patch boundary.

CPU: the facet tensors; the hand-over data A_i = A[I_i, I_i] + C_i against a literal patch-by-patch assembly; the oracle
continuation.  GPU: alfib_level_set_patch_corrections through the C-ABI against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.synth.fem import (FacetBlockPattern, VectorSpace, burman_facet_tensors, element_matrices, facet_adjacency)
from alfi_b200.synth.mesh import alfeld_split, kuhn_mesh
from alfi_b200.synth.problem import CONFIGS, lid_wind
from oracle import hotpath as hp

BURMAN = ["ldc2d-sv-k2-tiny-burman", "ldc2d-pkp0-tiny-burman", "ldc3d-sv-k3-tiny-burman"]


@pytest.mark.parametrize("dim,k", [(2, 2), (2, 3), (3, 3)])
def test_jump_form_vanishes_on_smooth_functions_and_is_psd(dim, k):
    """dot(jump(grad u, n), jump(grad v, n)) dS: every globally polynomial u of degree <= k lies in the space and has a
    continuous gradient, so it is in the kernel; the matrix is symmetric positive semi-definite and not zero."""
    mesh = alfeld_split(kuhn_mesh(dim, 2))
    mesh.build_topology()
    V = VectorSpace(mesh, k)
    rng = np.random.default_rng(0)
    _, fc, S = burman_facet_tensors(V, rng.standard_normal((V.nnodes, dim)), 5e-3)
    pat = FacetBlockPattern(V)
    assert np.array_equal(fc, pat.facet_cells)
    A = sp.csr_matrix((pat.scatter_facets(S), pat.colidx, pat.rowptr), shape=(V.nnodes,) * 2)
    x = V.node_coords
    scale = abs(A).max()
    for f in (np.ones(V.nnodes), x[:, 0], x[:, -1] - 2 * x[:, 0], x[:, 0] ** 2, x[:, 0] * x[:, -1], x[:, -1] ** k):
        assert np.abs(A @ f).max() <= 1e-11 * scale * max(np.abs(f).max(), 1.0)
    assert abs(A - A.T).max() <= 1e-14 * scale
    assert np.linalg.eigvalsh(A.toarray()).min() >= -1e-12 * scale
    r = rng.standard_normal(V.nnodes)
    assert r @ (A @ r) > 1e-3 * scale * (r @ r) / V.nnodes


def test_facet_adjacency_is_consistent():
    mesh = kuhn_mesh(3, 2)
    mesh.build_topology()
    fid, fc, fj = facet_adjacency(mesh)
    assert fid.size == np.unique(fid).size and (fc[:, 0] != fc[:, 1]).all()
    for a in range(2):
        assert np.array_equal(mesh.cell_facets[fc[:, a], fj[:, a]], fid)
    assert fid.size + (np.bincount(mesh.cell_facets.ravel()) == 1).sum() == mesh.nf


@pytest.mark.parametrize("name", BURMAN)
def test_patch_operators_are_the_pcpatch_integrals(problems, name):
    """Literal PCPATCH semantics, patch by patch: element tensors of the patch cells + facet tensors of the facets whose
    both cells are patch cells, restricted to the patch dofs — equals A[I, I] + C; A[I, I] alone does not."""
    prob = problems(name, gamma=10.0, nu=0.2)
    cfg = prob.config
    differs = 0
    for ld in prob.levels[1:]:
        V, ps, bs = ld.V, ld.patches, ld.V.bs
        nl = V.cell_nodes.shape[1]
        wind = V.interpolate(lambda xx: lid_wind(xx, V.mesh.extent))
        fid, fc, S = burman_facet_tensors(V, wind, cfg.stab_weight)
        corr = (ps.corrections.off, ps.corrections.rows, ps.corrections.cols, ps.corr_vals)
        A = ld.A.to_csr()
        got = hp.patch_matrices(A, ps.offsets, ps.dofs, corr)
        plain = hp.patch_matrices(A, ps.offsets, ps.dofs)
        Hc = ps.cells.tocsr()
        for p in range(ps.npatch):
            I = ps.patch(p)
            if I.size == 0:
                continue
            pos = {int(g): i for i, g in enumerate(I)}
            cells = Hc.indices[Hc.indptr[p]:Hc.indptr[p + 1]]
            M = np.zeros((I.size, I.size))
            E = element_matrices(V, prob.nu, prob.gamma, wind, 1.0, cfg.discretisation, cells)      # (nc, nl, d, nl, d)
            for ci, c in enumerate(cells):
                for i, ni in enumerate(V.cell_nodes[c]):
                    for j, nj in enumerate(V.cell_nodes[c]):
                        for r in range(bs):
                            for s_ in range(bs):
                                a, b = pos.get(int(ni) * bs + r), pos.get(int(nj) * bs + s_)
                                if a is not None and b is not None:
                                    M[a, b] += E[ci, i, r, j, s_]
            inpatch = np.isin(fc, cells).all(axis=1)
            for f in np.flatnonzero(inpatch):
                nodes = np.concatenate([V.cell_nodes[fc[f, 0]], V.cell_nodes[fc[f, 1]]])
                for i, ni in enumerate(nodes):
                    for j, nj in enumerate(nodes):
                        for r in range(bs):
                            a, b = pos.get(int(ni) * bs + r), pos.get(int(nj) * bs + r)
                            if a is not None and b is not None:
                                M[a, b] += S[f, i, j]
            assert np.abs(got[p] - M).max() <= 1e-11 * np.abs(M).max(), (name, ld.index, p)
            differs += np.abs(plain[p] - M).max() > 1e-6 * np.abs(M).max()
            assert 2 * nl >= 1
    assert differs > 0                          # the sub-matrix alone is NOT PCPATCH's operator


def test_oracle_continuation_with_burman():
    from alfi_b200.synth.outer import ContinuationSolver
    from oracle.backend import OracleBackend
    cfg = CONFIGS["ldc2d-sv-k2-tiny-burman"]
    s = ContinuationSolver(cfg, OracleBackend(cfg.m))
    for re in (10, 100):
        info = s.solve(re)
        assert info["nonlinear_iter"] <= 6 and info["linear_iter"] <= 12 * info["nonlinear_iter"]
        assert info["residual"] <= max(1e-8, 1e-9 * info["residual0"])
    assert np.linalg.norm(s.B @ s.u.ravel()) <= 1e-11


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", BURMAN)
def test_device_patch_inverses_and_cycle_with_corrections(problems, name):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems(name, gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    assert all(li.patch_corr_off is not None and li.patch_blocks is None for li in levels[1:])
    mg = DeviceMultigrid(levels, prob.config.m, deterministic=True)
    olv = [hp.level_from_host(l) for l in prob.levels]
    for l in range(1, len(levels)):
        ps = prob.levels[l].patches
        assert mg.ctx.patch_storage_form(l) == 0
        for p in (0, ps.npatch // 2, ps.npatch - 1):
            n = int(ps.sizes[p])
            if n == 0:
                continue
            X = mg.ctx.patch_inverse(l, p, n)
            want = olv[l].factors[p][1]
            assert np.abs(X - want).max() <= 1e-10 * np.abs(want).max(), (l, p)
        rng = np.random.default_rng(20261017 + l)
        x = rng.standard_normal(olv[l].n)
        x[olv[l].bc_dofs] = 0.0
        y = mg.ctx.smoother_apply(l, x, np.empty_like(x))
        yo = hp.smoother_apply(x, olv[l].offsets, olv[l].dofs, olv[l].order, olv[l].factors, olv[l].bc_dofs)
        assert np.linalg.norm(y - yo) <= 1e-11 * np.linalg.norm(yo)
        # and the corrections matter: the sub-matrix inverses give a different smoother
        plain = hp.factor_patches(hp.patch_matrices(olv[l].A, olv[l].offsets, olv[l].dofs))
        yp = hp.smoother_apply(x, olv[l].offsets, olv[l].dofs, olv[l].order, plain, olv[l].bc_dofs)
        assert np.linalg.norm(yp - yo) >= 1e-6 * np.linalg.norm(yo)
    b = np.random.default_rng(3).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0.0
    x = mg.apply(b, np.empty_like(b))
    xo = hp.fcycle(olv, b, prob.config.m)
    assert np.linalg.norm(x - xo) <= 1e-11 * np.linalg.norm(xo)


@pytest.mark.gpu
def test_stale_or_misplaced_corrections_are_refused(problems):
    from alfi_b200.lib import AlfibError
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems("ldc2d-sv-k2-tiny-burman", gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    mg = DeviceMultigrid(levels, prob.config.m, deterministic=True)
    mg.ctx.set_bsr_values(1, levels[1].vals)                   # new values, corrections not re-sent
    with pytest.raises(AlfibError, match="correction"):
        mg.ctx.factor(1)
    mg.ctx.set_patch_correction_values(1, levels[1].patch_corr_vals)
    mg.ctx.factor(1)
    li = levels[1]
    bad_rows = li.patch_corr_rows.copy()
    bad_rows[0] = 10 ** 6
    with pytest.raises(AlfibError, match="outside its patch"):
        mg.ctx.set_patch_corrections(1, li.patch_corr_off, bad_rows, li.patch_corr_cols)
    dup = li.patch_corr_rows.copy()
    dup[1], = dup[:1]
    cols = li.patch_corr_cols.copy()
    cols[1] = cols[0]
    with pytest.raises(AlfibError, match="distinct"):
        mg.ctx.set_patch_corrections(1, li.patch_corr_off, dup, cols)


@pytest.mark.gpu
def test_continuation_with_burman_on_gpu():
    from alfi_b200.multigrid import DeviceBackend
    from alfi_b200.synth.outer import ContinuationSolver
    from oracle.backend import OracleBackend
    cfg = CONFIGS["ldc2d-sv-k2-tiny-burman"]
    so = ContinuationSolver(cfg, OracleBackend(cfg.m))
    sd = ContinuationSolver(cfg, DeviceBackend(cfg.m, deterministic=True))
    for re in (10, 100, 200):
        a, b = so.solve(re), sd.solve(re)
        assert a["nonlinear_iter"] == b["nonlinear_iter"], (a, b)
        assert abs(a["linear_iter"] - b["linear_iter"]) <= a["nonlinear_iter"], (a, b)
    assert np.linalg.norm(sd.u - so.u) <= 1e-8 * np.linalg.norm(so.u)
    assert np.linalg.norm(sd.p - so.p) <= 1e-8 * np.linalg.norm(so.p)


@pytest.mark.gpu
def test_patchpc_with_corrections_through_the_adapter(problems):
    import alfi_b200
    from alfi_b200.synth.fakepetsc import FakePC, FakeVec, SynthAdapter
    prob = problems("ldc2d-sv-k2-tiny-burman", gamma=10.0, nu=0.2)
    opts = {"patch_pc_patch_construct_type": "python", "patch_pc_patch_construct_python_type": "alfi.MacroStar",
            "patch_pc_patch_construction_MacroStar_sort_order": "0+:1-", "patch_pc_patch_construction_MacroStar_expand": "vertices",
            "patch_sub_ksp_type": "preonly", "patch_sub_pc_type": "lu"}
    pc = FakePC(prob.levels[1].level.plex, options=opts, attrs={"alfi_b200_adapter": SynthAdapter(prob, 1, deterministic=True)})
    p = alfi_b200.PatchPC()
    p.setUp(pc)
    lv = hp.level_from_host(prob.levels[1])
    x = FakeVec(np.random.default_rng(0).standard_normal(lv.n))
    y = FakeVec(lv.n)
    for _ in range(2):                          # second pass: update() re-sends values and corrections
        p.apply(pc, x, y)
        want = hp.smoother_apply(x.array, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
        assert np.linalg.norm(y.array - want) <= 1e-11 * np.linalg.norm(want)
        p.setUp(pc)
