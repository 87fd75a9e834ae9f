"""Patch index sets against the REFERENCE'S OWN CODE.

tests/golden/reference_index_sets.npz holds what alfi/relaxation.py and alfi/transfer.py themselves produce
(loaded from the reference tree by oracle/refshim.py over stand-ins for Firedrake / petsc4py, see
tests/golden/make_reference_golden.py) on five synthetic hierarchies.  alfi_b200's plugin classes and its
vectorised builders must reproduce them: patch point lists exactly (order and duplicates included) for the
per-entity callbacks, as sets for the vectorised CSR builders (PCPATCH puts the points into a hash set),
iteration sets and coarse-boundary node lists exactly.  Where the reference tree is present the fixture is also
regenerated and must not have drifted.
"""
import os
import sys

import numpy as np
import pytest

from alfi_b200.patches import points_to_csr
from alfi_b200.relaxation import MacroStar, Star, macro_star_points, star_points
from alfi_b200.synth.fem import VectorSpace
from alfi_b200.transfer import (CoarseCellMacroPatches, CoarseCellPatches, coarse_cell_points,
                                fix_coarse_boundaries)
from oracle import refshim
from tests.test_patches import Ctx, FakePC

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_golden as gen  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_index_sets.npz"))


def unflatten(off, data):
    return [data[off[i]:off[i + 1]] for i in range(off.size - 1)]


def csr_of(sets, npoints):
    return points_to_csr([np.asarray(s) for s in sets], npoints)


@pytest.mark.parametrize("name", list(gen.CASES))
def test_patch_constructors_reproduce_the_reference(name):
    build, k, kind, sort = gen.CASES[name]
    levels = build()
    bary = levels[0].bary
    g = {key.split("/", 1)[1]: GOLD[key] for key in GOLD.files if key.startswith(name + "/")}
    for l, lev in enumerate(levels):
        plex = lev.plex
        # ---- Star (relaxation.py:153-160): callback class and vectorised builder
        ref = unflatten(g["l%d_star_off" % l], g["l%d_star_pts" % l])
        patches, order = Star()(FakePC(plex))
        assert len(patches) == len(ref) and all(np.array_equal(a, b) for a, b in zip(patches, ref))
        assert np.array_equal(order, g["l%d_star_iter" % l])
        H, _ = star_points(plex)
        assert (csr_of(ref, plex.npoints) != H).nnz == 0
        # ---- MacroStar (relaxation.py:163-177) with the case's sort order
        if bary:
            ref = unflatten(g["l%d_macro_off" % l], g["l%d_macro_pts" % l])
            opts = {} if sort is None else {"pc_patch_construction_MacroStar_sort_order": sort}
            patches, order = MacroStar()(FakePC(plex, opts))
            assert len(patches) == len(ref) and all(np.array_equal(a, b) for a, b in zip(patches, ref))
            assert np.array_equal(order, g["l%d_macro_iter" % l])
            H, _ = macro_star_points(plex, "all")
            assert (csr_of(ref, plex.npoints) != H).nnz == 0
        # ---- transfer patches and coarse-boundary nodes (transfer.py:13-88, 121-158)
        if l > 0:
            ref = unflatten(g["l%d_cell_off" % l], g["l%d_cell_pts" % l])
            maker = CoarseCellMacroPatches() if bary else CoarseCellPatches()
            patches, order = maker(FakePC(plex, ctx=Ctx(levels, l)))
            assert len(patches) == len(ref) and all(np.array_equal(a, b) for a, b in zip(patches, ref))
            assert np.array_equal(order, g["l%d_cell_iter" % l])
            assert (csr_of(ref, plex.npoints) != coarse_cell_points(levels, l, bary)).nnz == 0
            V = VectorSpace(lev.mesh, k, kind)
            assert np.array_equal(fix_coarse_boundaries(plex, V, l), g["l%d_cb_nodes" % l])


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", list(gen.CASES))
def test_fixture_is_what_the_reference_code_produces_now(name):
    """Re-run alfi's own relaxation.py / transfer.py and compare with the committed fixture."""
    out = gen.run_reference(name)
    keys = sorted(k.split("/", 1)[1] for k in GOLD.files if k.startswith(name + "/"))
    assert sorted(out) == keys
    for key in keys:
        assert np.array_equal(out[key], GOLD[name + "/" + key]), key
    assert "firedrake" not in sys.modules and "alfi" not in sys.modules        # the stand-ins do not leak


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_sort_order_parsing_is_the_reference_one():
    """keyfuncs (relaxation.py:88-108) on the key syntax of the examples ("0+:1-", bfs2d.py:32) and sweeps ("|")."""
    from alfi_b200.relaxation import iteration_order
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 4, size=(40, 3)).astype(float)          # many ties: stability matters
    with refshim.reference_modules() as (rel, _):
        for spec in ("0+:1-", "1-", "2+:0-:1+", "0+:1-|1+:0-", "0", "None", ""):
            refshim.set_options({"pc_patch_construction_MacroStar_sort_order": spec})
            ms = rel.MacroStar()
            ms.opts = refshim.Options("")
            kf = ms.keyfuncs(list(enumerate(coords)))
            # no key functions -> createStride(len(patches)) (relaxation.py:140-143), else concatenated sorts (:145-149)
            want = list(range(len(coords))) if kf is None else \
                sum(([i for i, _ in sorted(enumerate(coords), key=f)] for f in kf), [])
            assert np.array_equal(iteration_order(coords, spec), want), spec
    refshim.set_options({})


# ------------------------------------------------------------------------------------ solver dictionaries
import json  # noqa: E402

import alfi_b200  # noqa: E402
from alfi_b200 import pc as pcmod  # noqa: E402
from alfi_b200.pc import fieldsplit0_config  # noqa: E402

sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_parameters as genp  # noqa: E402

PARAMS = json.load(open(os.path.join(HERE, "golden", "reference_parameters.json")))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", list(genp.VARIANTS))
def test_parameter_fixture_is_what_get_parameters_returns_now(name):
    assert json.loads(json.dumps(genp.run(name), sort_keys=True)) == PARAMS[name]


@pytest.mark.parametrize("name,m,construct,sort", [
    ("ldc2d-sv-k2", 6, "alfi.MacroStar", "0+:1-"), ("bfs2d-sv-k2", 6, "alfi.MacroStar", "0+:1-"),
    ("ldc3d-sv-k3", 10, "alfi.MacroStar", "0+:1-"), ("ldc2d-pkp0", 6, "star", None), ("ldc3d-pkp0", 10, "star", None)])
def test_fieldsplit0_dictionary_is_accepted_unchanged(name, m, construct, sort):
    """The reference's own fieldsplit_0 dictionary (solver.py:359-379) configures the device cycle."""
    outer = PARAMS[name]["outer"]
    assert outer["pc_fieldsplit_type"] == "schur" and outer["ksp_type"] == "fgmres"
    cfg = fieldsplit0_config(outer["fieldsplit_0"])
    assert cfg["smoothing"] == m == PARAMS[name]["smoothing"]
    assert cfg["construct"] == construct and cfg["sort_order"] == sort
    assert PARAMS[name]["firedrake_parameters"]["default_sub_matrix_type"] == "baij"      # solver.py:512 -> BSR on the device
    from alfi_b200.synth.problem import CONFIGS
    c = CONFIGS[name]
    assert c.m == m and c.sort_order in (sort, None) and (c.patch == "macro") == (construct == "alfi.MacroStar")


@pytest.mark.parametrize("name", ["ldc3d-sv-k3-multiplicative", "ldc2d-pkp0-star-multiplicative"])
def test_multiplicative_dictionaries_are_read(name):
    """`--patch-composition multiplicative` (solver.py:306-308): symmetrised sequential sweeps in the relaxation direction."""
    got = fieldsplit0_config(PARAMS[name]["outer"]["fieldsplit_0"])
    assert got["local_type"] == "multiplicative" and got["symmetrise_sweep"] and got["sort_order"] == "0+:1-"


def test_unsupported_dictionaries_are_refused():
    import copy
    fs0 = copy.deepcopy(PARAMS["ldc3d-sv-k3"]["outer"]["fieldsplit_0"])
    fs0["mg_levels"]["patch_pc_patch_partition_of_unity"] = True
    with pytest.raises(NotImplementedError, match="partition_of_unity"):
        fieldsplit0_config(fs0)


class _Recorder:
    def __init__(self, *a, **k):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append((name, a, k))
            return 0
        return f


@pytest.mark.parametrize("config,problem,level", [("ldc2d-sv-k2", "ldc2d-sv-k2-tiny", 1), ("ldc2d-pkp0", "ldc2d-pkp0-tiny", 2),
                                                  ("bfs2d-sv-k2", "bfs2d-sv-k2-tiny", 1)])
def test_patchpc_takes_the_reference_mg_levels_options(problems, monkeypatch, config, problem, level):
    """mg_levels of the reference's dictionary, flattened to PETSc option names the way Firedrake does and
    put under the level prefix, is all alfi_b200.PatchPC needs: it builds the problem's patch dof sets and
    iteration order (the CUDA context is replaced by a recorder)."""
    from alfi_b200.synth.fakepetsc import FakePC as SynthPC, SynthAdapter
    monkeypatch.setattr(pcmod, "Context", _Recorder)
    prefix = "fieldsplit_0_mg_levels_%d_" % level
    opts = refshim.flatten_options(PARAMS[config]["outer"]["fieldsplit_0"]["mg_levels"], prefix)
    assert opts[prefix + "pc_python_type"] == "firedrake.PatchPC"           # the one string a maintainer changes
    prob = problems(problem, gamma=10.0, nu=0.2)
    ad = SynthAdapter(prob, level)
    pc = SynthPC(prob.levels[level].level.plex, options=opts, prefix=prefix, attrs={"alfi_b200_adapter": ad})
    p = alfi_b200.PatchPC()
    p.initialize(pc)
    ref = prob.levels[level].patches
    assert np.array_equal(p.patches.offsets, ref.offsets) and np.array_equal(p.patches.dofs, ref.dofs)
    assert np.array_equal(p.patches.order, ref.order)
    assert p.options_seen["pc_patch_sub_mat_type"] in ("seqaij", "seqdense")
    assert [c[0] for c in p.ctx.calls] == ["level_create", "set_bsr_pattern", "set_bc", "set_patches", "set_bsr_values", "factor"]


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_dg_mass_inv_is_the_reference_one():
    """solver.py:15-38: the reference's DGMassInv.apply on stand-in Mat/Vec objects == the continuation stand-in's
    Schur complement approximation (alfi_b200/synth/outer.py)."""
    import types

    import scipy.sparse as sp
    from alfi_b200.synth.outer import dg_mass_inv_apply
    rng = np.random.default_rng(0)
    Minv = sp.random(30, 30, density=0.2, random_state=1, format="csr") + sp.identity(30)
    x = rng.standard_normal(30)

    class Vec:
        def __init__(self, a):
            self.array = np.array(a, dtype=float)

        def scale(self, k):
            self.array *= k

    def mult(xv, yv):
        yv.array[:] = Minv @ xv.array
    with refshim.reference_modules(with_solver=True) as (_, _, sol):
        pc = sol.DGMassInv.__new__(sol.DGMassInv)
        pc.massinv = types.SimpleNamespace(mult=mult)
        pc.nu, pc.gamma = 0.004, 1.0e4
        y = Vec(np.zeros(30))
        pc.apply(None, Vec(x), y)
        with pytest.raises(NotImplementedError):
            pc.applyTranspose(None, None, None)
    assert np.array_equal(y.array, dg_mass_inv_apply(Minv, 0.004, 1.0e4, x))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("case", ["kuhn2d", "kuhn3d", "step"])
def test_bary_hierarchy_cell_maps_are_the_reference_ones(case):
    """alfi/bary.py:29-194 executed over stand-ins (oracle/refshim_bary.py): every level is Alfeld-split after
    uniform refinement with all uniform vertices labelled MacroVertices, the hierarchy is not nested, and the
    coarse-to-fine table of the barycentric meshes composed by the reference equals the synthetic one
    (alfi_b200/synth/hierarchy.py) — the table CoarseCellMacroPatches and the standard prolongation consume."""
    from fractions import Fraction

    from alfi_b200.synth.gmsh import step_mesh
    from alfi_b200.synth.hierarchy import build_hierarchy, build_hierarchy_from
    from oracle.refshim_bary import World
    levels = {"kuhn2d": lambda: build_hierarchy(2, 2, 2, True), "kuhn3d": lambda: build_hierarchy(3, 1, 1, True),
              "step": lambda: build_hierarchy_from(step_mesh(1, seed=3), 2, True)}[case]()
    nref = len(levels) - 1
    w = World(levels)
    with w.reference_function() as BaryMeshHierarchy:
        mh = BaryMeshHierarchy(w.base_mesh(), nref)
    d = levels[0].macro.dim
    assert mh.nested is False and len(mh.meshes) == nref + 1 and w.alfeld_applied == list(range(nref + 1))
    assert [m._topology_dm.refine_level for m in mh.meshes] == list(range(nref + 1))
    for l in range(nref):
        c2f = mh.coarse_to_fine_cells[Fraction(l, 1)]
        assert c2f.shape == (levels[l].mesh.nc, (d + 1) * 2 ** d)
        assert np.array_equal(c2f, levels[l].c2f)
        f2c = mh.fine_to_coarse_cells[Fraction(l + 1, 1)]
        assert f2c.shape == (levels[l + 1].mesh.nc, d + 1)
        # every fine bary cell lists the d+1 bary cells of its coarse macro cell
        parent = np.empty(levels[l + 1].macro.nc, dtype=np.int64)
        parent[levels[l].macro_c2f.ravel()] = np.repeat(np.arange(levels[l].macro.nc), 2 ** d)
        want = parent[np.arange(levels[l + 1].mesh.nc) // (d + 1)][:, None] * (d + 1) + np.arange(d + 1)[None, :]
        assert np.array_equal(f2c, want)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,solver", [("ldc2d-sv-k2-tiny", "ScottVogeliusSolver"), ("ldc2d-pkp0-tiny", "ConstantPressureSolver"),
                                         ("ldc3d-sv-k3-tiny", "ScottVogeliusSolver"), ("ldc3d-pkp0-tiny", "ConstantPressureSolver"),
                                         ("bfs2d-sv-k2-tiny", "ScottVogeliusSolver")])
def test_velocity_operator_is_the_reference_form(problems, name, solver):
    """Row M1's operator: the reference's `residual()` (solver.py:562-572, 613-623), evaluated numerically by
    oracle/ufl_eval.py on the synthetic mesh and element, against the parts alfi_b200.synth.fem assembles and
    hands to the library — nu (2 sym grad u, grad v) + gamma (div u, div v) [cell_avg(div u) for pkp0] exactly,
    and the Newton linearisation of advect ((grad u) u, v) through N(u+d) - N(u) - N(d)."""
    from alfi_b200.synth.fem import BSR, assemble_parts
    from oracle.ufl_eval import reference_velocity_residual
    prob = problems(name, gamma=10.0, nu=0.2)
    ld = prob.finest
    V = ld.V
    rng = np.random.default_rng(0)
    U, D = rng.standard_normal((V.nnodes, V.bs)), rng.standard_normal((V.nnodes, V.bs))
    nu, gamma = 0.3, 7.0
    mat = lambda vals: BSR(V.nnodes, V.bs, ld.pattern.rowptr, ld.pattern.colidx, vals).to_csr()      # noqa: E731
    parts = assemble_parts(V, ld.pattern, None, prob.config.discretisation, want=("visc", "div"))
    want = mat(nu * parts["visc"] + gamma * parts["div"]) @ U.ravel()
    got = reference_velocity_residual(solver, V, U, nu, gamma, 0.0)
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)
    N = lambda W: reference_velocity_residual(solver, V, W, 0.0, 0.0, 1.0)                           # noqa: E731
    adv = assemble_parts(V, ld.pattern, U, prob.config.discretisation, want=("adv1", "adv2"))
    want = mat(adv["adv1"] + adv["adv2"]) @ D.ravel()
    got = N(U + D) - N(U) - N(D)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,solver", [("ldc2d-sv-k2-tiny", "ScottVogeliusSolver"), ("ldc2d-pkp0-tiny", "ConstantPressureSolver"),
                                         ("ldc3d-sv-k3-tiny", "ScottVogeliusSolver")])
def test_saddle_point_residual_is_the_reference_form(problems, name, solver):
    """The whole residual of solver.py:562-572 / 613-623 with a pressure: F_u = A(u) u + B^T p, F_p = B u with
    B = -(div u, q) as alfi_b200.synth.fem.assemble_divergence builds it for the continuation stand-in
    (alfi_b200/synth/outer.py), and the pressure mass matrix DGMassInv inverts."""
    from alfi_b200.synth.fem import BSR, assemble_divergence, assemble_parts
    from oracle.ufl_eval import reference_residual
    prob = problems(name, gamma=10.0, nu=0.2)
    cfg = prob.config
    ld = prob.finest
    V = ld.V
    kq = cfg.k - 1 if cfg.discretisation == "sv" else 0
    B, Minv = assemble_divergence(V, kq)
    rng = np.random.default_rng(1)
    U, P = rng.standard_normal((V.nnodes, V.bs)), rng.standard_normal(B.shape[0])
    nu, gamma = 0.3, 7.0
    Fu, Fp = reference_residual(solver, V, kq, U, P, nu, gamma, 1.0)
    mat = lambda vals: BSR(V.nnodes, V.bs, ld.pattern.rowptr, ld.pattern.colidx, vals).to_csr()      # noqa: E731
    parts = assemble_parts(V, ld.pattern, U, cfg.discretisation, want=("visc", "div", "adv1"))
    want_u = mat(nu * parts["visc"] + gamma * parts["div"] + parts["adv1"]) @ U.ravel() + B.T @ P
    assert np.linalg.norm(Fu - want_u) <= 1e-12 * np.linalg.norm(want_u)
    assert np.linalg.norm(Fp - B @ U.ravel()) <= 1e-12 * np.linalg.norm(Fp)
    # Minv is the inverse of (p, q): applying the mass form to Minv's columns gives the identity on a cell
    _, Fq = reference_residual(solver, V, kq, np.zeros_like(U), P, 0.0, 0.0, 0.0)
    assert np.abs(Fq).max() == 0.0                       # no (p, q) term in the residual: mass only enters DGMassInv


@pytest.mark.parametrize("name,tdim", [("ldc2d-sv-k2", 2), ("ldc2d-pkp0", 2), ("ldc3d-sv-k3", 3), ("ldc3d-pkp0", 3)])
def test_continuation_tolerances_are_the_reference_ones(name, tdim):
    """The outer Newton / FGMRES settings of the continuation stand-in == the reference's dictionary."""
    from alfi_b200.synth import outer
    ref = PARAMS[name]["outer"]
    tol = outer.tolerances(tdim)
    for key in ("ksp_rtol", "ksp_atol", "snes_rtol", "snes_atol"):
        assert tol[key] == ref[key], key
    assert ref["snes_max_it"] == outer.SNES_MAX_IT and ref["ksp_max_it"] == outer.KSP_MAX_IT
    assert ref["ksp_type"] == "fgmres" and ref["snes_type"] == "newtonls" and ref["snes_linesearch_type"] == "basic"
    assert ref["pc_fieldsplit_schur_factorization_type"] == "full" and ref["pc_fieldsplit_schur_precondition"] == "user"
    assert ref["fieldsplit_1"] == {"ksp_type": "preonly", "pc_type": "python", "pc_python_type": "alfi.solver.DGMassInv"}


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_driver_defaults_and_reynolds_rule():
    """driver.get_default_parser (driver.py:9-51) and NavierStokesSolver.solve's parameter update (solver.py:257-268)
    executed from the reference tree: gamma, composition, smoothing defaults and nu = L U / Re, advect = [Re > 0]."""
    import importlib.util
    import types

    from alfi_b200.synth.problem import CONFIGS
    extra = {"alfi.solver": refshim._module("alfi.solver", ConstantPressureSolver=None, ScottVogeliusSolver=None)}
    with refshim.reference_modules(with_solver=True, extra_modules=extra) as (_, _, sol):
        spec = importlib.util.spec_from_file_location("_alfi_reference_driver", os.path.join(refshim.REFERENCE, "alfi", "driver.py"))
        drv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(drv)
        args = drv.get_default_parser().parse_args(["--discretisation", "sv"])
        assert args.gamma == CONFIGS["ldc3d-sv-k3"].gamma == 1e4
        assert args.patch_composition == "additive" and args.solver_type == "almg" and args.smoothing is None
        assert args.restriction is False and args.high_accuracy is False

        class Const:
            def __init__(self):
                self.v = None

            def assign(self, v):
                self.v = float(v)

            def values(self):
                return [self.v]
        for re in (0, 10, 5000):
            me = types.SimpleNamespace(
                z_last=types.SimpleNamespace(assign=lambda z: None), z=None, message=lambda m: None,
                advect=Const(), nu=Const(), char_L=2.0, char_U=1.0, stabilisation=None, nsp=None,
                check_nograddiv_residual=False,
                solver=types.SimpleNamespace(solve=lambda: None, snes=types.SimpleNamespace(
                    getLinearSolveIterations=lambda: 7, getIterationNumber=lambda: 2)))
            sol.GREEN = "%s"
            _, info = sol.NavierStokesSolver.solve(me, re)
            cfg = CONFIGS["ldc2d-sv-k2"]
            want_nu = cfg.length * 1.0 / re if re > 0 else cfg.length          # alfi_b200/synth/outer.py
            assert me.nu.v == want_nu == info["nu"] and me.advect.v == (1.0 if re > 0 else 0.0)
            assert info["linear_iter"] == 7 and info["nonlinear_iter"] == 2 and info["Re"] == re


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_transfer_wiring_is_the_reference_one():
    """Row W1: ScottVogeliusSolver.get_transfers (solver.py:632-653) and NullTransfer (transfer.py:359-366)
    executed from the reference tree — which callables go to the TransferManager for the velocity and the pressure
    element, with and without --restriction — against alfi_b200's drop-in classes."""
    import types

    import alfi_b200
    prolong, restrict, inject = object(), object(), object()
    with refshim.reference_modules(with_solver=True, extra_firedrake=dict(prolong=prolong, restrict=restrict, inject=inject)) as (_, tr, sol):
        for restriction in (True, False):
            V = types.SimpleNamespace(ufl_element=lambda: "V-element")
            Q = types.SimpleNamespace(ufl_element=lambda: "Q-element")
            me = types.SimpleNamespace(Z=types.SimpleNamespace(sub=lambda i: (V, Q)[i]), stabilisation_type=None,
                                       hierarchy="bary", nu=0.1, gamma=1e4, tdim=3, restriction=restriction)
            transfers = sol.ScottVogeliusSolver.get_transfers(me)
            vp, vr, vi = transfers["V-element"]
            assert isinstance(me.vtransfer, sol.SVSchoeberlTransfer) and me.vtransfer.parameters == (0.1, 1e4)
            assert vp == me.vtransfer.prolong and vi is inject
            assert (vr == me.vtransfer.restrict) if restriction else (vr is restrict)
            qp, qr, qi = transfers["Q-element"]
            assert qp is prolong and qr is restrict and isinstance(me.qtransfer, tr.NullTransfer) and qi == me.qtransfer.inject
            # the patch solver of the transfer: same option dictionary as alfi_b200 honours
            pp = me.vtransfer.patchparams
            assert pp["patch_pc_patch_construct_python_type"] == "alfi.transfer.CoarseCellMacroPatches"
            assert pp["patch_pc_patch_partition_of_unity"] is False and pp["patch_sub_pc_type"] == "lu"
            assert pp["patch_pc_patch_sub_mat_type"] == "seqaij"

        class Vec:
            def __init__(self):
                self.a = np.zeros(4)

            def set(self, v):
                self.a[:] = v
        dest = types.SimpleNamespace(dat=types.SimpleNamespace(vec_wo=None))
        import contextlib
        v = Vec()
        dest.dat.vec_wo = contextlib.nullcontext(v)
        tr.NullTransfer().inject(None, dest)
        assert np.isnan(v.a).all()
    ours = np.zeros(4)
    alfi_b200.NullTransfer().inject(None, ours)
    assert np.isnan(ours).all()
    for name in ("prolong", "restrict", "inject"):
        assert callable(getattr(alfi_b200.NullTransfer(), name)) and callable(getattr(alfi_b200.SVSchoeberlTransfer, name, None) or getattr(alfi_b200.AutoSchoeberlTransfer, name, None) or (lambda: 0))
