"""Synthetic lid-driven-cavity inputs for the velocity-block multigrid (host side).

Produces, for one of the BASELINE.json configurations, exactly what a Firedrake/alfi run would
hand to the hot path on every level: the BAIJ velocity operator (alfi/solver.py:512,562-572,
613-623), the Dirichlet node list (examples/ldc2d/ldc2d.py:23-26, ldc3d/ldc3d.py:17-20), the
smoother's patch dof sets (alfi/solver.py:331-344), the standard prolongation ``P_H`` and the
Schöberl-transfer operators (alfi/transfer.py:293-332).  Wind = nodal interpolant of the lid
profile extension, evaluated on every level (the reference injects it).
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field

import numpy as np

from ..patches import (PatchSet, facet_corrections, greedy_colouring, macro_interior_blocks, patch_dofs_from_points,
                       sweep_stages)
from ..relaxation import macro_star_points, star_points, iteration_order
from ..transfer import cell_patch_set
from .fem import (BSR, BlockPattern, FacetBlockPattern, VectorSpace, apply_dirichlet, assemble_parts, assemble_velocity_block,
                  burman_facet_tensors)
from .hierarchy import Level, build_hierarchy, build_hierarchy_from, prolongation_matrix

__all__ = ["Config", "CONFIGS", "LevelData", "Problem", "build_problem", "lid_wind"]


@dataclass(frozen=True)
class Config:
    name: str
    dim: int
    N: int
    nref: int
    discretisation: str          # "sv" | "pkp0"
    k: int
    patch: str                   # "star" | "macro"
    bary: bool
    re: float = 100.0
    gamma: float = 1.0e4
    smoothing: int | None = None
    macro_expand: str = "vertices"
    sort_order: str | None = None
    length: float = 2.0
    element: str = "lagrange"    # "lagrange" | "p1fb" ([P1+FacetBubble]^3, alfi/solver.py:576-579)
    domain: str = "ldc"          # "ldc": lid-driven cavity on [0, length]^d | "bfs": backward-facing step (2-D)
    mesh_file: str | None = None  # bfs: a Gmsh 2.2 file (examples/bfs2d/coarse*.msh); None = synth.gmsh.step_mesh(N)
    dirichlet_tags: tuple = (1, 2)  # bfs: Inflow + NoSlip (examples/bfs2d/bfs2d.py:25-27); Outflow stays natural
    shape: tuple = ()            # ldc: box [0, length * shape[a]] with N * shape[a] base cells per axis (() = cube)
    composition: str = "additive"   # "multiplicative": --patch-composition multiplicative (solver.py:306-308): sequential
                                    # sweep in the relaxation direction + the backward sweep (symmetrise_sweep)
    stabilisation: str = "none"     # "burman": --stabilisation-type burman (solver.py:226-228, stabilisation.py:139-162):
                                    # interior-facet jump term; the patch operators are then NOT sub-matrices
    stab_weight: float = 5e-3       # --stabilisation-weight of the reference's jobs (examples/Makefile:12-16)

    @property
    def m(self):
        return self.smoothing if self.smoothing is not None else (10 if self.dim > 2 else 6)

    @property
    def nu(self):
        return self.length * 1.0 / self.re          # char_L * char_U / Re (alfi/solver.py:267)


CONFIGS = {
    # BASELINE.json configs[0..4]; sizes per SURVEY §8d.  Macro-star patches are swept in the problem's
    # relaxation direction "0+:1-" (ldc2d.py:39, ldc3d.py:31 -> solver.py:342; confirmed by running the
    # reference's get_parameters, tests/test_reference_code.py); the builtin star construction has no order.
    "ldc2d-sv-k2": Config("ldc2d-sv-k2", 2, 10, 1, "sv", 2, "macro", True, re=1000.0, sort_order="0+:1-"),
    "ldc2d-pkp0": Config("ldc2d-pkp0", 2, 16, 3, "pkp0", 2, "star", False, re=10000.0),
    "ldc3d-sv-k3": Config("ldc3d-sv-k3", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    "ldc3d-pkp0": Config("ldc3d-pkp0", 3, 16, 2, "pkp0", 1, "star", False, re=5000.0, element="p1fb"),
    # configs[4] with the LITERAL 3-D MacroStar of the reference (relaxation.py:168-177 expands every non-macro point of
    # closure(star(v)), so in 3-D the macro edges on the link of v pull in the neighbouring macro cells: 2 175-dof
    # interior patches, 21-29 colours).  macro_expand="vertices" above is the open macro star (1 275 dofs)
    # that SURVEY §8 / BASELINE.md size the benchmark by — an extension, labelled so in bench.py's `config`.
    "ldc3d-sv-k3-literal": Config("ldc3d-sv-k3-literal", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-",
                                  macro_expand="all"),
    "ldc3d-sv-k3-half-literal": Config("ldc3d-sv-k3-half-literal", 3, 2, 2, "sv", 3, "macro", True, re=5000.0,
                                       sort_order="0+:1-", macro_expand="all"),
    "ldc3d-sv-k3-small-literal": Config("ldc3d-sv-k3-small-literal", 3, 2, 1, "sv", 3, "macro", True, re=5000.0,
                                        sort_order="0+:1-", macro_expand="all"),
    "ldc3d-sv-k3-tiny-literal": Config("ldc3d-sv-k3-tiny-literal", 3, 1, 1, "sv", 3, "macro", True, re=100.0,
                                       sort_order="0+:1-", macro_expand="all"),
    # larger members of configs[4]'s family: baseN 6 is the reference's own choice for ldc3d (generate_submission:75;
    # 4 899 531 dofs, 15 625 patches: its dense coarse inverse (78.9 k dofs) does not fit — needs the condensed coarse
    # inverse of scripts/r2_prep), baseN 5 (2 840 943 dofs, 9 261 patches, coarse 46 038 dofs) runs as is.  The host
    # generator needs ~33 / ~57 GB of RAM and minutes for them (cfg5: 17 GB, 73 s).
    "ldc3d-sv-k3-n5": Config("ldc3d-sv-k3-n5", 3, 5, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    "ldc3d-sv-k3-n6": Config("ldc3d-sv-k3-n6", 3, 6, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    # weak-scaling family of configs[4]: rank r of an sx x sy x sz rank grid gets one 16^3-cell brick (cfg5's finest
    # mesh: 1.46 M dofs per rank).  The coarsest mesh must keep >= 4 cells per direction: with 2 the F-cycle of this
    # unstabilised Re = 5000 problem AMPLIFIES a random right-hand side (measured in round 2: 2e4 on the 4 x 2 x 2
    # coarse mesh of a four-level 2 x 1 x 1 family, 8.5 on a 1-cube coarse mesh, against a reduction to 0.045 on cfg5).
    # So 2 and 4 ranks keep cfg5's three levels over an 8 x 4 x 4 / 8 x 8 x 4 coarse mesh (46 k / 92 k coarse dofs, held
    # as the condensed coarse inverse: 13 k / 26 k separator dofs), and 8 ranks use four levels over cfg5's own 4^3
    # coarse mesh instead of an 8^3 one (185 k dofs, 47 k separator dofs: no replicated dense factorisation of that).
    "ldc3d-sv-k3-w1": Config("ldc3d-sv-k3-w1", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    "ldc3d-sv-k3-w2": Config("ldc3d-sv-k3-w2", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-", shape=(2, 1, 1)),
    "ldc3d-sv-k3-w4": Config("ldc3d-sv-k3-w4", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-", shape=(2, 2, 1)),
    "ldc3d-sv-k3-w8": Config("ldc3d-sv-k3-w8", 3, 2, 3, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-", shape=(2, 2, 2)),
    # cfg5 itself as a 2 x 2 x 2 grid of 8^3-cell bricks (strong scaling at 8 ranks with rank-local generation and a
    # brick partition): base 2 cells per brick edge of length 1, so the domain is [0, 2]^3 and nu = 1 / 2500 = 2 / 5000
    "ldc3d-sv-k3-s8": Config("ldc3d-sv-k3-s8", 3, 2, 2, "sv", 3, "macro", True, re=2500.0, sort_order="0+:1-", length=1.0,
                             shape=(2, 2, 2)),
    # the same family at test size (one 4^3-cell brick per rank, three levels)
    "ldc3d-sv-k3-wtiny2": Config("ldc3d-sv-k3-wtiny2", 3, 1, 2, "sv", 3, "macro", True, re=100.0, sort_order="0+:1-", shape=(2, 1, 1)),
    # scaled-down members of the same families (tests, smoke, CPU-baseline sample)
    "ldc2d-sv-k2-tiny": Config("ldc2d-sv-k2-tiny", 2, 2, 1, "sv", 2, "macro", True, re=100.0, sort_order="0+:1-"),
    # --stabilisation-type burman --stabilisation-weight 5e-3: what the reference's own Scott-Vogelius jobs run with
    # (examples/Makefile:12-16, 24-25; generate_submission:77).  Patch operators are not sub-matrices, the
    # macro-cell interiors couple across macro faces: dense patch inverses + patch corrections.
    "ldc2d-sv-k2-burman": Config("ldc2d-sv-k2-burman", 2, 10, 1, "sv", 2, "macro", True, re=1000.0, sort_order="0+:1-",
                                 stabilisation="burman"),
    "ldc2d-sv-k2-tiny-burman": Config("ldc2d-sv-k2-tiny-burman", 2, 2, 1, "sv", 2, "macro", True, re=100.0, sort_order="0+:1-",
                                      stabilisation="burman"),
    "ldc3d-sv-k3-tiny-burman": Config("ldc3d-sv-k3-tiny-burman", 3, 1, 1, "sv", 3, "macro", True, re=100.0, sort_order="0+:1-",
                                      stabilisation="burman"),
    "ldc3d-sv-k3-small-burman": Config("ldc3d-sv-k3-small-burman", 3, 2, 1, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-",
                                       stabilisation="burman"),
    # configs[4] as the reference's own job runs it (examples/generate_submission:69-87: burman, weight 5e-3)
    "ldc3d-sv-k3-burman": Config("ldc3d-sv-k3-burman", 3, 4, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-",
                                 stabilisation="burman"),
    "ldc3d-sv-k3-half-burman": Config("ldc3d-sv-k3-half-burman", 3, 2, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-",
                                      stabilisation="burman"),
    "ldc2d-pkp0-tiny-burman": Config("ldc2d-pkp0-tiny-burman", 2, 2, 2, "pkp0", 2, "star", False, re=100.0, stabilisation="burman"),
    "ldc2d-pkp0-tiny": Config("ldc2d-pkp0-tiny", 2, 2, 2, "pkp0", 2, "star", False, re=100.0),
    "ldc3d-sv-k3-tiny": Config("ldc3d-sv-k3-tiny", 3, 1, 1, "sv", 3, "macro", True, re=100.0, sort_order="0+:1-"),
    "ldc3d-pkp0-tiny": Config("ldc3d-pkp0-tiny", 3, 1, 2, "pkp0", 1, "star", False, re=100.0, element="p1fb"),
    # configs[3] at BASELINE size (M = 64: 10.33 M dofs, 274 625 star patches) over FIVE levels: the reference's baseN 16 /
    # nref 2 hierarchy has a 167 k-dof coarse level without macro structure, which needs a sparse direct solver; the same
    # finest mesh over baseN 4 / nref 4 has a 2 967-dof one ("prefer smaller baseN + more levels, state it")
    "ldc3d-pkp0-l5": Config("ldc3d-pkp0-l5", 3, 4, 4, "pkp0", 1, "star", False, re=5000.0, element="p1fb"),
    "ldc3d-pkp0-l5-re100": Config("ldc3d-pkp0-l5-re100", 3, 4, 4, "pkp0", 1, "star", False, re=100.0, element="p1fb"),
    "ldc3d-pkp0-l5-tiny": Config("ldc3d-pkp0-l5-tiny", 3, 1, 4, "pkp0", 1, "star", False, re=100.0, element="p1fb"),
    "ldc3d-pkp0-mid": Config("ldc3d-pkp0-mid", 3, 8, 2, "pkp0", 1, "star", False, re=5000.0, element="p1fb"),
    "ldc3d-pkp0-small": Config("ldc3d-pkp0-small", 3, 4, 2, "pkp0", 1, "star", False, re=1000.0, element="p1fb"),
    "ldc3d-sv-k3-small": Config("ldc3d-sv-k3-small", 3, 2, 1, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    "ldc3d-sv-k3-half": Config("ldc3d-sv-k3-half", 3, 2, 2, "sv", 3, "macro", True, re=5000.0, sort_order="0+:1-"),
    # BASELINE.json configs[2]: backward-facing step, SV k=2, relaxation direction "0+:1-" (bfs2d.py:32),
    # char_length 1 (alfi/problem.py:43).  The reference meshes are Gmsh files (coarse09.msh: 2979 vertices);
    # N = 12 cells per unit length gives a base mesh of that size without them; pass mesh_file to use one.
    "bfs2d-sv-k2": Config("bfs2d-sv-k2", 2, 12, 4, "sv", 2, "macro", True, re=5000.0, length=1.0, domain="bfs",
                          sort_order="0+:1-"),
    "bfs2d-sv-k2-small": Config("bfs2d-sv-k2-small", 2, 4, 2, "sv", 2, "macro", True, re=1000.0, length=1.0,
                                domain="bfs", sort_order="0+:1-"),
    "bfs2d-sv-k2-tiny": Config("bfs2d-sv-k2-tiny", 2, 1, 1, "sv", 2, "macro", True, re=100.0, length=1.0,
                               domain="bfs", sort_order="0+:1-"),
}


def lid_wind(x, extent=None):
    """Lid profile of examples/ldc2d/ldc2d.py:32 / ldc3d/ldc3d.py:26 extended to the interior; on a box
    [0, extent] the profile of the [0, 2]^d cavity is stretched axis by axis."""
    d = x.shape[1]
    if extent is not None:
        x = x * (2.0 / np.asarray(extent, dtype=np.float64))[None, :]
    w = np.zeros_like(x)
    prof = x[:, 0] ** 2 * (2 - x[:, 0]) ** 2 * (0.25 * x[:, 1] ** 2)
    if d == 3:
        prof = prof * x[:, 2] ** 2 * (2 - x[:, 2]) ** 2
    w[:, 0] = prof
    return w


def step_wind(x):
    """Inflow profile of examples/bfs2d/bfs2d.py:20-22 extended along the channel (synthetic wind)."""
    w = np.zeros_like(x)
    y = x[:, 1]
    w[:, 0] = 4.0 * (2.0 - y) * (y - 1.0) * (y > 1.0)
    return w


@dataclass
class LevelData:
    index: int
    level: Level
    V: VectorSpace
    pattern: BlockPattern
    bc_nodes: np.ndarray                 # int32 node indices of the global Dirichlet condition
    A: BSR | None = None                 # level operator
    patches: PatchSet | None = None      # smoother patches (None on level 0)
    P: object | None = None              # scalar CSR prolongation from level index-1
    cell_patches: PatchSet | None = None
    cb_nodes: np.ndarray | None = None   # coarse-boundary nodes of the transfer (T2)
    P_dof_level: bool = False            # P acts on scalar dofs (BubbleTransfer) instead of per node
    A0: BSR | None = None                # nu*visc + gamma*div  (transfer patch operator)
    parts: dict | None = None            # unit-coefficient block values of the visc / div forms (cached)
    D: BSR | None = None                 # gamma * div-div form  (transfer rhs operator)

    @property
    def ndofs(self):
        return self.V.ndofs

    @property
    def bc_dofs(self):
        bs = self.V.bs
        return (self.bc_nodes[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int32)

    @property
    def cb_dofs(self):
        bs = self.V.bs
        return (self.cb_nodes[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int32)


@dataclass
class Problem:
    config: Config
    levels: list[LevelData]
    nu: float
    gamma: float

    @property
    def finest(self):
        return self.levels[-1]


def smoother_patches(cfg: Config, ld: LevelData) -> PatchSet:
    plex = ld.level.plex
    if cfg.patch == "macro":
        H, ents = macro_star_points(plex, cfg.macro_expand)
    else:
        H, ents = star_points(plex)
    order = None
    coords = np.array([plex.point_coords(p) for p in ents])
    if cfg.sort_order:
        order = iteration_order(coords, cfg.sort_order)
    ps = patch_dofs_from_points(plex, ld.V, H, bc_nodes=ld.bc_nodes, order=order)
    ps.centres = coords
    greedy_colouring(ps, ld.V.ndofs)
    if cfg.bary and cfg.stabilisation != "burman":     # the jump terms couple the macro-cell interiors across macro faces
        ps.blocks = macro_interior_blocks(plex, ld.V, ps)
    if cfg.stabilisation == "burman":
        ps.corrections = facet_corrections(ld.V, ps, ld.pattern.facet_cells)
    if cfg.composition == "multiplicative":
        ps.stages = sweep_stages(ps, ld.pattern.rowptr, ld.pattern.colidx)
        ps.symmetrise = True                     # solver.py:324: symmetrise_sweep = multiplicative
    return ps


def _linear_parts(cfg: Config, ld: LevelData):
    if ld.parts is None:
        ld.parts = assemble_parts(ld.V, ld.pattern, None, cfg.discretisation, want=("visc", "div"))
    return ld.parts


def _bsr(ld: LevelData, vals):
    return BSR(ld.V.nnodes, ld.V.bs, ld.pattern.rowptr, ld.pattern.colidx, vals)


def assemble_level(cfg: Config, ld: LevelData, nu: float, gamma: float, advect: float = 1.0, wind=None, stab_wind=None):
    """(Re)assemble the level operator — the once-per-Newton-step hand-over.  The viscous and
    div-div parts are linear in (nu, gamma) and cached; only the advection parts are re-assembled.
    stab_wind: the wind inside Burman's beta if it differs from the linearisation point (the reference updates it
    once per Reynolds number, before the Newton solve: solver.py:270-271)."""
    lin = _linear_parts(cfg, ld)
    vals = nu * lin["visc"] + gamma * lin["div"]
    if advect != 0.0:
        if wind is None:
            ext = ld.V.mesh.extent
            wind = ld.V.interpolate(step_wind if cfg.domain == "bfs" else (lambda xx: lid_wind(xx, ext)))
        adv = assemble_parts(ld.V, ld.pattern, wind, cfg.discretisation, want=("adv1", "adv2"))
        ld.adv1 = adv["adv1"]
        vals += advect * (adv["adv1"] + adv["adv2"])
        if cfg.stabilisation == "burman":
            # F += advect * stabilisation_form (solver.py:233-234); the wind inside beta is a frozen copy of the state
            # (stabilisation.py:19-44), so the term is linear in u and enters residual and Jacobian alike
            _, _, S = burman_facet_tensors(ld.V, wind if stab_wind is None else stab_wind, cfg.stab_weight)
            sv = ld.pattern.scatter_facets(S)
            ld.stab = np.zeros_like(vals)
            for r in range(ld.V.bs):
                ld.stab[:, r, r] = sv
            vals += advect * ld.stab
            if ld.patches is not None and ld.patches.corrections is not None:
                ld.patches.corr_vals = ld.patches.corrections.values(S, advect)
    elif cfg.stabilisation == "burman" and ld.patches is not None and ld.patches.corrections is not None:
        ld.patches.corr_vals = np.zeros(ld.patches.corrections.rows.size)
    ld.A = apply_dirichlet(_bsr(ld, vals), ld.bc_nodes, ld.pattern.rows)
    return ld.A


def assemble_transfer(cfg: Config, ld: LevelData, nu: float, gamma: float):
    """(Re)assemble the Schöberl transfer operators — once per (nu, gamma), transfer.py:238-244."""
    lin = _linear_parts(cfg, ld)
    ld.A0 = _bsr(ld, nu * lin["visc"] + gamma * lin["div"])
    ld.D = _bsr(ld, gamma * lin["div"])
    return ld.A0, ld.D


def build_problem(cfg: Config | str, nu: float | None = None, with_transfer: bool = True,
                  verbose: bool = False, gamma: float | None = None) -> Problem:
    """Generate the hand-over data of a configuration.  ALFIB_PROBLEM_CACHE=<dir> keeps a pickle per
    (config, nu, gamma) there, so that several benchmark scripts of one GPU lease do not each spend the
    ~40 s of host generation cfg5 takes (the cache is never read by tests of the generator itself)."""
    import dataclasses
    import os
    import pickle
    import time
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    if gamma is not None:
        cfg = dataclasses.replace(cfg, gamma=gamma)
    nu = cfg.nu if nu is None else nu
    t0 = time.time()
    cache_dir = os.environ.get("ALFIB_PROBLEM_CACHE")
    cache = None
    if cache_dir:
        cache = os.path.join(cache_dir, "%s_nu%.6e_g%.6e_t%d.pkl" % (cfg.name, nu, cfg.gamma, int(with_transfer)))
        if os.path.exists(cache):
            try:
                with open(cache, "rb") as f:
                    prob = pickle.load(f)
                if prob.config == cfg:
                    if verbose:
                        print("[synth] %s from cache in %.1fs" % (cfg.name, time.time() - t0), file=sys.stderr, flush=True)
                    return prob
            except Exception:       # noqa: BLE001 - a broken cache file is regenerated
                pass
    if cfg.domain == "bfs":
        from .gmsh import read_msh, step_mesh
        base = read_msh(cfg.mesh_file) if cfg.mesh_file else step_mesh(cfg.N)
        hier = build_hierarchy_from(base, cfg.nref, cfg.bary)
    else:
        hier = build_hierarchy(cfg.dim, cfg.N, cfg.nref, cfg.bary, cfg.length, cfg.shape)
    levels = []
    for lev in hier:
        V = VectorSpace(lev.mesh, cfg.k, cfg.element)
        bc = V.tagged_boundary_nodes(cfg.dirichlet_tags) if cfg.domain == "bfs" else V.boundary_nodes()
        burman = cfg.stabilisation == "burman"
        ld = LevelData(lev.index, lev, V, FacetBlockPattern(V) if burman else BlockPattern(V), bc.astype(np.int32))
        if lev.index > 0 and burman:
            ld.patches = smoother_patches(cfg, ld)       # before the assembly: it fills the patch corrections
        assemble_level(cfg, ld, nu, cfg.gamma)
        if lev.index > 0:
            if not burman:
                ld.patches = smoother_patches(cfg, ld)
            if cfg.element == "p1fb":
                # PkP0SchoeberlTransfer.standard_transfer -> BubbleTransfer (alfi/transfer.py:334-356)
                from ..bubble import bubble_transfer_matrix
                ld.P = bubble_transfer_matrix(levels[-1].V, V, hier[lev.index - 1].c2f)
                ld.P_dof_level = True
            else:
                ld.P = prolongation_matrix(levels[-1].V, V, hier[lev.index - 1].c2f)
            if with_transfer:
                ld.cell_patches, ld.cb_nodes = cell_patch_set(hier, lev.index, V, cfg.bary)
                if cfg.bary and not burman:
                    ld.cell_patches.blocks = macro_interior_blocks(lev.plex, V, ld.cell_patches)
                assemble_transfer(cfg, ld, nu, cfg.gamma)
        levels.append(ld)
        if verbose:
            print("[synth] level %d: %d dofs, %d patches, %.1fs" % (
                lev.index, V.ndofs, 0 if ld.patches is None else ld.patches.npatch, time.time() - t0),
                file=sys.stderr, flush=True)
    prob = Problem(cfg, levels, nu, cfg.gamma)
    if cache:
        try:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = cache + ".%d.tmp" % os.getpid()
            with open(tmp, "wb") as f:
                pickle.dump(prob, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, cache)
        except Exception:           # noqa: BLE001 - caching is best effort
            pass
    return prob
