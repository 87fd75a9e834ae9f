"""Box-shaped Kuhn hierarchies (Config.shape): the weak-scaling family of BASELINE configs[4] — rank r of an
sx x sy x sz rank grid gets one [0, 2]^3-sized brick of the mesh.  Counts follow SURVEY Appendix B brick by brick,
the oracle's cycle still contracts, and the cube (shape = ()) is unchanged."""
import dataclasses

import numpy as np
import pytest

from alfi_b200.synth.mesh import alfeld_split, kuhn_mesh, locate_in_kuhn
from alfi_b200.synth.problem import CONFIGS, build_problem
from oracle import hotpath as hp


@pytest.mark.parametrize("dim,shape", [(2, (2, 1)), (2, (1, 3)), (3, (2, 1, 1)), (3, (1, 2, 2)), (3, (2, 2, 2))])
def test_box_kuhn_mesh_counts_and_location(dim, shape):
    M = 2
    m = kuhn_mesh(dim, M, 2.0, shape)
    Ma = [M * s for s in shape]
    assert m.nv == int(np.prod([a + 1 for a in Ma]))
    assert m.nc == (2 if dim == 2 else 6) * int(np.prod(Ma))
    assert np.allclose(m.coords.max(axis=0), 2.0 * np.asarray(shape))
    # every cell is found again from its centroid, volumes are those of the cube mesh
    cent = m.coords[m.cells].mean(axis=1)
    assert np.array_equal(locate_in_kuhn(m, cent), np.arange(m.nc))
    e = m.coords[m.cells[:, 1:]] - m.coords[m.cells[:, :1]]
    vol = np.abs(np.linalg.det(e)) / (2 if dim == 2 else 6)
    assert np.allclose(vol, (2.0 / M) ** dim / (2 if dim == 2 else 6))
    # boundary vertices = those on the faces of the box
    nb = m.nv - int(np.prod([a - 1 for a in Ma]))
    assert int(m.boundary_vertex_mask().sum()) == nb
    if all(s == shape[0] for s in shape):
        cube = kuhn_mesh(dim, M * shape[0], 2.0 * shape[0])
        assert np.array_equal(cube.cells, m.cells) and np.allclose(cube.coords, m.coords)
    a = alfeld_split(m)
    assert a.nc == (dim + 1) * m.nc and a.shape == m.shape


@pytest.mark.parametrize("name,shape", [("ldc2d-sv-k2-tiny", (2, 1)), ("ldc3d-sv-k3-tiny", (2, 1, 1)), ("ldc2d-pkp0-tiny", (1, 2))])
def test_box_problem_patches_and_cycle(name, shape):
    cfg = dataclasses.replace(CONFIGS[name], shape=shape)
    prob = build_problem(cfg, gamma=10.0, nu=0.2)
    cube = build_problem(CONFIGS[name], gamma=10.0, nu=0.2)
    fine, cfine = prob.finest, cube.finest
    # one patch per non-ghost (macro) vertex: the vertex count of the box; interior patches have the cube's size
    Ma = [cfg.N * 2 ** cfg.nref * s for s in shape]
    assert fine.patches.npatch == int(np.prod([a + 1 for a in Ma]))
    assert fine.patches.sizes.max() == cfine.patches.sizes.max()
    nbricks = int(np.prod(shape))
    assert cfine.ndofs < fine.ndofs < nbricks * cfine.ndofs          # bricks share their interface dofs
    lv = [hp.level_from_host(l) for l in prob.levels]
    b = np.random.default_rng(1).standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    x = hp.fcycle(lv, b, cfg.m)
    r = b - lv[-1].A @ x
    r[lv[-1].bc_dofs] = 0
    assert np.linalg.norm(r) < 0.5 * np.linalg.norm(b)
    # the transfers keep their defining properties on the box (restrict = prolong^T)
    L = len(lv) - 1
    c = np.random.default_rng(2).standard_normal(lv[L - 1].n)
    c[lv[L - 1].bc_dofs] = 0
    f = np.random.default_rng(3).standard_normal(lv[L].n)
    f[lv[L].bc_dofs] = 0
    assert abs(f @ hp.prolong(lv[L], c) - hp.restrict(lv[L], f, lv[L - 1].bc_dofs) @ c) <= 1e-10 * np.linalg.norm(f) * np.linalg.norm(c)


def test_brick_grid_of_a_cube_is_the_cube_problem():
    """A 2 x 2 x 2 grid of unit bricks is the [0, 2]^3 cube (config ldc3d-sv-k3-s8 == cfg5): same mesh, numbering and
    operator when the viscosity is the same."""
    cube = dataclasses.replace(CONFIGS["ldc3d-sv-k3-tiny"], N=2, nref=1)
    grid = dataclasses.replace(cube, N=1, length=1.0, shape=(2, 2, 2), re=cube.re / 2)
    assert abs(cube.nu - grid.nu) < 1e-15
    a, b = build_problem(cube), build_problem(grid)
    for la, lb in zip(a.levels, b.levels):
        assert np.array_equal(la.level.mesh.cells, lb.level.mesh.cells)
        assert np.array_equal(la.A.colidx, lb.A.colidx) and np.allclose(la.A.vals, lb.A.vals, rtol=1e-13, atol=1e-13)
        if la.patches is not None:
            assert np.array_equal(la.patches.dofs, lb.patches.dofs) and np.array_equal(la.patches.order, lb.patches.order)
    s8, c5 = CONFIGS["ldc3d-sv-k3-s8"], CONFIGS["ldc3d-sv-k3"]
    assert abs(s8.nu - c5.nu) < 1e-18 and s8.N * 2 == c5.N and s8.nref == c5.nref and s8.length * 2 == c5.length
