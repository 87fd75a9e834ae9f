"""Index-set fixtures produced by the REFERENCE'S OWN CODE (run from the repo root, in the build container
where /root/reference exists: python tests/golden/make_reference_golden.py).

oracle/refshim.py loads alfi/relaxation.py and alfi/transfer.py from the reference tree over stand-ins for
Firedrake / petsc4py and runs, on our synthetic DMPlex look-alikes,

    Star()(pc), MacroStar()(pc)                        alfi/relaxation.py:110-177   (rows P1-P3)
    CoarseCellPatches()(pc), CoarseCellMacroPatches()(pc)   alfi/transfer.py:13-88  (row T1)
    AutoSchoeberlTransfer.fix_coarse_boundaries(V)     alfi/transfer.py:121-158     (row T2)

The results — patch point lists exactly as the reference builds them (order and duplicates included),
iteration sets, coarse-boundary node lists — go to tests/golden/reference_index_sets.npz; the tests compare
alfi_b200's builders with them wherever the suite runs (the GPU box has no reference tree).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from alfi_b200.synth.fem import VectorSpace  # noqa: E402
from alfi_b200.synth.gmsh import step_mesh  # noqa: E402
from alfi_b200.synth.hierarchy import build_hierarchy, build_hierarchy_from  # noqa: E402
from oracle import refshim  # noqa: E402

# name -> (hierarchy builder, polynomial degree, element kind, sort order)
CASES = {
    "kuhn2d-bary": (lambda: build_hierarchy(2, 2, 1, True), 2, "lagrange", "0+:1-"),
    "kuhn2d-plain": (lambda: build_hierarchy(2, 2, 2, False), 2, "lagrange", None),
    "kuhn3d-bary": (lambda: build_hierarchy(3, 1, 1, True), 3, "lagrange", "0+:1-:2+|2-"),
    "kuhn3d-p1fb": (lambda: build_hierarchy(3, 1, 1, False), 1, "p1fb", None),
    "step-bary": (lambda: build_hierarchy_from(step_mesh(1, seed=3), 1, True), 2, "lagrange", "0+:1-"),
}


def flatten(sets):
    off = np.concatenate(([0], np.cumsum([len(s) for s in sets]))).astype(np.int64)
    data = np.concatenate([np.asarray(s, dtype=np.int64) for s in sets]) if sets else np.empty(0, np.int64)
    return off, data


def run_reference(name):
    """Everything the reference's code produces for one case: dict of arrays."""
    build, k, kind, sort = CASES[name]
    levels = build()
    bary = levels[0].bary
    out = {}
    with refshim.reference_modules() as (rel, tr):
        hier = refshim.FakeHierarchy(levels)
        for l, lev in enumerate(levels):
            dm = refshim.coord_plex(lev.plex)
            refshim.set_options({})
            patches, iset = rel.Star()(refshim.FakePC(dm))
            out["l%d_star_off" % l], out["l%d_star_pts" % l] = flatten([p.getIndices() for p in patches])
            out["l%d_star_iter" % l] = iset.getIndices().astype(np.int64)
            if bary:
                table = {} if sort is None else {"pc_patch_construction_MacroStar_sort_order": sort}
                refshim.set_options(table)
                ms = rel.MacroStar()
                patches, iset = ms(refshim.FakePC(dm))
                out["l%d_macro_off" % l], out["l%d_macro_pts" % l] = flatten([p.getIndices() for p in patches])
                out["l%d_macro_iter" % l] = iset.getIndices().astype(np.int64)
            if l > 0:
                ctx = types.SimpleNamespace(_x=types.SimpleNamespace(ufl_domain=lambda m=hier[l]: m))
                maker = tr.CoarseCellMacroPatches() if bary else tr.CoarseCellPatches()
                patches, iset = maker(refshim.FakePC(dm, ctx=ctx))
                out["l%d_cell_off" % l], out["l%d_cell_pts" % l] = flatten([p.getIndices() for p in patches])
                out["l%d_cell_iter" % l] = iset.getIndices().astype(np.int64)
                V = VectorSpace(lev.mesh, k, kind)
                bc = tr.AutoSchoeberlTransfer.fix_coarse_boundaries(refshim.FakeFunctionSpace(hier, l, V))
                out["l%d_cb_nodes" % l] = np.asarray(bc.nodes, dtype=np.int64)
    refshim.set_options({})
    return out


if __name__ == "__main__":
    blob = {}
    for name in CASES:
        for key, val in run_reference(name).items():
            blob[name + "/" + key] = val
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_index_sets.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), len(blob), "arrays")
