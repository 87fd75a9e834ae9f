"""Generate the golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot run here, so these pin OUR
oracle (oracle/hotpath.py) and the host-side index-set builders against regressions: patch dof
sets, iteration order and colouring (integers, exact), and the oracle's smoother / transfer /
FGMRES / F-cycle outputs for seeded inputs (float64, mild parameters gamma=10, nu=0.2 so the
numbers are insensitive to LAPACK version).  "Parity unpinned" by the reference itself.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from alfi_b200.synth.problem import build_problem  # noqa: E402
from oracle import hotpath as hp  # noqa: E402

NAMES = ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny", "bfs2d-sv-k2-tiny"]


def make(name):
    prob = build_problem(name, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    out = {}
    rng = np.random.default_rng(20261017)
    for l, (ld, L) in enumerate(zip(prob.levels, lv)):
        if ld.patches is None:
            continue
        ps = ld.patches
        out["l%d_offsets" % l] = ps.offsets
        out["l%d_dofs" % l] = ps.dofs
        out["l%d_order" % l] = ps.order
        out["l%d_colours" % l] = ps.colours
        out["l%d_cell_offsets" % l] = ld.cell_patches.offsets
        out["l%d_cell_dofs" % l] = ld.cell_patches.dofs
        out["l%d_cb_dofs" % l] = ld.cb_dofs
        x = rng.standard_normal(L.n)
        x[L.bc_dofs] = 0
        c = rng.standard_normal(lv[l - 1].n)
        c[lv[l - 1].bc_dofs] = 0
        out["l%d_x" % l] = x
        out["l%d_c" % l] = c
        out["l%d_spmv" % l] = L.A @ x
        out["l%d_apply" % l] = hp.smoother_apply(x, L.offsets, L.dofs, L.order, L.factors, L.bc_dofs)
        out["l%d_prolong" % l] = hp.prolong(L, c)
        out["l%d_restrict" % l] = hp.restrict(L, x, lv[l - 1].bc_dofs)
        out["l%d_smooth" % l] = hp.smooth(L, x, np.zeros(L.n), prob.config.m)
    b = rng.standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    out["b"] = b
    out["fcycle"] = hp.fcycle(lv, b, prob.config.m)
    return out


if __name__ == "__main__":
    for name in (sys.argv[1:] or NAMES):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(path, **make(name))
        print(path, os.path.getsize(path))
