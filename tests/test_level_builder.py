"""alfi_b200.level_builder: the coarse-grained hand-over assembled level by level through the `access` protocol that
`FiredrakeAdapter.levels()` / `firedrake_adapter.transfer_backend()` use for a live solver — here over a stand-in that
answers from the synthetic problems, so that the result can be compared with the hand-over the GPU tests run on
(`level_input_from_synth`).  The standard prolongation is recovered by coloured probing of the stand-in's ``prolong``."""
import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.level_builder import build_level_inputs, colour_candidates, probe_prolongation, prolongation_candidates
from alfi_b200.multigrid import level_input_from_synth


class SynthAccess:
    """The `access` protocol answered by a synthetic Problem (every method is what one Firedrake call returns)."""

    def __init__(self, prob):
        self.prob = prob
        self.dof_level_transfer = any(getattr(ld, "P_dof_level", False) for ld in prob.levels)
        self.prolong_calls = 0

    def nlevels(self):
        return len(self.prob.levels)

    def space(self, l):
        return self.prob.levels[l].V

    def plex(self, l):
        return self.prob.levels[l].level.plex

    def coarse_to_fine_cells(self, l):
        return self.prob.levels[l].level.c2f

    def bc_nodes(self, l):
        return self.prob.levels[l].bc_nodes

    def operator_blocks(self, l):
        A = self.prob.levels[l].A
        return A.rowptr, A.colidx, A.vals

    def transfer_blocks(self, l, nu, gamma):
        assert (nu, gamma) == (self.prob.nu, self.prob.gamma)
        ld = self.prob.levels[l]
        return ld.A0.vals, ld.D.vals

    def prolong(self, l, coarse):
        self.prolong_calls += 1
        ld = self.prob.levels[l]
        return ld.P @ coarse                      # firedrake.prolong / BubbleTransfer.prolong of a nodal / dof array

    def parameters(self):
        return self.prob.nu, self.prob.gamma


def _same(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, what
    else:
        assert np.array_equal(np.asarray(a), np.asarray(b)), what


CASES = [("ldc2d-sv-k2-tiny", dict(construct="alfi.MacroStar", bary=True)),
         ("ldc3d-sv-k3-tiny", dict(construct="alfi.MacroStar", bary=True)),
         ("ldc2d-pkp0-tiny", dict(construct="star", bary=False)),
         ("ldc3d-pkp0-tiny", dict(construct="star", bary=False)),
         ("bfs2d-sv-k2-tiny", dict(construct="alfi.MacroStar", bary=True))]


@pytest.mark.parametrize("name,kw", CASES)
def test_levels_built_through_the_access_protocol_equal_the_synthetic_hand_over(problems, name, kw):
    prob = problems(name)
    cfg = prob.config
    acc = SynthAccess(prob)
    cache = {}
    levels = build_level_inputs(acc, sort_order=cfg.sort_order, macro_expand=cfg.macro_expand, prolongations=cache, **kw)
    want = [level_input_from_synth(ld) for ld in prob.levels]
    assert len(levels) == len(want)
    for l, (a, b) in enumerate(zip(levels, want)):
        for f in ("n_nodes", "bs", "P_dof_level", "symmetrise_sweep"):
            assert getattr(a, f) == getattr(b, f), (l, f)
        for f in ("rowptr", "colidx", "vals", "bc_dofs", "patch_offsets", "patch_dofs", "patch_order", "patch_colours",
                  "patch_blocks", "cell_offsets", "cell_dofs", "cell_blocks", "cb_dofs", "a0_vals", "d_vals",
                  "coarse_dofs", "coarse_blocks"):
            _same(getattr(a, f), getattr(b, f), (name, l, f))
        if l > 0:
            P, Q = a.P.tocsr(), b.P.tocsr()
            assert P.shape == Q.shape
            assert abs(P - Q).max() <= 1e-14 * abs(Q).max(), (name, l)
            assert P.nnz <= Q.nnz                       # probing found no entry the assembled matrix lacks
    calls = acc.prolong_calls
    build_level_inputs(acc, sort_order=cfg.sort_order, macro_expand=cfg.macro_expand, prolongations=cache, **kw)
    assert acc.prolong_calls == calls                    # cached per mesh


def test_probing_needs_a_bounded_number_of_prolong_applications(problems):
    """One application of the framework's prolong per colour (+ 1 check) and level; the number of colours is bounded by
    how many coarse nodes can influence the fine nodes around one coarse node, not by the mesh size."""
    prob = problems("ldc2d-sv-k2")
    acc = SynthAccess(prob)
    levels = build_level_inputs(acc, construct="alfi.MacroStar", bary=True, smoother=False, transfer=False)
    ncoarse = sum(li.P.shape[1] for li in levels[1:])
    assert acc.prolong_calls < ncoarse / 10, (acc.prolong_calls, ncoarse)
    for li, ld in zip(levels[1:], prob.levels[1:]):
        assert abs(li.P - ld.P).max() <= 1e-14


def test_transfer_only_hand_over_has_no_smoother_data(problems):
    prob = problems("ldc2d-sv-k2-tiny")
    levels = build_level_inputs(SynthAccess(prob), construct="alfi.MacroStar", bary=True, smoother=False)
    for l, li in enumerate(levels):
        assert li.vals is None and li.patch_offsets is None
        if l > 0:
            assert li.P is not None and li.cell_offsets is not None and li.a0_vals is not None and li.cb_dofs is not None


def test_colouring_separates_the_candidates_of_every_fine_node():
    rng = np.random.default_rng(3)
    nf, nc = 60, 25
    C = sp.random(nf, nc, density=0.15, random_state=4, format="csr")
    C.data[:] = 1
    colour, ncol = colour_candidates(C)
    for f in range(nf):
        cols = C.indices[C.indptr[f]:C.indptr[f + 1]]
        assert np.unique(colour[cols]).size == cols.size
    assert ncol <= nc
    # probing a matrix with exactly that pattern recovers it with ncol + 1 applications
    M = C.copy().astype(float)
    M.data[:] = rng.standard_normal(M.nnz)
    calls = []
    P = probe_prolongation(lambda x: (calls.append(1), M @ x)[1], C)
    assert abs(P - M).max() <= 1e-15 and len(calls) == ncol + 1
    # an operator with an entry outside the candidate pattern is refused, not silently truncated
    f0 = int(np.argmin(np.diff(C.indptr)))
    c0 = int(np.setdiff1d(np.arange(nc), C.indices[C.indptr[f0]:C.indptr[f0 + 1]])[0])
    M2 = (M + sp.csr_matrix(([1.0], ([f0], [c0])), shape=M.shape)).tocsr()
    with pytest.raises(ValueError):
        probe_prolongation(lambda x: M2 @ x, C)


def test_candidates_are_the_parent_cells_nodes():
    fine = np.array([[0, 1], [1, 2], [2, 3], [3, 4]])          # four fine cells of a 1-D mesh, two nodes each
    coarse = np.array([[0, 1], [1, 2]])
    c2f = np.array([[0, 1], [2, 3]])
    C = prolongation_candidates(fine, coarse, c2f, 5, 3).toarray()
    assert np.array_equal(C, [[1, 1, 0], [1, 1, 0], [1, 1, 1], [0, 1, 1], [0, 1, 1]])


# ---- the adapter pieces that sit on the builder (no Firedrake: stand-ins for the solver, the backend, Functions) -------------
def _reference_fieldsplit0(key):
    import copy
    import json
    import os
    params = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_parameters.json")))
    return copy.deepcopy(params[key]["outer"]["fieldsplit_0"])


def test_firedrake_adapter_levels_from_the_reference_dictionary(problems):
    """`FiredrakeAdapter.levels()` = what `alfi_b200.VelocityMGPC.initialize` asks for, driven by the reference's own
    fieldsplit_0 dictionary (smoothing, MacroStar, sort order) and an access object."""
    from alfi_b200.firedrake_adapter import FiredrakeAdapter
    prob = problems("ldc3d-sv-k3-tiny-literal")
    ad = FiredrakeAdapter(access=SynthAccess(prob), fieldsplit_0=_reference_fieldsplit0("ldc3d-sv-k3"), hierarchy="bary")
    assert ad.smoothing == prob.config.m == 10 and ad.parameters(None) == (prob.nu, prob.gamma)
    levels = ad.levels(None)
    want = [level_input_from_synth(ld) for ld in prob.levels]
    for a, b in zip(levels, want):
        for f in ("patch_offsets", "patch_dofs", "patch_order", "patch_colours", "patch_blocks", "cell_dofs", "cb_dofs", "vals"):
            _same(getattr(a, f), getattr(b, f), f)
    ad.access.pressure_operators = lambda: ("B", "Minv")          # the two assembled matrices come from the access object
    B, Minv, bc = ad.pressure_operators(None)
    assert (B, Minv) == ("B", "Minv") and np.array_equal(bc, prob.finest.bc_dofs)
    with pytest.raises(RuntimeError):
        FiredrakeAdapter().levels(None)


def test_transfer_backend_hands_over_the_transfer_levels_and_the_value_callback(problems):
    from alfi_b200.firedrake_adapter import transfer_backend
    prob = problems("ldc2d-sv-k2-tiny")
    acc = SynthAccess(prob)
    seen = {}

    def make_backend(levels, values_for=None, device=0):
        seen["levels"], seen["device"] = levels, device
        return {"backend": "ctx", "values_for": values_for}
    solver = type("Solver", (), {"hierarchy": "bary"})()
    kw = transfer_backend(solver, access=acc, make_backend=make_backend, device=3)
    assert kw["backend"] == "ctx" and seen["device"] == 3
    for l, (li, ld) in enumerate(zip(seen["levels"], prob.levels)):
        assert li.vals is None and li.patch_offsets is None            # nothing of the smoother
        if l > 0:
            assert abs(li.P - ld.P).max() <= 1e-14
            _same(li.cell_dofs, ld.cell_patches.dofs, "cell dofs")
            _same(li.cb_dofs, ld.cb_dofs, "cb dofs")
    a0, d = kw["values_for"](1, prob.nu, prob.gamma)
    assert a0 is prob.levels[1].A0.vals and d is prob.levels[1].D.vals


def test_transfer_classes_accept_firedrake_functions():
    """The TransferManager passes Functions (solver.py:593-596): read through .dat.data_ro, written through .dat.data."""
    from alfi_b200.transfer import SVSchoeberlTransfer

    class Dat:
        def __init__(self, a):
            self.data = a
            self.data_ro = a

    class Function:
        def __init__(self, n, bs):
            self.dat = Dat(np.zeros((n, bs)))

    class Backend:
        def level_sizes(self):
            return {0: 6, 1: 12}

        def transfer_update(self, level, a0, d):
            self.updated = level

        def prolong(self, level, coarse, fine):
            assert coarse.shape == (6,) and fine.shape == (12,) and level == 1
            fine[:] = np.repeat(coarse, 2)

        def restrict(self, level, fine, coarse):
            assert fine.shape == (12,) and coarse.shape == (6,) and level == 1
            coarse[:] = fine.reshape(6, 2).sum(axis=1)
    be = Backend()
    t = SVSchoeberlTransfer((0.1, 10.0), 2, "bary", backend=be, values_for=lambda l, nu, g: (None, None))
    c, f = Function(3, 2), Function(6, 2)
    c.dat.data[...] = np.arange(6.0).reshape(3, 2)
    t.prolong(c, f)
    assert np.array_equal(f.dat.data.reshape(-1), np.repeat(np.arange(6.0), 2)) and be.updated == 1
    t.restrict(f, c)
    assert np.array_equal(c.dat.data.reshape(-1), 2 * np.arange(6.0))
    # plain arrays keep working
    fa = np.zeros(12)
    t.prolong(np.arange(6.0), fa)
    assert np.array_equal(fa, np.repeat(np.arange(6.0), 2))


def test_cell_tables_are_converted_to_plex_numbering():
    from alfi_b200.firedrake_adapter import c2f_to_plex
    c2f = np.array([[0, 1], [2, 3], [4, 5]])                  # firedrake numbers, row = coarse firedrake cell
    f2p_c = np.array([2, 0, 1])                               # coarse firedrake cell i is plex cell f2p_c[i]
    f2p_f = np.array([5, 4, 3, 2, 1, 0])
    out = c2f_to_plex(c2f, f2p_c, f2p_f)
    assert np.array_equal(out, [[3, 2], [1, 0], [5, 4]])
