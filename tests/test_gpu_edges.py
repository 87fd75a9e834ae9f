"""Edge cases of the C-ABI on the GPU: layouts, ragged/empty patches, repeated iteration sets,
singular patches, option handling ("cover the edge cases the domain has")."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def random_bsr(n_nodes, bs, seed, density=0.2):
    rng = np.random.default_rng(seed)
    pat = sp.random(n_nodes, n_nodes, density=density, random_state=seed, format="csr")
    pat = (pat + sp.identity(n_nodes)).tocsr()
    pat.sort_indices()
    rowptr, colidx = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
    vals = rng.standard_normal((colidx.size, bs, bs))
    rows = np.repeat(np.arange(n_nodes), np.diff(rowptr))
    vals[rows == colidx] += 4.0 * np.sqrt(n_nodes) * np.eye(bs)          # comfortably non-singular
    A = sp.bsr_matrix((vals, colidx, rowptr), shape=(n_nodes * bs,) * 2).tocsr()
    return rowptr, colidx, vals, A


@pytest.mark.parametrize("bs", [2, 3])
def test_block_layouts_and_spmv(bs):
    from alfi_b200.lib import Context
    n_nodes = 57
    rowptr, colidx, vals, A = random_bsr(n_nodes, bs, 1)
    x = np.random.default_rng(2).standard_normal(n_nodes * bs)
    ctx = Context()
    ctx.level_create(0, n_nodes, bs)
    ctx.set_bsr_pattern(0, rowptr, colidx)
    ctx.set_bsr_values(0, vals, block_col_major=False)
    y0 = ctx.spmv(0, x, np.empty_like(x))
    ctx.set_bsr_values(0, np.ascontiguousarray(vals.transpose(0, 2, 1)), block_col_major=True)   # PETSc BAIJ layout
    y1 = ctx.spmv(0, x, np.empty_like(x))
    assert rel(y0, A @ x) < 1e-14 and np.array_equal(y0, y1)
    ctx.close()


@pytest.mark.parametrize("bs", [2, 3])
def test_ragged_empty_and_repeated_patches(bs):
    """Patch sizes 0, 1, odd, even, > 64 (several tiles); a patch twice in the iteration set."""
    from alfi_b200.lib import Context
    n_nodes = 120
    n = n_nodes * bs
    rowptr, colidx, vals, A = random_bsr(n_nodes, bs, 3, density=0.1)
    rng = np.random.default_rng(4)
    sizes = [0, 1, 2, 3, 17, 64, 65, 131, 0, 200]
    patches = [rng.choice(n, size=s, replace=False).astype(np.int32) for s in sizes]
    offsets = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    dofs = np.concatenate(patches) if patches else np.empty(0, np.int32)
    bc = np.array([0, 5, n - 1], dtype=np.int32)
    x = rng.standard_normal(n)

    def want(order):
        y = np.zeros(n)
        for p in order:
            I = patches[p]
            if I.size:
                y[I] += np.linalg.solve(A[I][:, I].toarray(), x[I])
        y[bc] = x[bc]
        return y

    for order in (np.arange(len(sizes)), np.array([9, 3, 3, 7, 5, 9, 1])):
        for det in (False, True):
            ctx = Context(deterministic=det)
            ctx.level_create(0, n_nodes, bs)
            ctx.set_bsr_pattern(0, rowptr, colidx)
            ctx.set_bsr_values(0, vals)
            ctx.set_bc(0, bc)
            ctx.set_patches(0, offsets, dofs, order.astype(np.int32), None)
            ctx.factor(0)
            y = ctx.smoother_apply(0, x, np.empty(n))
            assert rel(y, want(order)) < 1e-11, (order, det)
            for p, s in enumerate(sizes):
                if s:
                    I = patches[p]
                    inv = ctx.patch_inverse(0, p, s)
                    assert np.abs(inv @ A[I][:, I].toarray() - np.eye(s)).max() < 1e-10
            ctx.close()


@pytest.mark.parametrize("bs", [2, 3])
def test_condensed_edge_cases(bs):
    """Condensed (block/separator) patch sets on a synthetic clustered operator: empty, separator-only
    and block-only patches, blocks without separator neighbours, blocks and neighbourhoods at the
    64-dof limit, > 64 separator dofs (several tiles), overlapping patches, a repeated visit."""
    from alfi_b200.lib import Context
    from tests.condense_cases import clustered_problem, dense_reference
    case = clustered_problem(bs, seed=bs)
    n = case["n_nodes"] * bs
    bc = np.array([1, n - 2], dtype=np.int32)
    x = np.random.default_rng(11).standard_normal(n)
    npatch = len(case["patches"])
    for order in (np.arange(npatch), np.array([7, 3, 3, 6, 5, 4, 1, 2])):
        for det in (False, True):
            ctx = Context(deterministic=det)
            ctx.level_create(0, case["n_nodes"], bs)
            ctx.set_bsr_pattern(0, case["rowptr"], case["colidx"])
            ctx.set_bsr_values(0, case["vals"])
            ctx.set_bc(0, bc)
            ctx.set_patches(0, case["offsets"], case["dofs"], order.astype(np.int32), None)
            dense_bytes = ctx.patch_storage_bytes(0)
            ctx.set_patch_blocks(0, case["blocks"])
            assert ctx.patch_storage_bytes(0) < dense_bytes
            ctx.factor(0)
            y = ctx.smoother_apply(0, x, np.empty(n))
            assert rel(y, dense_reference(case, order, x, bc)) < 1e-12, (order, det)
            if det and len(set(order.tolist())) == order.size:
                # bitwise reproducible; a repeated visit falls back to atomics like the dense path
                assert np.array_equal(y, ctx.smoother_apply(0, x, np.empty(n)))
            for p, I in enumerate(case["patches"]):
                if I.size:
                    inv = ctx.patch_inverse(0, p, I.size)
                    assert np.abs(inv @ case["A"][I][:, I].toarray() - np.eye(I.size)).max() < 1e-10
            # new values, same structure: the per-Newton-step path
            ctx.set_bsr_values(0, 2.0 * case["vals"])
            ctx.factor(0)
            y2 = ctx.smoother_apply(0, x, np.empty(n))
            want = 0.5 * dense_reference(case, order, x)
            want[bc] = x[bc]
            assert rel(y2, want) < 1e-12
            # back to dense inverses on the same context
            ctx.set_patch_blocks(0, None)
            ctx.factor(0)
            assert rel(ctx.smoother_apply(0, x, np.empty(n)), want) < 1e-12
            ctx.close()


def test_singular_patch_is_reported():
    from alfi_b200.lib import AlfibError, Context
    n_nodes, bs = 10, 2
    rowptr = np.arange(n_nodes + 1, dtype=np.int32)
    colidx = np.arange(n_nodes, dtype=np.int32)
    vals = np.tile(np.eye(bs), (n_nodes, 1, 1))
    vals[3] = 0.0                                           # node 3: zero block -> singular patch
    ctx = Context()
    ctx.level_create(0, n_nodes, bs)
    ctx.set_bsr_pattern(0, rowptr, colidx)
    ctx.set_bsr_values(0, vals)
    ctx.set_patches(0, np.array([0, 4, 8], np.int64), np.array([0, 1, 2, 3, 4, 5, 6, 7], np.int32), None, None)
    with pytest.raises(AlfibError, match="singular"):
        ctx.factor(0)
    ctx.close()


def test_bad_arguments_are_rejected():
    from alfi_b200.lib import AlfibError, Context
    ctx = Context()
    ctx.level_create(0, 4, 2)
    with pytest.raises(AlfibError):
        ctx.level_create(0, 4, 2)                            # exists
    with pytest.raises(AlfibError):
        ctx.set_bsr_pattern(0, np.array([0, 1, 2, 3, 4], np.int32), np.array([0, 1, 2, 9], np.int32))   # column 9
    with pytest.raises(AlfibError):
        ctx.set_patches(0, np.array([0, 2], np.int64), np.array([1, 1], np.int32), None, None)          # duplicate dof
    with pytest.raises(AlfibError):
        ctx.set_patches(0, np.array([0, 1], np.int64), np.array([99], np.int32), None, None)            # out of range
    with pytest.raises(AlfibError):
        ctx.smoother_apply(0, np.zeros(8), np.zeros(8))      # not factored
    with pytest.raises(AlfibError):
        ctx.cycle_setup(2, 6)                                # level 1 does not exist
    ctx.close()


def test_graph_and_eager_cycles_agree(problems):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems("ldc2d-pkp0-tiny", gamma=10.0, nu=0.2)
    b = np.random.default_rng(5).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    out = []
    for graph in (0, 1):
        mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=True)
        mg.ctx.set_option(5, graph)
        xs = [mg.apply(b, np.empty_like(b)).copy() for _ in range(4)]       # eager, capture, replay, replay
        assert all(np.array_equal(xs[0], x) for x in xs[1:])
        out.append(xs[-1])
        mg.ctx.close()
    assert np.array_equal(out[0], out[1])


def test_replayed_cycle_refuses_stale_inverses_and_follows_new_values(problems):
    """a captured cycle must not run on inverses that are older than the operator values, and a Newton step's
    refresh (same buffers, new values) must show in the replayed result exactly as in an eager one."""
    from alfi_b200.lib import AlfibError
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    b = np.random.default_rng(9).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    mg = DeviceMultigrid(levels, prob.config.m, deterministic=True)
    x0 = [mg.apply(b, np.empty_like(b)).copy() for _ in range(3)][-1]       # eager, capture, replay
    mg.ctx.set_bsr_values(1, levels[1].vals)                                # values changed, no re-factorisation
    with pytest.raises(AlfibError, match="alfib_level_factor"):
        mg.apply(b, np.empty_like(b))
    mg.ctx.factor(1)
    mg.ctx.set_bsr_values(0, levels[0].vals)
    with pytest.raises(AlfibError, match="alfib_coarse_factor"):
        mg.apply(b, np.empty_like(b))
    mg.ctx.coarse_factor()
    assert np.array_equal(mg.apply(b, np.empty_like(b)), x0)
    # scaled operator on every level: the cycle is linear in the inverse, so x scales by 1 / 2 — through the replay
    import dataclasses
    scaled = [dataclasses.replace(li, vals=2.0 * li.vals, a0_vals=None if li.a0_vals is None else 2.0 * li.a0_vals,
                                  d_vals=None if li.d_vals is None else 2.0 * li.d_vals) for li in levels]
    mg.update_operators(scaled)
    mg.update_transfers(scaled)
    x2 = mg.apply(b, np.empty_like(b))
    assert np.linalg.norm(2.0 * x2 - x0) <= 1e-12 * np.linalg.norm(x0)
    mg.ctx.close()


def test_caller_colours_are_validated():
    """two patches of one colour that share a dof would race in the deterministic scatter."""
    from alfi_b200.lib import AlfibError, Context
    ctx = Context()
    ctx.level_create(0, 4, 2)
    off, dofs = np.array([0, 3, 6], np.int64), np.array([0, 1, 2, 2, 3, 4], np.int32)       # dof 2 in both
    with pytest.raises(AlfibError, match="same colour"):
        ctx.set_patches(0, off, dofs, None, np.array([0, 0], np.int32))
    ctx.set_patches(0, off, dofs, None, np.array([0, 1], np.int32))
    assert ctx.colours(0, 2).tolist() == [0, 1]
    ctx.close()


def test_five_level_cycle_equals_oracle(problems):
    """configs[3] at BASELINE size runs over five levels (ldc3d-pkp0-l5: baseN 4, nref 4); the same hierarchy depth on a
    16^3 mesh here: F-cycle (PCMG full: 1 + 2 + 3 + 4 + 5 level visits) against the oracle."""
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    from oracle import hotpath as hp
    prob = problems("ldc3d-pkp0-l5-tiny", gamma=10.0, nu=0.2)
    assert len(prob.levels) == 5 and prob.finest.ndofs == 3 * (17 ** 3 + 12 * 16 ** 3 + 6 * 16 ** 2)
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=True)
    olv = [hp.level_from_host(l) for l in prob.levels]
    b = np.random.default_rng(12).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    x = mg.apply(b, np.empty_like(b))
    want = hp.fcycle(olv, b, prob.config.m)
    assert rel(x, want) <= 1e-11
    for _ in range(2):
        x = mg.apply(b, np.empty_like(b))                   # captured, replayed
    assert rel(x, want) <= 1e-11
    mg.ctx.close()
