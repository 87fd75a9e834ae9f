"""Print relative differences CUDA vs oracle for every hot-path op (run on a GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
from alfi_b200.synth.problem import build_problem
from oracle import hotpath as hp

def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

names = sys.argv[1:] or ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny"]
print("%-22s %-6s %3s %9s %9s %9s %9s %9s %9s %9s %9s" % ("config", "regime", "lvl", "kappa", "spmv", "apply", "apply_be", "prolong", "restrict", "fgmres", "cycle"))
for name in names:
    for regime in ("mild", "prod"):
        prob = build_problem(name, gamma=10.0, nu=0.2) if regime == "mild" else build_problem(name)
        mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=True)
        olv = [hp.level_from_host(l) for l in prob.levels]
        rng = np.random.default_rng(20261017)
        b = rng.standard_normal(olv[-1].n); b[olv[-1].bc_dofs] = 0
        cyc = rel(mg.apply(b, np.empty_like(b)), hp.fcycle(olv, b, prob.config.m))
        for l in range(1, len(olv)):
            lv, lc = olv[l], olv[l - 1]
            x = rng.standard_normal(lv.n); x[lv.bc_dofs] = 0
            c = rng.standard_normal(lc.n); c[lc.bc_dofs] = 0
            mats = hp.patch_matrices(lv.A, lv.offsets, lv.dofs)
            kappa = max(np.linalg.cond(M) for M in mats if M.size)
            e_spmv = rel(mg.ctx.spmv(l, x, np.empty_like(x)), lv.A @ x)
            y = mg.ctx.smoother_apply(l, x, np.empty_like(x))
            e_app = rel(y, hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs))
            # backward error of the device patch solves
            be = 0.0
            for p, M in enumerate(mats):
                if not M.size: continue
                I = lv.dofs[lv.offsets[p]:lv.offsets[p + 1]]
                u = mg.ctx.patch_inverse(l, p, M.shape[0]) @ x[I]
                be = max(be, np.linalg.norm(M @ u - x[I]) / (np.linalg.norm(M, 2) * np.linalg.norm(u) + np.linalg.norm(x[I])))
            e_pro = rel(mg.ctx.prolong(l, c, np.empty(lv.n)), hp.prolong(lv, c))
            e_res = rel(mg.ctx.restrict(l, x, np.empty(lc.n)), hp.restrict(lv, x, lc.bc_dofs))
            x0 = rng.standard_normal(lv.n); x0[lv.bc_dofs] = 0
            e_fg = rel(mg.ctx.smooth(l, prob.config.m, x, x0.copy()), hp.smooth(lv, x, x0, prob.config.m))
            print("%-22s %-6s %3d %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e" % (name, regime, l, kappa, e_spmv, e_app, be, e_pro, e_res, e_fg, cyc), flush=True)
        mg.ctx.close()
