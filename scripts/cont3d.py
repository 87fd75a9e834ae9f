"""Newton continuation of the 3-D Scott-Vogelius k = 3 lid-driven cavity to Re 5000 (north-star condition 3) around a
`fieldsplit_0` backend — the reference's own ladder, Re = 1, 10, 100, 200, ..., 5000 (examples/iters.py:33-37).

    python scripts/cont3d.py oracle CONFIG OUT.npz      CPU oracle (numpy/C port): writes the fixture (minutes to hours); an
                                                        existing OUT.npz holding a prefix of the ladder is resumed
    python scripts/cont3d.py device CONFIG FIXTURE.npz [RE,RE,...|-] [host|schur|device]
                                                        CUDA library; compares iteration counts (+-1 per Newton step) and
                                                        the final velocity / pressure (<= 1e-8) with the fixture; the last
                                                        argument moves the Schur-complement application / the whole linear
                                                        solve of a Newton step onto the device as well (csrc/outer.cu)

The host stand-in of the outer solver (alfi_b200/synth/outer.py) assembles in numpy; the time split is printed."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from alfi_b200.synth.outer import ContinuationSolver  # noqa: E402
from alfi_b200.synth.problem import CONFIGS  # noqa: E402

mode, name, path = sys.argv[1], sys.argv[2], sys.argv[3]
res = [1, 10, 100] + list(range(200, 5001, 100))
if len(sys.argv) > 4 and sys.argv[4] != "-":
    res = [float(r) for r in sys.argv[4].split(",")]
outer = sys.argv[5] if len(sys.argv) > 5 else "host"
cfg = CONFIGS[name]
if mode == "oracle":
    from oracle.backend import OracleBackend
    # patch sub-solver of the reference's Scott-Vogelius path: preonly + LU solves (solver.py:326-327, 655-659), not an
    # explicit inverse — the device applies explicit (condensed) inverses, so equal iteration counts are the evidence
    # that this difference does not matter
    backend = OracleBackend(cfg.m, mode="lu")
else:
    from alfi_b200.multigrid import DeviceBackend
    backend = DeviceBackend(cfg.m, deterministic=False)
s = ContinuationSolver(cfg, backend, outer=outer if mode != "oracle" else "host")
t0 = time.time()
rows = []
import os
if mode == "oracle" and os.path.exists(path):       # resume: the state after the last Reynolds number of the fixture
    old = np.load(path)
    nold = len(old["re"])
    assert str(old["config"]) == name and [float(r) for r in old["re"]] == [float(r) for r in res[:nold]]
    rows = [(float(a), int(b), int(c), float(d)) for a, b, c, d in zip(old["re"], old["nonlinear_iter"], old["linear_iter"], old["residual"])]
    s.u[:], s.p[:] = old["u"], old["p"]
    res = res[nold:]
    t0 -= float(old["time_s"])
    print("resuming after Re %g (%d steps, %.0fs so far)" % (rows[-1][0], nold, float(old["time_s"])), flush=True)
if mode != "oracle":
    ref = np.load(path)
    res = [float(r) for r in ref["re"]]          # the ladder (or the prefix of it) the fixture holds
for re in res:
    t1 = time.time()
    info = s.solve(re)
    rows.append((re, info["nonlinear_iter"], info["linear_iter"], float(info["residual"])))
    print("Re %6g  Newton %d  Krylov %3d  residual %.2e  %.1fs" % (re, info["nonlinear_iter"], info["linear_iter"],
                                                                 info["residual"], time.time() - t1), flush=True)
    if mode == "oracle":                         # written after every Reynolds number: a long run can be used as far as it got
        r_ = np.array(rows)
        np.savez_compressed(path, config=name, re=r_[:, 0], nonlinear_iter=r_[:, 1].astype(int), linear_iter=r_[:, 2].astype(int),
                            residual=r_[:, 3], u=s.u, p=s.p, time_s=time.time() - t0)
total = time.time() - t0
rows = np.array(rows)
if mode == "oracle":
    print("fixture written:", path, "%.0fs" % total)
else:
    assert str(ref["config"]) == name and np.array_equal(ref["re"], rows[:, 0])
    nl_ok = np.array_equal(ref["nonlinear_iter"], rows[:, 1].astype(int))
    dk = np.abs(ref["linear_iter"] - rows[:, 2].astype(int))
    k_ok = bool((dk <= ref["nonlinear_iter"]).all())
    du = float(np.linalg.norm(s.u - ref["u"]) / np.linalg.norm(ref["u"]))
    dp = float(np.linalg.norm(s.p - ref["p"]) / np.linalg.norm(ref["p"]))
    out = {"config": name, "outer": outer, "velocity_dofs": int(s.nu_dofs), "re_max": float(rows[-1, 0]), "steps": int(len(res)),
           "newton_iterations": int(rows[:, 1].sum()), "krylov_iterations": int(rows[:, 2].sum()),
           "newton_counts_equal": bool(nl_ok), "krylov_counts_within_1_per_newton_step": k_ok,
           "max_krylov_count_difference": int(dk.max()), "velocity_rel_diff": du, "pressure_rel_diff": dp,
           "time_s_device": total, "time_s_cpu_oracle": float(ref["time_s"]),
           "pass": bool(nl_ok and k_ok and du <= 1e-8 and dp <= 1e-8)}
    print(json.dumps(out), flush=True)
