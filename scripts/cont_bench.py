"""Where the Newton continuation (BASELINE configs[0]) spends its time with the CUDA library as fieldsplit_0:
wall-clock per backend method, for the condensed-form variants (environment switches of csrc/condense.cu).

    python scripts/cont_bench.py [N [host|schur|device]]     N runs of the default variant; the last argument moves the
                                                            Schur-complement application / the whole linear solve of a
                                                            Newton step onto the device (csrc/outer.cu)
"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceBackend  # noqa: E402
from alfi_b200.synth.outer import ContinuationSolver  # noqa: E402
from alfi_b200.synth.problem import CONFIGS  # noqa: E402

CONT_CONFIG = "ldc2d-sv-k2"
CONT_RES = [10, 100] + list(range(200, 1001, 100))


class Timed:
    def __init__(self, inner):
        self.inner, self.t, self.n = inner, {}, {}

    def _wrap(self, name, *a):
        t0 = time.perf_counter()
        out = getattr(self.inner, name)(*a)
        self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
        self.n[name] = self.n.get(name, 0) + 1
        return out

    def setup(self, levels):
        out = self._wrap("setup", levels)
        ctx = self.inner.mg.ctx                      # wall-clock of the library calls inside update_operators
        for name in ("set_bsr_values", "factor", "coarse_factor", "transfer_update"):
            fn = getattr(ctx, name)

            def timed(*a, _fn=fn, _name="ctx." + name, **kw):
                t0 = time.perf_counter()
                r = _fn(*a, **kw)
                self.t[_name] = self.t.get(_name, 0.0) + time.perf_counter() - t0
                self.n[_name] = self.n.get(_name, 0) + 1
                return r
            setattr(ctx, name, timed)
        return out

    def update_operators(self, levels):
        return self._wrap("update_operators", levels)

    def update_transfers(self, levels):
        return self._wrap("update_transfers", levels)

    def apply(self, b):
        return self._wrap("apply", b)

    def setup_outer(self, *a):
        return self._wrap("setup_outer", *a)

    def schur_apply(self, *a):
        return self._wrap("schur_apply", *a)

    def outer_solve(self, *a):
        return self._wrap("outer_solve", *a)


cfg = CONFIGS[CONT_CONFIG]
OLD = {"ALFIB_CONDENSE_SHARED": "0", "ALFIB_TILE_V1": "1"}
VARIANTS = (("warm-up", {}), ("default", {}), ("per-instance blocks, tile v1", OLD), ("default", {}),
            ("per-instance blocks, tile v1", OLD), ("dense", {"dense": "1"}), ("default", {}))
if len(sys.argv) > 1:                       # python scripts/cont_bench.py N: N runs of the default variant
    VARIANTS = (("warm-up", {}),) + (("default", {}),) * int(sys.argv[1])
OUTER = sys.argv[2] if len(sys.argv) > 2 else "host"
for label, env in VARIANTS:
    for k in ("ALFIB_CONDENSE_SHARED", "ALFIB_TILE_V1"):
        os.environ.pop(k, None)
    os.environ.update({k: v for k, v in env.items() if k.startswith("ALFIB")})
    backend = Timed(DeviceBackend(cfg.m, device=0, condense=("dense" not in env)))
    solver = ContinuationSolver(cfg, backend, outer=OUTER)
    t0 = time.perf_counter()
    infos = [solver.solve(re) for re in CONT_RES]
    dt = time.perf_counter() - t0
    print(json.dumps({"variant": label, "outer": OUTER, "time_s": dt, "backend_s": backend.t, "calls": backend.n,
                      "linear_iter": [i["linear_iter"] for i in infos]}), flush=True)
    backend.inner.mg.ctx.close()
