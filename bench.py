#!/usr/bin/env python
"""Benchmark of the velocity-block multigrid hot path (BASELINE.json metric).

A "step" is one application of `fieldsplit_0` (richardson(1) + PCMG-full F-cycle with
FGMRES(m)/patch smoothing, Schoeberl transfers, direct coarse solve — alfi/solver.py:359-379)
on the ldc3d Scott-Vogelius k=3 barycentric workload (BASELINE.json configs[4] at the size that
fits one GPU: baseN 4, nref 2 — 1 458 867 velocity dofs, 4 913 macro-star patches; their inverses
are 48 GB dense, ~6 GB in the condensed block/separator form the library uses by default).  `value` = finest-level velocity dofs / time of one step, inputs resident in
HBM; `e2e` = the same through the C-ABI with pinned host vectors (H2D + D2H inside the timed
region).  The JSON line also carries the roofline of the dominant kernel (finest-level patch
apply), the CPU baseline (oracle restatement timed on a bounded sample) and the clocks seen.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_CONFIG = "ldc3d-sv-k3"
CPU_SAMPLE_CONFIG = {"ldc3d-sv-k3-literal": "ldc3d-sv-k3-half-literal", "ldc3d-sv-k3": "ldc3d-sv-k3-half", "ldc3d-sv-k3-w1": "ldc3d-sv-k3-half", "ldc3d-sv-k3-w2": "ldc3d-sv-k3-half",
                     "ldc3d-sv-k3-w4": "ldc3d-sv-k3-half", "ldc3d-sv-k3-w8": "ldc3d-sv-k3-half", "ldc3d-sv-k3-s8": "ldc3d-sv-k3-half", "ldc3d-sv-k3-n5": "ldc3d-sv-k3-half", "ldc3d-sv-k3-n6": "ldc3d-sv-k3-half", "ldc2d-sv-k2": "ldc2d-sv-k2", "ldc2d-pkp0": "ldc2d-pkp0",
                     "ldc3d-pkp0": "ldc3d-pkp0-small", "ldc3d-pkp0-mid": "ldc3d-pkp0-small", "bfs2d-sv-k2": "bfs2d-sv-k2-small",
                     "ldc3d-pkp0-l5": "ldc3d-pkp0-small", "ldc3d-pkp0-l5-re100": "ldc3d-pkp0-small", "ldc3d-sv-k3-burman": "ldc3d-sv-k3-half-burman", "ldc3d-sv-k3-half-burman": "ldc3d-sv-k3-half-burman"}
METRIC = "V-cycle DoF/s (finest-level velocity dofs per second of one fieldsplit_0 PCMG-full application)"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- algorithmic bytes
def smoother_bytes(ld):
    """B_s = sum_i (8 n_i^2 + 4 n_i) + 8 N (x read) + 8 N (y written)   (SURVEY §8d)."""
    n = ld.patches.sizes.astype(np.float64)
    return float((8 * n * n + 4 * n).sum() + 16 * ld.ndofs)


def spmv_bytes(ld):
    bs = ld.V.bs
    return float(ld.A.nnzb * (8 * bs * bs + 4) + 4 * (ld.V.nnodes + 1) + 16 * ld.ndofs)


# --------------------------------------------------------------------------- the workload, described identically by both arms
def workload_config(name):
    """`config` of the JSON line: what is being solved, not how.  Both arms (`--impl ours` / `--impl reference`) print
    exactly this dictionary.  Sizes of the 3-D Scott-Vogelius family follow from the mesh formulas of SURVEY App. B
    (box of Mx x My x Mz Kuhn cubes, Alfeld split, P3: nodes = V + 2E + F + 15T); other configurations are generated."""
    from alfi_b200.synth.problem import CONFIGS
    cfg = CONFIGS[name]
    shape = tuple(cfg.shape) if cfg.shape else (1,) * cfg.dim
    if cfg.dim == 3 and cfg.discretisation == "sv" and cfg.k == 3 and cfg.bary and cfg.domain == "ldc":
        mx, my, mz = (cfg.N * s * 2 ** cfg.nref for s in shape)
        V = (mx + 1) * (my + 1) * (mz + 1)
        sq = mx * my * (mz + 1) + my * mz * (mx + 1) + mx * mz * (my + 1)
        E = mx * (my + 1) * (mz + 1) + my * (mx + 1) * (mz + 1) + mz * (mx + 1) * (my + 1) + sq + mx * my * mz
        F = 2 * sq + 6 * mx * my * mz
        T = 6 * mx * my * mz
        ndofs, npatch, maxn = 3 * (V + 2 * E + F + 15 * T), V, (2175 if cfg.macro_expand == "all" else 1275)
    elif cfg.dim == 3 and cfg.element == "p1fb" and cfg.domain == "ldc" and not cfg.shape:
        M = cfg.N * 2 ** cfg.nref                                   # nodes = vertices + faces
        ndofs, npatch, maxn = 3 * ((M + 1) ** 3 + 12 * M ** 3 + 6 * M ** 2), (M + 1) ** 3, 111
    else:
        from alfi_b200.synth.problem import build_problem
        fine = build_problem(name).finest
        ndofs, npatch, maxn = int(fine.ndofs), int(fine.patches.npatch), int(fine.patches.sizes.max())
    return {"workload": name,
            "mesh": "Kuhn %s x 2^%d%s" % (" x ".join(str(cfg.N * s) for s in shape), cfg.nref, ", Alfeld split" if cfg.bary else ""),
            "velocity_dofs": int(ndofs), "levels": cfg.nref + 1, "smoothing": cfg.m, "re": cfg.re, "gamma": cfg.gamma,
            "rank_grid": list(cfg.shape) if cfg.shape else None, "patches_finest": int(npatch), "max_patch_dofs": int(maxn),
            "stabilisation": ("burman, weight %g (interior-facet jump term: patch operators are not sub-matrices, dense inverses + "
                              "patch corrections)" % cfg.stab_weight) if cfg.stabilisation == "burman" else "none",
            "patch_constructor": ("alfi.MacroStar, literal semantics of relaxation.py:168-177" if cfg.macro_expand == "all" else
                                  "alfi.MacroStar restricted to the open macro star (-pc_patch_construction_MacroStar_expand vertices: "
                                  "an extension; in 2-D identical to the reference, in 3-D the 1275-dof sets SURVEY §8 sizes the "
                                  "benchmark by instead of the reference's 2175-dof ones)") if cfg.patch == "macro" else "star"}


def default_workload(args, world):
    """N = 1: BASELINE.json's configuration (cfg5).  N > 1 (unless --config / --scaling strong say otherwise): the
    weak-scaling family, one cfg5-sized brick per rank."""
    if world > 1 and args.scaling == "weak" and args.config == DEFAULT_CONFIG:
        return "ldc3d-sv-k3-w%d" % world
    return args.config


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU oracle arm
def run_oracle(sample_name, steps, warmup):
    """Time the CPU restatement (oracle/) on a bounded sample of the workload."""
    from alfi_b200.synth.problem import build_problem
    from oracle import cport
    from oracle import hotpath as hp
    # all host cores, whatever the launcher put into the environment (torchrun exports OMP_NUM_THREADS=1)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cport.set_num_threads(ncores)
    t0 = time.time()
    prob = build_problem(sample_name)
    levels = [hp.level_from_host(l, "inverse") for l in prob.levels]
    cport.accelerate(levels, [None if l.patches is None else l.patches.colours for l in prob.levels])
    log("oracle sample %s: %d dofs, setup %.1fs" % (sample_name, prob.finest.ndofs, time.time() - t0))
    rng = np.random.default_rng(20261017)
    b = rng.standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0.0
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1, user_api="blas"):      # BLAS-1 only; keeps the cores for OpenMP
        for _ in range(warmup):
            hp.fcycle(levels, b, prob.config.m)
        t0 = time.perf_counter()
        for _ in range(steps):
            hp.fcycle(levels, b, prob.config.m)
        dt = (time.perf_counter() - t0) / steps
    cores = cport.num_threads()
    return {"value": prob.finest.ndofs / dt, "unit": "DoF/s", "cores": cores, "kind": "port",
            "sample": "%s: same mesh family/element/patches, %d velocity dofs, %d levels, %.2f s per F-cycle "
                      "(C/OpenMP patch apply + SpMV, numpy BLAS-1; CPU restatement, not PETSc)" % (sample_name, prob.finest.ndofs, len(levels), dt)}, dt


CONT_CONFIG = "ldc2d-sv-k2"                      # BASELINE.json configs[0]: Re 10 -> 1000 continuation
CONT_RES = [10, 100] + list(range(200, 1001, 100))


class TimedBackend:
    """Wraps a fieldsplit_0 backend and accumulates the wall time spent inside it (setup / per-Newton refresh /
    applications), so a continuation's time splits into library time and host assembly + outer loop."""

    def __init__(self, inner):
        self.inner = inner
        self.t = {"setup": 0.0, "update_operators": 0.0, "update_transfers": 0.0, "apply": 0.0}
        self.n = {k: 0 for k in self.t}

    def _timed(self, key, fn, *a):
        t0 = time.perf_counter()
        out = fn(*a)
        self.t[key] += time.perf_counter() - t0
        self.n[key] += 1
        return out

    def setup(self, levels):
        return self._timed("setup", self.inner.setup, levels)

    def update_operators(self, levels):
        return self._timed("update_operators", self.inner.update_operators, levels)

    def update_transfers(self, levels):
        return self._timed("update_transfers", self.inner.update_transfers, levels)

    def apply(self, b):
        return self._timed("apply", self.inner.apply, b)


def run_continuation(backend, label, config=None, res=None):
    """Full Newton continuation (alfi/driver.py:95-129) of BASELINE configs[0] (or `config`) around `backend` as
    the fieldsplit_0 preconditioner; assembly and the outer FGMRES/Schur loop run on the host."""
    from alfi_b200.synth.outer import ContinuationSolver
    from alfi_b200.synth.problem import CONFIGS
    config = config or CONT_CONFIG
    res = list(res or CONT_RES)
    cfg = CONFIGS[config]
    timed = TimedBackend(backend)
    solver = ContinuationSolver(cfg, timed)
    t0 = time.perf_counter()
    infos = [solver.solve(re) for re in res]
    dt = time.perf_counter() - t0
    return {"config": config, "velocity_dofs": solver.nu_dofs, "pressure_dofs": solver.np_dofs, "re": res,
            "time_s": dt, "nonlinear_iter": [i["nonlinear_iter"] for i in infos],
            "linear_iter": [i["linear_iter"] for i in infos], "final_residual": float(infos[-1]["residual"]),
            "fieldsplit_0": label,
            "time_in_fieldsplit_0_s": {k: round(v, 4) for k, v in timed.t.items()},
            "calls_of_fieldsplit_0": dict(timed.n),
            "time_on_host_s": round(dt - sum(timed.t.values()), 4)}, solver


def reference_arm(args):
    """`--impl reference`: the reference's CPU path cannot be built here (PETSc/Firedrake absent),
    so the oracle port of the same algorithm is timed on the host cores (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = default_workload(args, args.gpus)
    sample = CPU_SAMPLE_CONFIG.get(workload, workload)
    steps, warmup = max(1, args.steps), max(0, args.warmup)        # exactly the K / W of the command line
    base, dt = run_oracle(sample, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "DoF/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if (args.scaling == "strong" or "-s%d" % args.gpus in workload) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(workload), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import torch.distributed as dist
    from alfi_b200.build import build
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    from alfi_b200.synth.problem import build_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build()

    t0 = time.time()
    # N > 1, --scaling weak: one cfg5-sized brick per rank (configs ldc3d-sv-k3-w{N}), generated rank-locally
    # (alfi_b200/synth/bricks.py) — no rank ever builds a global level >= 1 — and run with distributed vectors
    weak = world > 1 and args.scaling == "weak"
    if weak:
        from alfi_b200.synth.bricks import build_rank_local
        from alfi_b200.synth.problem import CONFIGS as _CW
        args.config = default_workload(args, world)
        cfg = _CW[args.config]
        if int(np.prod(cfg.shape or (1,))) != world:
            raise SystemExit("config %s is a %s rank grid, not %d ranks" % (args.config, cfg.shape, world))
        rlp = build_rank_local(cfg, rank, verbose=(rank == 0))
        prob = None
        nlev = len(rlp.local)
        fine_ll = rlp.local[-1]
    else:
        prob = build_problem(args.config, verbose=(rank == 0))
        cfg = prob.config
        nlev = len(prob.levels)
    log("rank %d: problem built in %.1fs" % (rank, time.time() - t0))
    # N > 1: level vectors distributed (owned + ghosts per rank, neighbour exchanges) or replicated
    distributed = world > 1 and (args.vectors == "distributed" or weak)

    def make_mg(condense):
        """Device hierarchy + one per-Newton-step refresh; returns (mg, setup_s, newton_setup_s)."""
        t0 = time.time()
        uid = None
        if world > 1:
            from alfi_b200.dist import bootstrap_unique_id
            uid = bootstrap_unique_id(rank)
        if weak:
            from alfi_b200.multigrid import DistributedMultigrid
            mg = DistributedMultigrid.from_local(rlp, cfg.m, uid, device=local, deterministic=bool(args.deterministic),
                                                 torch_storage=True, condense=bool(condense),
                                                 peer_memory=bool(args.peer_memory))
        elif distributed:
            from alfi_b200.multigrid import DistributedMultigrid
            mg = DistributedMultigrid([level_input_from_synth(l) for l in prob.levels], cfg.m, rank, world, uid,
                                      device=local, deterministic=bool(args.deterministic), torch_storage=True,
                                      condense=bool(condense), peer_memory=bool(args.peer_memory))
        else:
            mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], cfg.m, device=local,
                                 deterministic=bool(args.deterministic), torch_storage=True,
                                 rank=rank, nranks=world, unique_id=uid, peer_memory=bool(args.peer_memory),
                                 condense=bool(condense))
        mg.ctx.synchronize()
        setup_s = time.time() - t0
        log("rank %d: device setup (upload + factor) %.1fs" % (rank, setup_s))
        # what every Newton step pays again (alfi re-assembles J, PCSetUp_PATCH refactors the patches and the
        # coarse LU; solver.py:320-327, 369-378): values hand-over + all patch inverses + coarse inverse.  The
        # second refresh is the steady state (the first one still allocates workspaces); the hand-over of the
        # values (host numpy -> HBM over PCIe, 1.8 GB on cfg5) is timed separately from the device work.
        lv_in = None if weak else [level_input_from_synth(l) for l in prob.levels]
        newton_setup_s, handover = None, [0.0]
        orig_set = mg.ctx.set_bsr_values

        def timed_set(*a, **k):
            mg.ctx.synchronize()
            t = time.perf_counter()
            r = orig_set(*a, **k)
            mg.ctx.synchronize()
            handover[0] += time.perf_counter() - t
            return r
        for rep in range(2):
            handover[0] = 0.0
            mg.ctx.set_bsr_values = timed_set
            t0 = time.time()
            mg.update_operators(lv_in)
            mg.ctx.synchronize()
            newton_setup_s = time.time() - t0
            mg.ctx.set_bsr_values = orig_set
            log("rank %d: per-Newton-step setup #%d %.3fs (values hand-over %.3fs, patch inverses + coarse inverse %.3fs)"
                % (rank, rep, newton_setup_s, handover[0], newton_setup_s - handover[0]))
        make_mg.handover_s = handover[0]
        # the same with the value arrays page-locked once (alfib_host_register: what a shim does with PETSc's value
        # arrays, whose addresses do not change between Newton steps): first pass registers, second is the steady state
        make_mg.pinned = None
        if lv_in is not None and hasattr(mg, "update_operators") and not distributed:
            try:
                for rep in range(2):
                    handover[0] = 0.0
                    mg.ctx.set_bsr_values = timed_set
                    t0 = time.time()
                    mg.update_operators(lv_in, pin_values=True)
                    mg.ctx.synchronize()
                    dt = time.time() - t0
                    mg.ctx.set_bsr_values = orig_set
                    log("rank %d: per-Newton-step setup, page-locked values #%d %.3fs (hand-over %.3fs)" % (rank, rep, dt, handover[0]))
                make_mg.pinned = {"per_newton_step": dt, "values_handover": handover[0], "device_work": dt - handover[0]}
            except Exception as e:      # noqa: BLE001
                mg.ctx.set_bsr_values = orig_set
                make_mg.pinned = {"error": repr(e)}
        return mg, setup_s, newton_setup_s

    def all_ok(flag):
        """True iff `flag` holds on every rank (the ranks must take the same branch)."""
        if world == 1:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    # N > 1 with condensed inverses: that combination had only its host side checked (CPU test of the rank-local
    # condensed sets) when this was written, so the sharded run verifies itself — setup on every rank, then one
    # cycle must reduce the residual — and otherwise repeats with the dense inverses, which were measured on
    # 2/4/8 GPUs.  The line says which form ran (`config.patch_inverses`, `config.fallback`).
    fallback = None
    mg = None
    try:
        mg, setup_s, newton_setup_s = make_mg(args.condense)
        ok = True
    except Exception as e:          # noqa: BLE001
        if world == 1 or not args.condense:
            raise
        log("rank %d: sharded condensed setup failed: %r" % (rank, e))
        ok = False
    if world > 1 and args.condense and not all_ok(ok):
        fallback = "sharded condensed setup failed on a rank; dense inverses used"
        if mg is not None:
            mg.ctx.close()
        mg = None
        torch.cuda.empty_cache()
        args.condense = 0
        mg, setup_s, newton_setup_s = make_mg(0)

    if weak:
        rng = np.random.default_rng(20261017 + rank)
        bnp = rng.standard_normal(fine_ll.n_local)     # this rank's local vector: owned dofs, then ghosts (refreshed by the library)
        bnp[fine_ll.bc_dofs] = 0.0
        tn = torch.tensor([float(fine_ll.n_owned), float(fine_ll.patch_ids.size)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tn)
        n, npatch_total = int(tn[0].item()), int(tn[1].item())
    else:
        n = prob.finest.ndofs
        rng = np.random.default_rng(20261017)          # same right-hand side on every rank (replicated vectors)
        bnp = rng.standard_normal(n)
        bnp[prob.finest.bc_dofs] = 0.0
        if distributed:
            bnp = mg.scatter(bnp)                      # this rank's local vector: owned dofs, then ghosts
    nvec = bnp.size
    bh = torch.empty(nvec, dtype=torch.float64).pin_memory()
    xh = torch.empty(nvec, dtype=torch.float64).pin_memory()
    bh.copy_(torch.from_numpy(bnp))
    bd = bh.cuda()
    xd = torch.empty_like(bd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduction(xvec):
        rr = torch.empty_like(xvec)
        mg.ctx.residual(nlev - 1, bd, xvec, rr)
        mg.ctx.synchronize()
        if distributed:                            # norms over the owned entries of all ranks
            t = torch.stack([(rr[:mg.n_owned] ** 2).sum(), (bd[:mg.n_owned] ** 2).sum()])
            dist.all_reduce(t)
            return float(torch.sqrt(t[0] / t[1]))
        return float(torch.linalg.norm(rr) / torch.linalg.norm(bd))

    # ---- device-resident steps (value) ------------------------------------------------------
    for _ in range(args.warmup):
        mg.apply(bd, xd)
    mg.ctx.synchronize()
    if world > 1 and args.condense and not distributed:
        red0 = reduction(xd)
        if not all_ok(np.isfinite(red0) and red0 < 0.9):
            log("rank %d: sharded condensed cycle does not reduce the residual (%.3e); dense inverses instead" % (rank, red0))
            fallback = "sharded condensed cycle failed its residual check (%.3e); dense inverses used" % red0
            mg.ctx.close()
            mg = None
            torch.cuda.empty_cache()
            args.condense = 0
            mg, setup_s, newton_setup_s = make_mg(0)
            for _ in range(args.warmup):
                mg.apply(bd, xd)
            mg.ctx.synchronize()
    stream = torch.cuda.ExternalStream(mg.ctx.stream, device=torch.device("cuda", local))
    # pass 1 (the number): K steps exactly as a user runs them — CUDA-graph replay of the cycle, no per-kernel events
    launches0 = mg.ctx.launches
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        mg.apply(bd, xd)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (mg.ctx.launches - launches0) // args.steps
    # pass 2 (the explanation): the same K steps with the library's per-kernel-family CUDA events on (eager launches:
    # the breakdown and the roofline's kernel time come from here; `profile_pass_ms_per_step` says what it cost)
    mg.ctx.profile(True)
    mg.ctx.profile_reset()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        mg.apply(bd, xd)
    p1.record(stream)
    p1.synchronize()
    barrier()
    clocks = sampler.stop()
    prof_ms = p0.elapsed_time(p1) / args.steps
    prof_fine = mg.ctx.profile_get(nlev - 1)
    prof_all = mg.ctx.profile_get(-1)
    mg.ctx.profile(False)
    patch_apply_bytes = mg.ctx.patch_apply_bytes(nlev - 1)
    mg_form = mg.ctx.patch_storage_form(nlev - 1)
    factor_bytes = mg.ctx.patch_storage_bytes(nlev - 1)

    # ---- end to end through the C-ABI with host buffers --------------------------------------
    bhn, xhn = bh.numpy(), xh.numpy()
    for _ in range(max(1, args.warmup // 2)):
        mg.apply(bhn, xhn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mg.apply(bhn, xhn)          # H2D of b, cycle, D2H of x, stream sync — all inside the call
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    # sanity: the cycle must reduce the residual (a fast wrong answer is not a result)
    red = reduction(torch.from_numpy(xhn.copy()).cuda())

    if rank != 0:
        return
    # ---- roofline (rank 0's own kernel; with N > 1 the bytes are those of rank 0's patches) of the dominant kernel ------------------------------------------------------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, sustained copy)"
    except Exception:       # noqa: BLE001
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if weak:
        sizes = np.diff(fine_ll.patch_offsets)
        fine_npatch, fine_maxn, fine_dense = npatch_total, int(sizes.max()), float((sizes.astype(float) ** 2).sum() * 8)
        has_blocks = fine_ll.patch_blocks is not None
    else:
        fine = prob.finest
        fine_npatch, fine_maxn = int(fine.patches.npatch), int(fine.patches.sizes.max())
        fine_dense = float((fine.patches.sizes.astype(float) ** 2).sum() * 8)
        has_blocks = fine.patches.blocks is not None
    # algorithmic bytes of one finest-level PCApply_PATCH on this rank: stored factors
    # (dense: 8 n_i^2; condensed: X_SS + per-block V and [D | -W]) + index data + 16 N
    condensed = bool(args.condense) and has_blocks
    bs_bytes = float(patch_apply_bytes)
    app_ms, app_calls = prof_fine["PCPATCHApply"]
    achieved = bs_bytes / (app_ms / max(app_calls, 1) * 1e-3) / 1e9 if app_calls else None
    form = mg_form
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "patch_apply_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config + ((":condensed-shared" if form == 2 else ":condensed")
                                                                if condensed else ""))
        except Exception:   # noqa: BLE001
            traffic = None
    tile = ("tile_ops_kernel" if os.environ.get("ALFIB_TILE_V1", "0")[:1] == "1" else
            "tile_ops_kernel_v2" if os.environ.get("ALFIB_TILE_TMA", "1")[:1] == "0" else "tile_ops_kernel_tma")
    roofline = {"kernel": ("%s x3 + sep_rhs_kernel + slot_sum_kernel (finest-level PCApply_PATCH, condensed inverses with "
                           "shared blocks: memset, Vf ops, separator rhs, X_SS ops, z sums, [D|-Wf] ops, bc fix-up)" % tile)
                if form == 2 else
                ("%s x3 + sep_rhs_kernel (finest-level PCApply_PATCH, condensed inverses: memset, "
                 "V ops, separator rhs, X_SS ops, [D|-W] ops, bc fix-up)" % tile) if condensed else
                "patch_apply_kernel (finest-level PCApply_PATCH: memset + colour launches + bc fix-up)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bs_bytes, "launches_timed": app_calls,
                "avg_ms": app_ms / max(app_calls, 1),
                "share_of_step": app_ms / args.steps / prof_ms if app_calls else None,
                "timed_in": "second pass of K steps with per-kernel-family CUDA events on the library stream (eager launches, "
                            "%.2f ms per step against %.2f ms with graph replay)" % (prof_ms, ms)}
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "calls_per_step": v[1] / args.steps} for k, v in prof_all.items()}

    cpu = None
    continuation = None
    if not args.no_cpu_baseline:
        try:
            cpu, _ = run_oracle(CPU_SAMPLE_CONFIG.get(args.config, args.config), 1, 1)
        except Exception as e:      # noqa: BLE001
            cpu = {"value": None, "unit": "DoF/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    if not args.no_continuation and world == 1:
        try:
            from alfi_b200.multigrid import DeviceBackend
            from alfi_b200.synth.problem import CONFIGS as _C
            mg.ctx.close()
            del mg
            torch.cuda.empty_cache()
            continuation, sdev = run_continuation(DeviceBackend(_C[CONT_CONFIG].m, device=local), "CUDA library (alfib_cycle_apply)")
            if cpu is not None:
                from oracle.backend import OracleBackend
                cref, sref = run_continuation(OracleBackend(_C[CONT_CONFIG].m), "CPU oracle (numpy)")
                cpu["continuation"] = cref
                per_newton_ok = all(abs(a - b) <= max(n, 1) for a, b, n in
                                    zip(cref["linear_iter"], continuation["linear_iter"], cref["nonlinear_iter"]))
                continuation["iteration_parity"] = bool(per_newton_ok and
                                                        cref["nonlinear_iter"] == continuation["nonlinear_iter"])
                continuation["velocity_rel_diff_vs_cpu"] = float(np.linalg.norm(sdev.u - sref.u) / np.linalg.norm(sref.u))
                continuation["pressure_rel_diff_vs_cpu"] = float(np.linalg.norm(sdev.p - sref.p) / np.linalg.norm(sref.p))
        except Exception as e:      # noqa: BLE001
            continuation = {"error": repr(e)}
        if not args.no_continuation_3d and isinstance(continuation, dict) and "error" not in continuation:
            # North-star condition 3 on the family BASELINE's metric names: the 3-D Scott-Vogelius k = 3 continuation on the
            # reference's ladder Re = 1, 10, 100, 200, ... against the committed fixture of the CPU oracle with LU patch
            # solves (tests/golden/continuation_3d_small.npz, written by scripts/cont3d.py oracle in hours of CPU time):
            # Newton counts equal, Krylov counts within +-1 per Newton step, final u / p within 1e-8.  The host stand-in
            # assembles in numpy, which is not the library's job; `outer` says where the linear solves ran.
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("cont3d", os.path.join(ROOT, "scripts", "cont3d.py"))
                cont3d = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(cont3d)
                quiet = lambda *a: print("[bench] 3-D continuation:", *a, file=sys.stderr, flush=True)   # noqa: E731
                # three_d: no stabilisation (the benchmark's operator); three_d_burman: --stabilisation-type burman,
                # weight 5e-3, what the reference's own ldc3d Scott-Vogelius job runs with (generate_submission:69-87)
                for key, name3, fname in (("three_d", args.continuation_3d, "continuation_3d_small.npz"),
                                          ("three_d_burman", args.continuation_3d + "-burman", "continuation_3d_small_burman.npz")):
                    fixture = os.path.join(ROOT, "tests", "golden", fname)
                    if not os.path.exists(fixture):
                        continue
                    try:
                        continuation[key] = cont3d.compare_with_fixture(name3, fixture, args.continuation_3d_outer,
                                                                        device=local, log=quiet)
                    except Exception as e:      # noqa: BLE001
                        continuation[key] = {"error": repr(e)}
            except Exception as e:      # noqa: BLE001
                continuation["three_d"] = {"error": repr(e)}

    total = n            # all ranks together
    wl_config = workload_config(args.config)
    generated = {"velocity_dofs": int(n), "levels": int(nlev), "patches_finest": int(fine_npatch), "max_patch_dofs": int(fine_maxn)}
    for k, v in generated.items():
        if wl_config[k] != v:
            raise SystemExit("bench: workload_config(%s)[%s] = %r but the generated problem has %r" % (args.config, k, wl_config[k], v))
    line = {
        "metric": METRIC, "value": total / (ms * 1e-3), "unit": "DoF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        # ldc3d-sv-k3-s8 is cfg5 itself cut into bricks: rank-local generation, but the total problem is fixed
        "scaling": "strong" if (args.scaling == "strong" or "-s%d" % world in args.config) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl_config,
        "implementation": {
            "factor_bytes_finest": float(factor_bytes), "dense_factor_bytes_finest": fine_dense,
            "patch_inverses": ("condensed (block/separator form, blocks shared between patches, csrc/condense.cu)"
                               if mg_form == 2 else "condensed (block/separator form, csrc/condense.cu)")
            if condensed else "dense",
            "tile_ops": tile,
            "l2_policy": "inputs larger than L2 (%.1f GB of patch inverses streamed per finest-level smoother "
                         "application, 126 MB L2)" % (factor_bytes / 1e9),
            "fallback": fallback, "deterministic": bool(args.deterministic), "parallelism": "1 GPU" if world == 1 else
            ("one mesh brick per GPU (%d), level vectors distributed (owned + ghost dofs per rank): ghost exchanges with the "
             "neighbouring ranks around every patch apply / SpMV and the dots' small all-reduces %s; level 0 replicated"
             % (world, "as single kernels over NVLink peer memory (128-bit flag-in-data stores, csrc/comm.cu)"
                if args.peer_memory else "over NCCL send/recv + all-reduce")) if distributed else
            "patches + operator rows sharded over %d GPUs, level vectors replicated; ncclAllReduce after "
            "every patch apply, grouped ncclBroadcast after every SpMV" % world},
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "DoF/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 8 * nvec, "d2h_bytes_per_step": 8 * nvec},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "breakdown_ms": breakdown, "profile_pass_ms_per_step": prof_ms, "setup_s": {"device_upload_factor": setup_s, "per_newton_step": newton_setup_s,
                    "per_newton_step_values_handover": getattr(make_mg, "handover_s", None),
                    "per_newton_step_device_work": newton_setup_s - getattr(make_mg, "handover_s", 0.0),
                    "page_locked_values": getattr(make_mg, "pinned", None),
                    "note": "steady state (second refresh); Schur-complement setup of the condensed inverses"}, "continuation": continuation,
        "residual_reduction": red,
    }
    print(json.dumps(line), flush=True)


def _finish():
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:       # noqa: BLE001
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--deterministic", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-continuation", action="store_true")
    ap.add_argument("--no-continuation-3d", action="store_true")
    ap.add_argument("--continuation-3d", default="ldc3d-sv-k3-small", metavar="CONFIG",
                    help="3-D configuration of the committed continuation fixture (tests/golden/continuation_3d_small.npz)")
    ap.add_argument("--continuation-3d-outer", default="device", choices=["host", "schur", "device"],
                    help="where the outer linear solves of the 3-D continuation run (csrc/outer.cu for schur / device)")
    ap.add_argument("--peer-memory", type=int, default=1, help="N > 1: NVLink peer-memory exchanges (default) instead of NCCL")
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"],
                    help="N > 1: the same problem sharded (strong) or one cfg5-sized brick per rank, generated rank-locally (weak)")
    ap.add_argument("--vectors", default="distributed", choices=["replicated", "distributed"],
                    help="N > 1: replicated level vectors (measured in round 1) or distributed ones (DESIGN §6.1)")
    ap.add_argument("--condense", type=int, default=1,
                    help="1 (default): condensed block/separator patch inverses where the mesh has macro structure; 0: dense")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
        _finish()


if __name__ == "__main__":
    main()
