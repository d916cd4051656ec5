#!/usr/bin/env python
"""bench.py -- q-evals/sec of the batched thrust-network evaluation on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (z, f, grad f, g, dense J, feasibility margin, status) over one batch of
synthetic candidates of the discretisation-20 cross vault (the configuration north_star quotes its target on).

    python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 2 --warmup 1     # CPU arm: the oracle port of the reference callbacks
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 10 --warmup 3               # candidates sharded across ranks, weak scaling

Prints ONE JSON line on rank 0.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "q-evals/sec (z+constraints+Jacobian, fp64)"
UNIT = "evals/s"


# ------------------------------------------------------------------------------------------------------
def build_workload(name, n_states=12):
    from compas_tno_b200 import workloads as W
    if name.startswith("crossvault_disc"):
        n = int(name[len("crossvault_disc"):])
        P, xs = W.crossvault_problem(n, n_states=n_states)
        objective = "min"
    elif name == "dome_16x20":
        P, xs = W.dome_problem(20, 16, constraints=("funicular", "envelope", "reac_bounds"), n_states=n_states)
        objective = "max"
    else:
        raise SystemExit("unknown workload %r" % name)
    return P, xs, objective


def candidates(P, xs, N, seed):
    from compas_tno_b200 import workloads as W
    return W.sample_candidates(P, xs, N, seed=seed, basis=getattr(P, "candidate_basis", None))


def slsqp_config1(device):
    """BASELINE configuration 1 (examples/_tutorial: cross vault, span 10 m, thickness 0.5 m, discretisation 10, min
    thrust, SLSQP): scipy's fmin_slsqp exactly as solvers/scipy.py:172 calls it, driven once by the GPU drop-in
    callbacks and once by the CPU port of the reference callbacks (oracle/, checker arm).  Same start, same bounds."""
    import copy
    from scipy.optimize import fmin_slsqp
    from compas_tno_b200 import accelerate
    from oracle import callbacks_np as onp
    P, xs, objective = build_workload("crossvault_disc10")
    x0 = np.asarray(xs, dtype=np.float64).copy()
    lbz, ubz = P.lb.ravel()[P.fixed], P.ub.ravel()[P.fixed]
    bounds = [(-1e4, 1e-8)] * P.k + [(float(a), float(b)) for a, b in zip(lbz, ubz)]
    kw = dict(bounds=bounds, iter=500, iprint=0, full_output=True)

    Pg = copy.copy(P)
    accelerate(Pg, objective, device=device)
    Pg.fobj(x0, Pg)                                   # plan creation and first launch outside the timed region
    t0 = time.perf_counter()
    og = fmin_slsqp(Pg.fobj, x0, args=[Pg], f_ieqcons=Pg.fconstr, fprime=Pg.fgrad, fprime_ieqcons=Pg.fjac, **kw)
    t_gpu = time.perf_counter() - t0

    M = onp.clone_problem(P)
    fobj, fgrad = onp.objective_selector(objective)
    # the reference objectives read the state the constraint call left: same order here
    t0 = time.perf_counter()
    oc = fmin_slsqp(lambda x, m_: (onp.constr_wrapper(x, M), fobj(x, M))[1], x0, args=[None],
                    f_ieqcons=lambda x, m_: onp.constr_wrapper(x, M),
                    fprime=lambda x, m_: (onp.constr_wrapper(x, M), np.asarray(fgrad(x, M)).ravel())[1],
                    fprime_ieqcons=lambda x, m_: onp.sensitivities_wrapper(x, M), **kw)
    t_cpu = time.perf_counter() - t0
    return {"what": "config 1: crossvault_disc10 min thrust, scipy fmin_slsqp(iter=500) as solvers/scipy.py:172 calls it",
            "gpu_callbacks": {"wall_s": t_gpu, "iterations": int(og[2]), "exit_mode": int(og[3]), "f": float(og[1])},
            "cpu_port_callbacks": {"wall_s": t_cpu, "iterations": int(oc[2]), "exit_mode": int(oc[3]), "f": float(oc[1])},
            "max_abs_dx": float(np.max(np.abs(og[0] - oc[0]))),
            "note": "wall time includes scipy's own QP work per iteration, identical in both arms"}


def _block_rng(seed, lo, hi, block=4096):
    """Per-candidate random streams that do not depend on how the sweep is sharded: candidate j always draws from the
    generator of its 4096-block, so rank slices [lo, hi) of any world size see the same candidates."""
    for b in range(lo // block, (hi - 1) // block + 1):
        a0, a1 = max(lo, b * block), min(hi, (b + 1) * block)
        yield a0 - lo, a1 - lo, a0 - b * block, a1 - b * block, np.random.default_rng([seed, b]), block


def _ladder(center, lo, hi, seed, sigma_lo, sigma_hi):
    """Rows [lo, hi) of the sharding-independent candidate ladder around ``center`` (see workloads.sample_around)."""
    xc = np.asarray(center, dtype=np.float64).ravel()
    X = np.empty((hi - lo, len(xc)))
    for o0, o1, b0, b1, rng, block in _block_rng(seed, lo, hi):
        sig = np.exp(rng.uniform(np.log(sigma_lo), np.log(sigma_hi), size=(block, 1)))
        xi = rng.uniform(-1, 1, size=(block, len(xc)))
        X[o0:o1] = xc[None, :] * (1.0 + sig[b0:b1] * xi[b0:b1])
    return np.ascontiguousarray(X)


def _bcast(arr, device):
    """Rank 0's copy of a small host array on every rank (set-up only): the centre of a candidate ladder comes out of a
    linear-programming loop whose last bits may differ between processes; the sweep must not depend on the sharding."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(torch.device("cuda", device))
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def sweep_specs():
    """The BASELINE configurations 2-5 (+ a full-output discretisation-40 sweep) as sweeps sharded over the ranks.
    ``total(world)`` is the number of candidates of the whole sweep; ``build(device)`` returns the problem, its plan and
    ``make(lo, hi)`` -> (X, extras) for the candidates [lo, hi) of the sweep, identical for every sharding.  Candidates are
    scattered around a strictly feasible state (workloads.max_margin_state) so that part of every sweep is feasible and
    the best-feasible reduction has something to find."""
    from compas_tno_b200 import workloads as W
    from compas_tno_b200.plan import TnoPlan
    FULL = ("f", "grad", "g", "J", "z", "gmin", "status")
    LIGHT = ("f", "gmin", "status")
    cache = {}

    def disc40(device):
        if "disc40" not in cache:
            P, xs, obj = build_workload("crossvault_disc40", n_states=6)
            plan = TnoPlan(P, objective=obj, device=device)
            xc, margin = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
            xc = _bcast(xc, device)
            cache["disc40"] = (P, plan, obj, xc, margin)
        return cache["disc40"]

    def dome_multistart(device):
        P, xs, obj = build_workload("dome_16x20", n_states=6)
        plan = TnoPlan(P, objective=obj, device=device)
        xc, margin = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
        xc = _bcast(xc, device)
        return dict(P=P, plan=plan, names=FULL, margin=margin,
                    make=lambda lo, hi: (_ladder(xc, lo, hi, 20261019 + 2, 1e-4, 0.1), {}))

    def fan_thickness(device):
        # fan form diagram of the reference fixture (tests/golden/fan_q_zb_max.npz carries its arrays) under the
        # closed-form cross-vault envelope the fixture's lb / ub follow; the pavilion-vault envelope of the
        # configuration is not pinned by any fixture, so the cross vault stands in for it
        import copy
        from compas_tno_b200.envelopes import CrossVaultEnvelope
        from compas_tno_b200.problems import FixedProblem
        z = np.load(os.path.join(ROOT, "tests", "golden", "fan_q_zb_max.npz"))
        P = FixedProblem.from_arrays(z["xyz"], z["edges"], z["fixed"], z["loads"], z["lb"], z["ub"], z["target"], z["q_form"],
                                     ind=z["ind"])
        P.B, P.d = z["B"].copy(), z["d"].reshape(-1, 1).copy()
        P.variables, P.constraints, P.features = ["q", "zb"], ["funicular", "envelope"], ["fixed"]
        P.thk, P.envelope = 0.5, CrossVaultEnvelope((0.0, 10.0), (0.0, 10.0), 0.5)
        p0 = TnoPlan(P, objective="max", device=device)           # the design: widest margin at thickness 0.50
        xc, margin = W.max_margin_state(p0, z["x"][0], rows=W.margin_rows(P, p0))
        xc = _bcast(xc, device)
        p0.close()
        Pt = copy.copy(P)
        Pt.variables = ["q", "zb", "t"]
        plan = TnoPlan(Pt, objective="t", device=device)

        def make(lo, hi):
            X = np.repeat(np.concatenate([xc, [0.5]])[None, :], hi - lo, axis=0)
            X[:, -1] = 0.50 - np.arange(lo, hi) * (0.45 / 255.0)
            return np.ascontiguousarray(X), {}
        return dict(P=Pt, plan=plan, names=FULL, margin=margin, make=make)

    def disc40_montecarlo(device):
        P, plan, obj, xc, margin = disc40(device)
        n = P.n
        ub0, lb0 = P.ub.ravel(), P.lb.ravel()

        def make(lo, hi):
            X = _ladder(xc, lo, hi, 20261019 + 4, 1e-5, 1e-2)
            pzs = np.empty(hi - lo)
            bounds = np.empty((hi - lo, 2 * n))
            for o0, o1, b0, b1, rng, block in _block_rng(20261019 + 40, lo, hi):
                pz = np.clip(rng.normal(1.0, 0.05, size=block), 0.8, 1.2)
                du = rng.normal(0.0, 0.02, size=(block, 2 * n))        # intrados / extrados heights perturbed by 2 cm (1 sigma)
                pzs[o0:o1] = pz[b0:b1]
                bounds[o0:o1, :n] = ub0[None, :] + du[b0:b1, :n]
                bounds[o0:o1, n:] = lb0[None, :] + du[b0:b1, n:]
            return X, {"pz_scale": np.ascontiguousarray(pzs), "bounds": np.ascontiguousarray(bounds)}
        return dict(P=P, plan=plan, names=LIGHT, margin=margin, make=make, keep_plan=True)

    def disc40_full(device):
        P, plan, obj, xc, margin = disc40(device)
        return dict(P=P, plan=plan, names=FULL, margin=margin, keep_plan=True,
                    make=lambda lo, hi: (_ladder(xc, lo, hi, 20261019 + 41, 1e-4, 0.1), {}))

    def dome_lambdh(device):
        P0, x0, _ = build_workload("dome_16x20", n_states=6)
        p0 = TnoPlan(P0, objective="max", device=device)           # vertically loaded dome: the state the load is added to
        xc, margin = W.max_margin_state(p0, x0, rows=W.margin_rows(P0, p0))
        xc = _bcast(xc, device)
        p0.close()
        P, xs = W.dome_direction_problem(20, 16, n_states=6)
        plan = TnoPlan(P, objective="max_load", device=device)
        k = P.k
        NT = 8192

        def make(lo, hi):
            theta = 2.0 * np.pi * np.arange(lo, hi) / NT
            load_dir = np.ascontiguousarray(np.column_stack([np.cos(theta), np.sin(theta)]))
            X = np.empty((hi - lo, plan.nvar))
            for o0, o1, b0, b1, rng, block in _block_rng(20261019 + 5, lo, hi):
                lam = np.exp(rng.uniform(np.log(1e-4), np.log(0.05), size=block))[b0:b1]    # load multiplier ladder
                shift = load_dir[o0:o1, [0]] * P.lambdh_shift[None, :] + load_dir[o0:o1, [1]] * P.lambdh_shift_b[None, :]
                X[o0:o1, :k] = xc[None, :k] - lam[:, None] * shift      # q = B y* + lambdh (d0(theta) - B B^+ d0(theta))
                X[o0:o1, k:-1] = xc[None, k:]
                X[o0:o1, -1] = lam
            return np.ascontiguousarray(X), {"load_dir": load_dir}
        return dict(P=P, plan=plan, names=FULL, margin=margin, make=make)

    return [
        {"name": "dome_16x20", "config": 2, "scaling": "strong", "total": lambda world: 4096, "build": dome_multistart,
         "what": "config 2: dome, radial form, max thrust + reac_bounds, 4096 multi-start candidates scattered (relative sigma 1e-4 .. 0.1) "
                 "around the widest-margin state"},
        {"name": "fan_thickness", "config": 3, "scaling": "strong", "total": lambda world: 256, "build": fan_thickness,
         "what": "config 3: fan form diagram (reference fixture), thickness sweep 0.50 -> 0.05 m in 256 steps at the widest-margin "
                 "state of thickness 0.50, objective t; cross-vault closed-form envelope (pavilion envelope unpinned)"},
        {"name": "crossvault_disc40_mc", "config": 4, "scaling": "weak", "total": lambda world: 8192 * world, "build": disc40_montecarlo,
         "reps": 3,
         "what": "config 4: cross vault disc 40, Monte-Carlo: self-weight scale N(1, 0.05^2) clipped to [0.8, 1.2] (pz_scale) AND "
                 "intrados / extrados heights perturbed N(0, 2 cm) per vertex (bounds), distinct x per candidate; 8192 candidates "
                 "per GPU = 65,536 on 8 GPUs; objectives + feasibility only"},
        {"name": "crossvault_disc40_full", "config": 4, "scaling": "weak", "total": lambda world: 8192 * world, "build": disc40_full,
         "reps": 3,
         "what": "cross vault disc 40, every output incl. the dense 10082 x 29 Jacobian, 8192 candidates per GPU"},
        {"name": "dome_lambdh", "config": 5, "scaling": "strong", "total": lambda world: 8192, "build": dome_lambdh, "reps": 3,
         "what": "config 5: dome, horizontal-load multiplier lambdh as a variable + reac_bounds, 8192 load directions "
                 "theta_j = 2 pi j / 8192 (d0(theta) = cos d0x + sin d0y per candidate), multiplier ladder 1e-4 .. 0.05, batched Jacobians"},
    ]


def run_sweeps(dev, local, world, rank, barrier):
    """Every sweep of sweep_specs(), sharded over the ranks through distributed.ShardedSweep: each rank evaluates its
    contiguous slice on its own plan replica; ONE packed all_gather of (f, gmin, status) per sweep is the only exchange,
    every rank then knows the best feasible candidate.  Timed with CUDA events between barriers, max over ranks."""
    import torch
    import torch.distributed as dist
    from compas_tno_b200.distributed import ShardedSweep, best_of_packed
    sweeps = {}
    plans = []
    for spec in sweep_specs():
        built = spec["build"](local)
        plan, namesw = built["plan"], built["names"]
        NT = int(spec["total"](world))
        sw = ShardedSweep(plan, NT)
        Xh, extras_h = built["make"](sw.lo, sw.hi)
        nloc = sw.hi - sw.lo
        Xw = torch.from_numpy(Xh).to(dev)
        extras = {key: torch.from_numpy(val).to(dev) for key, val in extras_h.items()}
        outw = {key: torch.empty(plan._shape(key, nloc), device=dev, dtype=torch.int32 if key == "status" else torch.float64)
                for key in namesw}
        gbuf = None

        def one():
            nonlocal gbuf
            res = sw.evaluate(Xw, namesw, out=outw, **extras)
            gbuf = sw.gather_packed(res, out=gbuf if world > 1 else None)
            return gbuf

        for _ in range(2):
            one()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = spec.get("reps", 5)
        a0.record()
        for _ in range(reps):
            packed = one()
        a1.record()
        barrier()
        msw = a0.elapsed_time(a1) / reps
        if world > 1:
            t = torch.tensor([msw], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            msw = float(t.item())
        best_f, best_i, nfeas = best_of_packed(packed, tol=1e-9)
        bad = int((packed[:, 2] != 0).sum().item())
        sweeps[spec["name"]] = {
            "what": spec["what"], "baseline_config": spec["config"], "candidates": NT, "candidates_per_gpu": nloc, "n_gpus": world,
            "scaling": spec["scaling"], "outputs": list(namesw), "per_candidate_inputs": sorted(extras),
            "wall_ms_per_sweep": msw, "evals_per_s": NT / (msw * 1e-3),
            "kernel_family": {1: "interleaved", 2: "onchip"}.get(plan.info.get("last_kernel") or plan.info["kernel"]),
            "nvar": plan.nvar, "ng": plan.ng, "bytes_per_eval": plan.bytes_per_eval,
            "hbm_frac_whole_step": (NT / world) / (msw * 1e-3) * plan.bytes_per_eval / 1e9 / measured_peak()[0] if "J" in namesw else None,
            "status_nonzero": bad, "feasible": nfeas, "centre_margin": built["margin"],
            "best_feasible": {"f": best_f if nfeas else None, "index": best_i},
            "exchange": "one all_gather of (f, gmin, status), 24 B per candidate" if world > 1 else "none (single GPU)"}
        del outw, Xw, extras
        if built.get("keep_plan"):
            plans.append(plan)
        else:
            plan.close()
        torch.cuda.empty_cache()
    for pl in set(plans):
        pl.close()
    return sweeps


def run_multistart(dev, local, world, rank, barrier):
    """Configuration 2 as real optimisations: S starting points sharded over the ranks, every rank drives its
    slice with the batched SQP driver (compas_tno_b200.multistart: evaluations, QP subproblems and BFGS updates on the
    device, iterates resident in HBM); the best converged feasible optimum is found with one all_gather of (f, gmin, status).
    Wall time between barriers, max over ranks."""
    import time
    import torch
    import torch.distributed as dist
    from compas_tno_b200 import workloads as W
    from compas_tno_b200.distributed import shard_bounds
    from compas_tno_b200.multistart import MultiStartSolver, variable_bounds
    from compas_tno_b200.plan import TnoPlan
    out = {}
    for name, wl, S_total, spread, max_iter, cfg in (("multistart_crossvault_disc10", "crossvault_disc10", 4096, 0.02, 100, 1),
                                                     ("multistart_dome_16x20", "dome_16x20", 1024, 0.02, 100, 2)):
        P, xs, obj = build_workload(wl, n_states=6)
        plan = TnoPlan(P, objective=obj, device=local)
        xc, margin = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
        xc = _bcast(xc, local)
        lb, ub = variable_bounds(P)
        lo, hi = shard_bounds(S_total, world, rank)
        X0 = _ladder(xc, lo, hi, 20261019 + 60 + cfg, 0.2 * spread, spread)
        solver = MultiStartSolver(plan, lb, ub, acc=1e-6, max_iter=max_iter)
        solver.solve(X0[:min(64, len(X0))])        # warm-up: buffers, kernels, clocks
        barrier()
        t0 = time.perf_counter()
        res = solver.solve(X0)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        ok = res.success & (res.gmin >= -1e-6) & (res.status == 0)
        stats = torch.tensor([dt, float(ok.sum()), float(res.n_evals), float(res.f[ok].min()) if ok.any() else float("inf"),
                              float(res.n_iterations)], device=dev, dtype=torch.float64)
        if world > 1:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            mn = stats.clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dt, nok, nev, fbest, nit = float(mx[0]), float(sm[1]), float(sm[2]), float(mn[3]), float(mx[4])
        else:
            dt, nok, nev, fbest, nit = (float(v) for v in stats)
        fo = res.f[ok]
        out[name] = {
            "what": "config %d as %d simultaneous optimisations (%s, objective %s): starts scattered %.1f %% around the widest-margin "
                    "state, batched SQP on the device (interior-point QP on the dense J, damped BFGS, L1-merit line search), "
                    "SLSQP's stopping rule at acc 1e-6" % (cfg, S_total, wl, obj, 100 * spread),
            "baseline_config": cfg, "starts": S_total, "starts_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong",
            "wall_s": dt, "optimisations_per_s": S_total / dt, "converged_feasible": int(nok), "sqp_iterations": int(nit),
            "evaluations": int(nev), "evals_per_s": nev / dt, "qp_iterations_mean": res.qp_iterations,
            "f_best": fbest if np.isfinite(fbest) else None,
            "f_spread_rank0": [float(fo.min()), float(np.median(fo)), float(fo.max())] if len(fo) else None,
            "nvar": plan.nvar, "ng": plan.ng, "centre_margin": margin}
        plan.close()
        torch.cuda.empty_cache()
    return out


def run_geometry_montecarlo(dev, local, world, rank, barrier, n_div=20, G_total=64, per_geometry=1024):
    """The geometry half of configuration 4: G_total plan geometries of the cross form diagram (grid-line
    spacings perturbed, compas_tno_b200.preprocess.symmetric_grid_perturbations) sharded over the ranks.  For its slice a
    rank computes B and d of every geometry in ONE batched pre-processing call on the device (blocked Householder QR of
    E[:, dep], trailing updates on the FP64 tensor pipe), builds one plan per geometry and evaluates `per_geometry`
    self-weight-scaled candidates on each.  The host recipe of the reference (numpy pinv, problems.py:516-528) is timed on
    one geometry beside it."""
    import time
    import torch
    import torch.distributed as dist
    from compas_tno_b200 import preprocess as pre
    from compas_tno_b200 import workloads as W
    from compas_tno_b200.distributed import shard_bounds
    from compas_tno_b200.plan import TnoPlan
    P0, x_star = W.crossvault_problem(n_div, n_states=6)
    xy0 = np.asarray(P0.X, dtype=np.float64)[:, :2]
    XY = pre.symmetric_grid_perturbations(xy0, n_div, G_total, sigma=0.02, seed=20261019 + 43)
    lo, hi = shard_bounds(G_total, world, rank)
    ph = np.zeros(2 * P0.ni)
    timings = {}
    for tag, use_mma in (("warm", True), ("dfma", False), ("dmma", True)):
        barrier()
        t0 = time.perf_counter()
        res = pre.batched_B_d(P0, XY[lo:hi], ph=ph, use_mma=use_mma, device=local)
        torch.cuda.synchronize()
        timings[tag] = time.perf_counter() - t0
    fits = res.fits()
    Bh, dh = res.B.cpu().numpy(), res.d.cpu().numpy()
    # host recipe on one geometry (rank 0), same inputs
    cpu_s = err = None
    if rank == 0:
        from scipy.sparse import diags, vstack
        from scipy.sparse import csr_matrix
        Ct = csr_matrix(P0.C).transpose().tocsr()[list(P0.free), :]
        u, v = P0.C @ XY[lo, :, 0], P0.C @ XY[lo, :, 1]
        t0 = time.perf_counter()
        E = vstack((Ct @ diags(np.asarray(u).ravel()), Ct @ diags(np.asarray(v).ravel()))).toarray()
        ind = list(P0.ind)
        dep = [e for e in range(P0.m) if e not in set(ind)]
        Edinv = -np.linalg.pinv(E[:, dep], rcond=1e-17)
        Bref = np.zeros((P0.m, len(ind)))
        Bref[dep] = Edinv @ E[:, ind]
        Bref[ind] = np.identity(len(ind))
        cpu_s = time.perf_counter() - t0
        err = float(np.abs(Bh[0] - Bref).max() / np.abs(Bref).max())
    # one plan per geometry, per_geometry candidates each: self-weight scale N(1, 0.05^2) clipped to [0.8, 1.2]
    X = torch.from_numpy(W.sample_candidates(P0, x_star, per_geometry, sigma=0.01)).to(dev)
    rng = np.random.default_rng([20261019 + 44, rank])
    pzs = torch.from_numpy(np.clip(rng.normal(1.0, 0.05, size=per_geometry), 0.8, 1.2)).to(dev)
    names = ("f", "gmin", "status")
    outw = None
    t_plan = t_eval = 0.0
    feasible = bad = 0
    barrier()
    t_all0 = time.perf_counter()
    for g in range(hi - lo):
        t0 = time.perf_counter()
        Pg = pre.problem_with_geometry(P0, XY[lo + g], Bh[g], dh[g])
        plan = TnoPlan(Pg, objective="min", device=local)
        t_plan += time.perf_counter() - t0
        if outw is None:
            outw = {key: torch.empty(plan._shape(key, per_geometry), device=dev, dtype=torch.int32 if key == "status" else torch.float64)
                    for key in names}
        t0 = time.perf_counter()
        plan.evaluate_device(X, names, out=outw, pz_scale=pzs)
        torch.cuda.synchronize()
        t_eval += time.perf_counter() - t0
        feasible += int(((outw["gmin"] >= -1e-9) & (outw["status"] == 0)).sum())
        bad += int((outw["status"] != 0).sum())
        plan.close()
    barrier()
    t_all = time.perf_counter() - t_all0
    vals = torch.tensor([timings["dmma"], timings["dfma"], t_plan, t_eval, t_all], device=dev, dtype=torch.float64)
    cnt = torch.tensor([float(fits.sum()), float(feasible), float(bad)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    pre_s, pre_fma_s, plan_s, eval_s, all_s = (float(v) for v in vals)
    M, N = 2 * P0.ni, P0.m - P0.k
    flops = 2.0 * M * N * N - 2.0 / 3.0 * N ** 3 + 4.0 * M * N * (P0.k + 1)
    return {
        "what": "config 4, geometry half: %d plan geometries of the disc-%d cross form (grid-line spacings perturbed 2 %% of the "
                "cell, symmetric), B and d of every geometry by ONE batched device call (blocked Householder QR of the %d x %d "
                "matrix E[:, dep]), one plan per geometry, %d self-weight-scaled candidates each (f, gmin, status)"
                % (G_total, n_div, M, N, per_geometry),
        "baseline_config": 4, "geometries": G_total, "geometries_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong",
        "candidates": G_total * per_geometry,
        "preprocess_s": pre_s, "geometries_per_s": G_total / pre_s, "preprocess_tflops": flops * (hi - lo) / pre_s / 1e12,
        "preprocess_dfma_s": pre_fma_s, "tensor_pipe_speedup": pre_fma_s / pre_s,
        "independents_still_fit": int(cnt[0]), "plan_creation_s": plan_s, "evaluation_s": eval_s, "wall_s_after_preprocess": all_s,
        "evals_per_s_incl_plans": G_total * per_geometry / all_s, "feasible": int(cnt[1]), "status_nonzero": int(cnt[2]),
        "host_recipe_s_per_geometry": cpu_s, "host_recipe": "numpy pinv of E[:, dep] as problems.py:516-528, one geometry, rank 0",
        "max_rel_diff_B_vs_host": err,
        "speedup_vs_host_per_geometry": (cpu_s * G_total / pre_s) if cpu_s else None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).
    The sampler is started before the warm-up (nvidia-smi needs ~0.1 s to start); only samples whose time stamp falls
    inside [mark_begin, mark_end] are reported."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons, total = [], [], [], set(), 0
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                vals = (float(f[2]), float(f[3]), float(f[4]))
            except ValueError:
                continue
            total += 1
            if self.t0 is not None and not (self.t0 - 0.02 <= ts <= (self.t1 or ts) + 0.02):
                continue
            sm.append(vals[0])
            smax.append(vals[1])
            power.append(vals[2])
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "samples_total": total, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "samples_total": total, "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Evaluate a slice of candidates with the oracle port (one private Problem copy per worker)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    import warnings
    warnings.simplefilter("ignore")
    from threadpoolctl import threadpool_limits
    P, X, objective = args
    from oracle import callbacks_np as onp
    M = onp.clone_problem(P)
    t = time.perf_counter()
    with threadpool_limits(limits=1):      # one BLAS thread per worker process: no oversubscription
        for x in X:
            onp.evaluate(x, M, objective)
    return time.perf_counter() - t


def cpu_evals_per_s(P, X, objective, cores):
    """Oracle port (same scipy calls and call order as the reference callbacks) on ``cores`` processes."""
    import multiprocessing as mp
    from threadpoolctl import threadpool_limits
    chunks = [c for c in np.array_split(X, cores) if len(c)]
    with threadpool_limits(limits=1):
        if cores == 1:
            t = time.perf_counter()
            _cpu_worker((P, chunks[0], objective))
            return len(X) / (time.perf_counter() - t)
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            pool.map(_cpu_worker, [(P, c[:1], objective) for c in chunks])   # spin the workers up (untimed)
            t = time.perf_counter()
            pool.map(_cpu_worker, [(P, c, objective) for c in chunks])
            return len(X) / (time.perf_counter() - t)


def strip_problem(P):
    """Picklable, light copy for the worker processes."""
    from oracle import callbacks_np as onp
    return onp.clone_problem(P)


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the pure-Python reference cannot travel to the box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P, xs, objective = build_workload(args.workload, n_states=4)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = max(cores * args.cpu_evals_per_core, cores)
    X = candidates(P, xs, per_step * (args.steps + args.warmup), seed=20261019)
    SP = strip_problem(P)
    for w in range(args.warmup):
        cpu_evals_per_s(SP, X[w * per_step:(w + 1) * per_step], objective, cores)
    dt = 0.0
    for s in range(args.steps):
        lo = (args.warmup + s) * per_step
        dt += per_step / cpu_evals_per_s(SP, X[lo:lo + per_step], objective, cores)   # worker start-up is not timed
    value = per_step * args.steps / dt
    sample = "%d candidates per step (%d per core) of %s, oracle/callbacks_np.py: constr_wrapper + sensitivities_wrapper + objective + gradient" % (
        per_step, args.cpu_evals_per_core, args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "candidates_per_step": per_step, "objective": objective},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from compas_tno_b200.plan import TnoPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # host threads of the Jacobian fetch: the ranks of one box share its cores
    os.environ.setdefault("TNO_HOST_THREADS", str(max(2, min(16, (os.cpu_count() or 16) // world))))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from compas_tno_b200.distributed import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local) if world > 1 else None   # pinned buffers and fill threads on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P, xs, objective = build_workload(args.workload)
    plan = TnoPlan(P, objective=objective, device=local, kernel=args.kernel)
    info = plan.info
    B_alg = plan.bytes_per_eval
    N = args.batch                                       # candidates per GPU per step (weak scaling)
    X = candidates(P, xs, N, seed=20261019 + rank)       # each rank owns a different slice of the sweep
    x_dev = torch.from_numpy(X).to(dev)
    names = ("f", "grad", "g", "J", "z", "gmin", "status")
    out = {k: torch.empty(plan._shape(k, N), device=dev, dtype=torch.int32 if k == "status" else torch.float64) for k in names}
    from compas_tno_b200.distributed import best_of_packed, gather_packed
    gather_buf = torch.empty((world * N, 3), device=dev, dtype=torch.float64) if world > 1 else None
    packed_all = None

    def step():
        nonlocal packed_all
        plan.evaluate_device(x_dev, names, out=out)
        if world > 1:   # the path's only exchange: ONE all_gather of (f, gmin, status), 24 B / candidate
            packed_all = gather_packed(out["f"], out["gmin"], out["status"], world * N, out=gather_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    # The per-kernel CUDA events of the library sit between the launches.  With two launches per step (on-chip family) they
    # cost nothing and the timed region is profiled directly; the level-scheduled interleaved family issues ~10^3 launches
    # per step, where the events themselves take ~10 % -- there the timed region runs unprofiled and the kernel
    # durations come from a second pass of the same K steps.
    profiled_inline = (plan.info.get("last_kernel") or plan.info["kernel"]) == 2
    launches0 = plan.launch_count
    if profiled_inline:
        plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = plan.launch_count - launches0
    profiled_ms = ms
    if not profiled_inline:
        plan.profile(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            step()
        p1.record()
        barrier()
        profiled_ms = p0.elapsed_time(p1)
    prof = plan.profile_read()
    plan.profile(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * N * args.steps / (ms * 1e-3)
    status_bad = int((out["status"] != 0).sum().item())

    # ---- roofline of the dominant kernel (CUDA-event timed inside the library, same timed region) ----
    # Algorithmic bytes are attributed to the kernel that moves them: in the on-chip family the candidate-independent
    # funicular rows of J (2m x nvar doubles) are written by the streaming k_fill kernel ("k_emit(J)"), everything else
    # (x in; z, f, grad, g, the envelope / reaction rows of J out) by the fused k_onchip kernel.
    peak, peak_src = measured_peak()
    roofline = None
    if prof:
        jfun_bytes = 8 * 2 * plan.m * plan.nvar if "funicular" in plan.constraints else 0
        info = plan.info                       # after the timed calls: last_kernel = the family that actually ran
        family = info.get("last_kernel") or info["kernel"]
        total_kernel_ms = sum(v[0] for v in prof.values())
        profile_pass = "timed region" if profiled_inline else "second pass of the same %d steps (%.3f ms/step with events)" % (args.steps, profiled_ms / args.steps)
        if family == 2:
            # on-chip family: two kernels, each moves its own share of the algorithmic bytes exactly once
            kbytes = {"k_onchip": B_alg - jfun_bytes, "k_emit(J)": jfun_bytes}
            name, (kms, kcount) = max(prof.items(), key=lambda kv: kv[1][0])
            bytes_per_eval = kbytes.get(name, 0)
            evals_per_launch = N * args.steps / kcount
            achieved = bytes_per_eval * evals_per_launch / (kms / kcount * 1e-3) / 1e9
            roof_kernel, share = name, kms / total_kernel_ms
        else:
            # interleaved family: ~10^3 launches per step keep their state in HBM scratch and only the emitters write
            # outputs, so no single kernel owns the algorithmic bytes.  The roofline figure is the WHOLE STEP on B_alg
            # (the dominant kernel is named for information, without a byte claim of its own).
            kbytes = {"k_emit(J)": 8 * plan.ng * plan.nvar}
            name, (kms, kcount) = max(prof.items(), key=lambda kv: kv[1][0])
            bytes_per_eval = B_alg
            evals_per_launch = N
            achieved = value / world * B_alg / 1e9
            roof_kernel, share = "whole step (all launches; dominant by time: %s)" % name, 1.0
        per_kernel = {}
        for kname, (ms_k, cnt_k) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            ent = {"ms_per_step": ms_k / args.steps, "launches_per_step": cnt_k / args.steps}
            if kname in kbytes and ms_k > 0:
                ent["bytes_per_eval"] = kbytes[kname]
                ent["GB/s"] = kbytes[kname] * N * args.steps / (ms_k * 1e-3) / 1e9
            per_kernel[kname] = ent
        traffic = None   # DRAM bytes per launch (per step for the interleaved family) from the committed ncu --set full captures
        for tfile in ("r2_traffic.json", "r1_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tfile)))
                ent = tj.get(args.workload, tj if args.workload == "crossvault_disc20" else {})
                key = name if family == 2 else "whole_step"
                if key in ent:
                    traffic = ent[key]["dram_bytes_per_candidate"] * evals_per_launch
                    break
            except Exception:
                continue
        fp64_peak, fp64_src = 40.0, "nominal B200 FP64 (measurement failed)"
        try:
            from compas_tno_b200 import _lib as _tl
            fp64_peak = _tl.device_peak("fp64_fma", local)
            fp64_src = "measured in place: tno_device_peak (DFMA chains, %.1f TFLOP/s; FP64 tensor-core mma.sync m8n8k4: %.1f TFLOP/s)" % (
                fp64_peak, _tl.device_peak("fp64_mma", local))
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": roof_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel_share_of_step": share,
                    "bytes_per_eval": bytes_per_eval, "evals_per_launch": evals_per_launch,
                    "kernel_durations_from": profile_pass,
                    # secondary bound (north_star: "fraction of the HBM or FP64 roofline"): algorithmic flops of SURVEY 8d
                    "fp64": (lambda fl: {"flops_per_eval": fl, "achieved_tflops": value / world * fl / 1e12, "peak_tflops": fp64_peak,
                                         "peak_source": fp64_src, "frac": value / world * fl / 1e12 / fp64_peak})(
                        int(info["factor_flops"] + info["solve_flops"] + 6 * plan.m * plan.k + 8 * plan.m)),
                    "whole_step": {"bytes_per_eval": B_alg, "achieved": value / world * B_alg / 1e9,
                                   "frac": value / world * B_alg / 1e9 / peak},
                    "kernels": per_kernel}

    # ---- end to end through the host-buffer API: pinned x -> device, evaluate, every output -> pinned host ----
    e2e = None
    Ne = min(args.e2e_batch, N)
    xh = torch.from_numpy(X[:Ne].copy()).pin_memory()
    bufs = {}
    host_names = ("f", "grad", "g", "J", "z", "gmin", "status")
    for _ in range(2):
        plan.evaluate_batch(xh, to_host=host_names, on_device=(), buffers=bufs)
    barrier()
    t0 = time.perf_counter()
    esteps = max(2, args.steps // 2)
    for _ in range(esteps):
        res = plan.evaluate_batch(xh, to_host=host_names, on_device=(), buffers=bufs)
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    d2h = sum(res[k].numel() * res[k].element_size() for k in host_names)
    # host-filled share of the constant funicular rows: the split tno_fetch_jacobian_host computes from its thread count
    _thr = max(1, min(int(os.environ.get("TNO_HOST_THREADS", "16")), os.cpu_count() or 1, Ne // 64 + 1))
    _cst, _per = 2 * plan.m * plan.nvar if "funicular" in plan.constraints else 0, plan.ng * plan.nvar
    _phi = float(os.environ.get("TNO_HOST_FILL_FRACTION", min(1.0, _per / (_cst * (1.0 + 52.0 / (4.5 * _thr)))) if _cst else 0.0))
    _fs = plan.fetch_stats   # the split the library actually used in the last fetch (it follows the measured host / DMA rates)
    if _cst and "TNO_HOST_FILL_FRACTION" not in os.environ and _fs["host_threads"] > 0:
        _phi, _thr = _fs["host_fill_fraction"], _fs["host_threads"]
    _rows = min(2 * plan.m, int(_phi * 2 * plan.m + 0.5)) if _cst else 0
    _rows -= (_rows * plan.nvar) & 1
    host_filled = Ne * 8 * _rows * plan.nvar
    e2e = {"value": world * Ne * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": int(xh.numel() * 8), "d2h_bytes_per_step": int(d2h),
           "d2h_bytes_over_pcie_per_step": int(d2h - host_filled), "host_filled_bytes_per_step": int(host_filled),
           "candidates_per_step": Ne, "numa": numa, "host_fill_threads": _thr, "host_fill_fraction_of_constant_rows": _phi,
           "note": "every output incl. the dense J delivered to pinned host memory; d2h_bytes_per_step = delivered bytes, of which the "
                   "candidate-independent funicular rows of J (host_filled_bytes_per_step) are written by host threads from the plan's "
                   "copy and only d2h_bytes_over_pcie_per_step cross PCIe; the split follows the measured host-fill and DMA rates "
                   "(tno_plan_fetch_stats)"}
    e2e["delivered_gb_per_s"] = e2e["value"] * d2h / Ne / 1e9
    e2e["measured_host_fill_gb_per_s"], e2e["measured_dma_gb_per_s"] = _fs["host_fill_gb_per_s"], _fs["dma_gb_per_s"]
    if rank == 0:
        e2e["host_dram_copy_gb_per_s"] = host_copy_bandwidth()
    # the sweep use-case: only objectives + feasibility cross PCIe, g / J / z stay in HBM for an on-device consumer
    bufs2 = {}
    xh2 = torch.from_numpy(X).pin_memory()
    for _ in range(2):
        plan.evaluate_batch(xh2, buffers=bufs2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        plan.evaluate_batch(xh2, buffers=bufs2)
    barrier()
    dt2 = time.perf_counter() - t0
    e2e["objectives_only"] = {"value": world * N * esteps / dt2, "unit": UNIT, "h2d_bytes_per_step": int(xh2.numel() * 8),
                              "d2h_bytes_per_step": int(N * (8 + 8 + 4)), "note": "f, gmin, status to host; grad, g, J, z materialised in HBM"}

    # feasibility sweeps (Monte-Carlo, multi-start screening) need neither J nor the gradient: the dz/dq_ind solve is skipped
    bufs3 = {}
    for _ in range(2):
        plan.evaluate_batch(xh2, to_host=("f", "gmin", "status"), on_device=(), buffers=bufs3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        plan.evaluate_batch(xh2, to_host=("f", "gmin", "status"), on_device=(), buffers=bufs3)
    barrier()
    dt3 = time.perf_counter() - t0
    e2e["screening_only"] = {"value": world * N * esteps / dt3, "unit": "candidates/s", "h2d_bytes_per_step": int(xh2.numel() * 8),
                             "d2h_bytes_per_step": int(N * (8 + 8 + 4)),
                             "note": "NOT the headline metric: f, gmin, status only; no J, no gradient, sensitivity solve skipped"}
    del bufs2, bufs3

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        ns = max(cores * args.cpu_evals_per_core, cores)
        SP = strip_problem(P)
        cpu_evals_per_s(SP, X[:cores], objective, cores)  # warm the pool / caches
        v = cpu_evals_per_s(SP, X[:ns], objective, cores)
        v1 = cpu_evals_per_s(SP, X[:max(4, args.cpu_evals_per_core // 2)], objective, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "single_core_value": v1,
               "sample": "%d candidates of %s through oracle/callbacks_np.py (reference call order: constr_wrapper, "
                         "sensitivities_wrapper, objective, gradient), one process per core" % (ns, args.workload)}

    # ---- sweep wall-times of the other BASELINE configurations, sharded over the ranks (every world size) ----
    sweeps = None
    if not args.no_sweeps:
        sweeps = run_sweeps(dev, local, world, rank, barrier)
        for extra in (run_multistart, lambda *a: {"crossvault_disc20_geometry_mc": run_geometry_montecarlo(*a)}):
            try:
                sweeps.update(extra(dev, local, world, rank, barrier))
            except Exception as exc:    # a failed extra sweep must not lose the headline line; it is reported instead
                if world > 1:
                    raise
                sweeps["error_%d" % len(sweeps)] = "%s: %s" % (type(exc).__name__, exc)

    # ---- batch-size dependence (N in {1, 256, 4096, 65536}), device-resident, all outputs ----
    batch_sizes = None
    if rank == 0 and world == 1 and not args.no_sweeps:
        batch_sizes = {}
        for nb_ in (1, 256, 4096, 65536):
            if nb_ * B_alg > 60e9:     # keep the output buffers well inside HBM
                continue
            Xn = torch.from_numpy(candidates(P, xs, nb_, seed=20261019 + 100)).to(dev)
            outn = {k: torch.empty(plan._shape(k, nb_), device=dev, dtype=torch.int32 if k == "status" else torch.float64) for k in names}
            reps = 200 if nb_ == 1 else (50 if nb_ <= 4096 else 5)
            for _ in range(3):
                plan.evaluate_device(Xn, names, out=outn)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(reps):
                plan.evaluate_device(Xn, names, out=outn)
            b1.record()
            torch.cuda.synchronize()
            msn = b0.elapsed_time(b1) / reps
            batch_sizes[str(nb_)] = {"ms_per_call": msn, "evals_per_s": nb_ / (msn * 1e-3)}
            del outn, Xn
            torch.cuda.empty_cache()

    # ---- configuration 1: the reference's own CPU-runnable case, SLSQP on the disc-10 cross vault ----
    slsqp = None
    if rank == 0 and world == 1 and not args.no_sweeps:
        slsqp = slsqp_config1(local)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "candidates_per_gpu_per_step": N, "objective": objective,
                       "variables": plan.variables, "constraints": plan.constraints, "n": plan.n, "m": plan.m, "k": plan.k,
                       "nvar": plan.nvar, "ng": plan.ng, "nnz_l": info["nnz_l"], "etree_height": info["etree_height"],
                       "kernel_family": {1: "interleaved", 2: "onchip"}.get(info.get("last_kernel") or info["kernel"], str(info["kernel"])),
                       "l2_policy": "outputs per step (%.2f GB) exceed the 126 MB L2; inputs re-read each step" % (N * B_alg / 1e9),
                       "parallelism": "candidates sharded, plan replicated, all_gather(f, status)" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "sweeps": sweeps,
            "best_feasible_of_last_step": (lambda b: {"f": b[0] if b[2] else None, "index": b[1], "feasible": b[2]})(
                best_of_packed(packed_all if packed_all is not None else gather_packed(out["f"], out["gmin"], out["status"], N), tol=1e-9)),
            "batch_sizes": batch_sizes, "slsqp_config1": slsqp,
            "status_nonzero": status_bad,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def host_copy_bandwidth(threads=16, mib=128, reps=3):
    """Host DRAM copy bandwidth (read + written GB/s) of this box with `threads` threads: the yardstick of the full-J host
    path, whose delivered bytes all land in host DRAM (numpy copies release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    threads = max(1, min(threads, os.cpu_count() or 1))
    n = mib * (1 << 20) // 8
    src = [np.ones(n) for _ in range(threads)]
    dst = [np.empty(n) for _ in range(threads)]
    best = 0.0
    with ThreadPoolExecutor(max_workers=threads) as pool:
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            list(pool.map(lambda i: np.copyto(dst[i], src[i]), range(threads)))
            best = max(best, 2.0 * threads * n * 8 / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="crossvault_disc20")
    ap.add_argument("--batch", type=int, default=16384, help="candidates per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=2048)
    ap.add_argument("--kernel", default="auto", choices=["auto", "interleaved", "onchip"])
    ap.add_argument("--cpu-evals-per-core", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweeps", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
