"""Batched multi-start driver: S optimisations of ONE problem from S starting points advance
together on the GPU.

The reference runs one optimisation at a time through ``scipy.optimize.fmin_slsqp`` (``solvers/scipy.py:127-174``:
``fn, x0, bounds, fprime, f_ieqcons, fprime_ieqcons``), calling the four callbacks once per iteration on the host.  Here
every iteration is, for all starts at once,

1. one ``tno_eval_batch`` (f, grad f, g, dense J stay in HBM),
2. ``tno_sqp_qp_solve``: the QP subproblem of every start (interior point on the dense J, one CTA per start),
3. a line search on the L1 merit function f + rho * sum max(0, -g) with batched f / g evaluations,
4. ``tno_sqp_bfgs_update``: damped BFGS update of every start's Hessian approximation,

and nothing but a few scalars per start crosses PCIe.  The stopping rule is SLSQP's: ``|grad.d| + sum lam |g| < acc`` with
the constraint violation below ``acc``.  ``variable_bounds`` builds the per-variable bounds like the reference's
``_build_variable_vector`` (``problems/setup.py:168-293``).

torch is used for device buffers, streams and trivially parallel element-wise glue only; the QP, the BFGS update and the
evaluations are the library's CUDA kernels (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .plan import TnoPlan


def variable_bounds(problem, thk_bounds=(0.001, None), max_lambd=None):
    """(lb, ub) of x = [q_ind | zb | t | lambdh | lambdv] as problems/setup.py:168-293 builds them: q_ind in [qmin, qmax],
    zb in [lb, ub] of the support vertices, t in [min_thk, max_thk = thk], lambdh in [0, 10], lambdv in [0, 100]."""
    variables = list(problem.variables or [])
    qmin = np.broadcast_to(np.asarray(problem.qmin, dtype=np.float64).reshape(-1), (problem.m,))
    qmax = np.broadcast_to(np.asarray(problem.qmax, dtype=np.float64).reshape(-1), (problem.m,))
    ind = np.asarray(problem.ind, dtype=int)
    lo, hi = [qmin[ind]], [qmax[ind]]
    if "xyb" in variables:
        raise NotImplementedError("multi-start driver: xyb bounds live on the form diagram (xmin / xmax attributes)")
    if "zb" in variables:
        fixed = np.asarray(problem.fixed, dtype=int)
        lo.append(np.asarray(problem.lb, dtype=np.float64).reshape(-1)[fixed])
        hi.append(np.asarray(problem.ub, dtype=np.float64).reshape(-1)[fixed])
    if "t" in variables or "n" in variables:
        thk = float(np.asarray(problem.thk).reshape(-1)[0])
        if "t" in variables:
            lo.append([thk_bounds[0]])
            hi.append([thk if thk_bounds[1] is None else thk_bounds[1]])
        else:
            lo.append([0.0])
            hi.append([thk / 2.0])
    if "lambdh" in variables:
        lo.append([0.0])
        hi.append([10.0 if max_lambd is None else max_lambd])
    if "lambdv" in variables:
        lo.append([0.0])
        hi.append([100.0 if max_lambd is None else max_lambd])
    return np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in lo]), \
        np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in hi])


@dataclass
class MultiStartResult:
    """Per-start results (host arrays) and the best feasible start."""
    x: np.ndarray          # S x nvar
    f: np.ndarray          # S
    gmin: np.ndarray       # S     min_r g[r] at x
    niter: np.ndarray      # S     SQP iterations taken
    success: np.ndarray    # S     bool: SLSQP's stopping rule met
    status: np.ndarray     # S     TNO_STATUS_* of the last evaluation
    best: int              # index of the best converged feasible start, -1 if none
    n_evals: int           # candidate evaluations spent (all starts, line searches included)
    n_iterations: int      # iterations of the batch loop
    qp_iterations: float   # mean interior point iterations per QP

    @property
    def xopt(self):
        return None if self.best < 0 else self.x[self.best].reshape(-1, 1)

    @property
    def fopt(self):
        return None if self.best < 0 else float(self.f[self.best])


class MultiStartSolver:
    """S simultaneous SQP runs on one plan.  ``solve(X0)`` returns a :class:`MultiStartResult`."""

    def __init__(self, plan: TnoPlan, lb, ub, acc: float = 1e-6, max_iter: int = 100, qp_iter: int = 40, qp_tol: float = 1e-10,
                 max_backtracks: int = 10, use_mma: bool = False, feas_tol: float = 1e-8, patience: int = 20):
        import torch
        self.plan = plan
        self.torch = torch
        self.dev = torch.device("cuda", plan.device)
        self.lib = _lib.load()
        self.acc, self.max_iter, self.qp_iter, self.qp_tol = float(acc), int(max_iter), int(qp_iter), float(qp_tol)
        self.max_backtracks, self.use_mma, self.feas_tol = int(max_backtracks), bool(use_mma), float(feas_tol)
        self.patience = int(patience)   # iterations without progress after which the last <= 0.5 % of the starts are given up
        nv = plan.nvar
        if nv > 64:
            raise NotImplementedError("multi-start driver: more than 64 variables per start")
        lb = np.asarray(lb, dtype=np.float64).reshape(-1)
        ub = np.asarray(ub, dtype=np.float64).reshape(-1)
        if lb.shape != (nv,) or ub.shape != (nv,):
            raise ValueError("bounds must have %d entries" % nv)
        self.lb = torch.from_numpy(np.where(np.isfinite(lb), lb, -1e20)).to(self.dev)
        self.ub = torch.from_numpy(np.where(np.isfinite(ub), ub, 1e20)).to(self.dev)
        # the funicular rows come first and do not depend on the candidate: every start's QP reads ONE copy of them
        self.nshared = 2 * plan.m if "funicular" in plan.constraints else 0
        self._buf = {}

    # ------------------------------------------------------------------------------------------
    def _alloc(self, S):
        t, nv, ng = self.torch, self.plan.nvar, self.plan.ng
        if self._buf.get("S") == S:
            return self._buf
        f64 = dict(dtype=t.float64, device=self.dev)
        mt = ng + 2 * nv
        B = {"S": S}
        for name, shape in (("f", (S,)), ("grad", (S, nv)), ("g", (S, ng)), ("J", (S, ng, nv)), ("gmin", (S,))):
            B[name] = t.empty(shape, **f64)
        B["status"] = t.empty((S,), dtype=t.int32, device=self.dev)
        B["ft"], B["gt"], B["gmint"] = t.empty((S,), **f64), t.empty((S, ng), **f64), t.empty((S,), **f64)
        B["statust"] = t.empty((S,), dtype=t.int32, device=self.dev)
        B["H"] = t.empty((S, nv, nv), **f64)
        B["d"], B["lam"] = t.empty((S, nv), **f64), t.empty((S, mt), **f64)
        B["info"], B["vold"] = t.empty((S, 8), **f64), t.empty((S, nv), **f64)
        B["work"] = t.empty((S, 4, mt), **f64)
        B["step"] = t.empty((S, nv), **f64)
        self._buf = B
        return B

    def _eval(self, x, full, B):
        if full:
            out = {k: B[k] for k in ("f", "grad", "g", "J", "gmin", "status")}
            self.plan.evaluate_device(x, tuple(out), out=out)
        else:
            out = {"f": B["ft"], "g": B["gt"], "gmin": B["gmint"], "status": B["statust"]}
            self.plan.evaluate_device(x, tuple(out), out=out)

    def _ptr(self, tns):
        return C.c_void_p(tns.data_ptr()) if tns is not None else None

    # ------------------------------------------------------------------------------------------
    def solve(self, X0) -> MultiStartResult:
        t = self.torch
        nv, ng = self.plan.nvar, self.plan.ng
        X0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64).reshape(-1, nv))
        S = X0.shape[0]
        B = self._alloc(S)
        x = t.from_numpy(X0).to(self.dev)
        x = t.minimum(t.maximum(x, self.lb), self.ub).contiguous()
        stream = C.c_void_p(t.cuda.current_stream(self.dev).cuda_stream)
        B["H"].copy_(t.eye(nv, dtype=t.float64, device=self.dev).expand(S, nv, nv))
        self._eval(x, True, B)
        n_evals = S
        Jsh = B["J"][0, :self.nshared].contiguous() if self.nshared else None
        active = t.ones(S, dtype=t.uint8, device=self.dev)
        success = t.zeros(S, dtype=t.bool, device=self.dev)
        niter = t.zeros(S, dtype=t.int32, device=self.dev)
        rho = t.zeros(S, dtype=t.float64, device=self.dev)
        qp_its, qp_cnt, it = 0.0, 0, 0
        stall_count, stall_iters = -1, 0
        for it in range(1, self.max_iter + 1):
            _lib.check(self.lib.tno_sqp_qp_solve(
                S, nv, ng, self.nshared, self._ptr(B["H"]), self._ptr(B["grad"]), self._ptr(B["g"]), self._ptr(B["J"]),
                self._ptr(Jsh), self._ptr(x), self._ptr(self.lb), self._ptr(self.ub), self._ptr(active), self._ptr(B["d"]),
                self._ptr(B["lam"]), self._ptr(B["info"]), self._ptr(B["vold"]), self._ptr(B["work"]), self.qp_iter, self.qp_tol,
                1 if self.use_mma else 0, stream))
            info, act = B["info"], active.bool()
            gd, lamb, viol, lmax = info[:, 0], info[:, 1], info[:, 2], info[:, 3]
            bad = ~t.isfinite(B["d"]).all(dim=1) | (B["status"] != 0)
            # SLSQP's stopping rule (scipy slsqp: |g.d| + sum |lam| |c| < acc with the violation below acc)
            done = act & ~bad & ((gd.abs() + lamb) < self.acc) & (viol < self.acc)
            success |= done
            qp_its += float((info[:, 4] * act).sum())
            qp_cnt += int(act.sum())
            if getattr(self, "debug", False):
                nf = ~t.isfinite(B["d"]).all(dim=1)
                print("[ms] it %3d active %d done %d nonfinite-d %d status!=0 %d qp-unconverged %d  max|d| %.3e viol %.3e" % (
                    it, int(act.sum()), int(done.sum()), int((act & nf).sum()), int((act & (B["status"] != 0)).sum()),
                    int((act & (info[:, 5] == 0)).sum()), float(B["d"][act & ~nf].abs().max()) if bool((act & ~nf).any()) else 0.0,
                    float(viol[act].max())), flush=True)
            act = act & ~done & ~bad
            n_act = int(act.sum())
            if n_act == 0:
                active = act.to(t.uint8)
                break
            # stragglers: the starts advance in lock step, so a handful that do not converge would keep the whole batch
            # iterating to max_iter.  Once at most 0.5 % of the starts (at least one) are left and none of them has
            # finished for `patience` iterations, they are given up (reported as not converged).
            if n_act <= max(1, S // 200) and n_act == stall_count:
                stall_iters += 1
                if stall_iters >= self.patience:
                    active = act.to(t.uint8)
                    break
            else:
                stall_count, stall_iters = n_act, 0
            niter += act.to(t.int32)
            # penalty of the L1 merit function: above the largest multiplier, never decreased by more than half a step
            rho = t.where(act, t.maximum(0.5 * (rho + lmax), 1.01 * lmax), rho)
            phi0 = B["f"] + rho * viol
            dphi = gd - rho * viol
            alpha = t.where(act, 1.0, 0.0).to(t.float64)
            accepted = ~act
            xt = x.clone()
            for ls in range(self.max_backtracks + 1):
                xt = t.where(accepted.unsqueeze(1), xt, x + alpha.unsqueeze(1) * B["d"]).contiguous()
                self._eval(xt, False, B)
                n_evals += int((~accepted).sum())
                phit = B["ft"] + rho * t.clamp(-B["gt"], min=0.0).sum(dim=1)
                ok = t.isfinite(phit) & (B["statust"] == 0) & (phit <= phi0 + 0.1 * alpha * dphi + 1e-14 * phi0.abs())
                accepted = accepted | ok
                if bool(accepted.all()) or ls == self.max_backtracks:
                    break
                alpha = t.where(accepted, alpha, alpha * (0.3 if ls == 0 else 0.5))
            # starts whose line search failed take the shortest trial step (the merit function may be badly scaled there)
            B["step"].copy_(t.where(act.unsqueeze(1), xt - x, t.zeros_like(x)))
            x = t.where(act.unsqueeze(1), xt, x).contiguous()
            self._eval(x, True, B)
            n_evals += int(act.sum())
            active = act.to(t.uint8)
            _lib.check(self.lib.tno_sqp_bfgs_update(
                S, nv, ng, self.nshared, self._ptr(B["H"]), self._ptr(B["grad"]), self._ptr(B["J"]), self._ptr(Jsh),
                self._ptr(B["lam"]), self._ptr(B["vold"]), self._ptr(B["step"]), self._ptr(active), stream))
        t.cuda.synchronize(self.dev)
        f, gmin = B["f"].cpu().numpy().copy(), B["gmin"].cpu().numpy().copy()
        status = B["status"].cpu().numpy().copy()
        succ = success.cpu().numpy()
        feas = succ & (gmin >= -max(self.feas_tol, self.acc)) & (status == 0)
        best = int(np.argmin(np.where(feas, f, np.inf))) if feas.any() else -1
        return MultiStartResult(x=x.cpu().numpy(), f=f, gmin=gmin, niter=niter.cpu().numpy(), success=succ, status=status,
                                best=best, n_evals=n_evals, n_iterations=it, qp_iterations=qp_its / max(qp_cnt, 1))


def solve_multistart(problem, X0, objective="min", bounds=None, device=0, plan: Optional[TnoPlan] = None, **kw):
    """One call: plan (unless given), bounds from the problem (unless given), S simultaneous optimisations from the rows of
    ``X0``.  The counterpart of ``ScipySolver.solve`` (solvers/scipy.py:58-125) for a batch of starting points."""
    plan = plan if plan is not None else TnoPlan(problem, objective=objective, device=device)
    lb, ub = bounds if bounds is not None else variable_bounds(problem)
    return MultiStartSolver(plan, lb, ub, **kw).solve(X0)
