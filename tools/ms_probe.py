"""Multi-start driver on a bench workload: starts scattered around the widest-margin state.
usage: python tools/ms_probe.py [workload] [S] [max_iter] [spread]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import build_workload  # noqa: E402
from compas_tno_b200 import workloads as W  # noqa: E402
from compas_tno_b200.multistart import MultiStartSolver, variable_bounds  # noqa: E402
from compas_tno_b200.plan import TnoPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "dome_16x20"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 100
spread = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
P, xs, obj = build_workload(wl, n_states=6)
plan = TnoPlan(P, objective=obj)
xc, margin = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
lb, ub = variable_bounds(P)
rng = np.random.default_rng(5)
X0 = xc[None, :] * (1.0 + spread * rng.uniform(-1, 1, size=(S, plan.nvar)))
X0[0] = xc
sol = MultiStartSolver(plan, lb, ub, acc=1e-6, max_iter=max_iter, use_mma=bool(int(os.environ.get("MS_MMA", "0"))))
sol.debug = bool(int(os.environ.get("MS_DEBUG", "0")))
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sol.solve(X0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = res.success & (res.gmin >= -1e-6)
    fo = res.f[ok]
    print("MS %s obj=%s S=%d nvar=%d ng=%d margin=%.3g: %.3f s  %.1f optimisations/s  converged+feasible %d  iterations %d  evals %d  "
          "qp-its %.1f  f best %s  f spread %s" % (wl, obj, S, plan.nvar, plan.ng, margin, dt, S / dt, int(ok.sum()), res.n_iterations,
                                                res.n_evals, res.qp_iterations, res.fopt,
                                                (float(fo.min()), float(np.median(fo)), float(fo.max())) if len(fo) else None), flush=True)
