"""Does a start's trajectory in the batched SQP driver depend on the batch it is in?  (round-2 diagnostic)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from bench import build_workload, _ladder
from compas_tno_b200 import workloads as W
from compas_tno_b200.multistart import MultiStartSolver, variable_bounds
from compas_tno_b200.plan import TnoPlan
P, xs, obj = build_workload("crossvault_disc10", n_states=6)
plan = TnoPlan(P, objective=obj, device=0)
xc, margin = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
lb, ub = variable_bounds(P)
X0 = _ladder(xc, 0, 4096, 20261019 + 60 + 1, 0.2 * 0.02, 0.02)
solver = MultiStartSolver(plan, lb, ub, acc=1e-6, max_iter=100)
for name, sl in (("all", slice(0, 4096)), ("first half", slice(0, 2048)), ("second half", slice(2048, 4096)), ("all again", slice(0, 4096))):
    r = solver.solve(X0[sl])
    bad = np.where(~r.success)[0]
    print(name, "iterations", r.n_iterations, "converged", int(r.success.sum()), "of", len(r.success), "unconverged idx", bad[:8] + sl.start,
          "niter max", int(r.niter.max()), flush=True)
    if len(bad):
        i = int(bad[0]) + sl.start
        r1 = solver.solve(X0[i:i + 1])
        print("   start", i, "alone: iterations", r1.n_iterations, "success", bool(r1.success[0]), "f", float(r1.f[0]), flush=True)
