import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from golden_io import load_case
from compas_tno_b200.plan import TnoPlan
from compas_tno_b200 import multistart as ms

case = sys.argv[1] if len(sys.argv) > 1 else "cross6_q_zb_min"
P, objective, xs, _ = load_case(case)
lb, ub = ms.variable_bounds(P)
rng = np.random.default_rng(11)
X0 = xs[0] * (1.0 + 0.10 * rng.uniform(-1, 1, size=(96, xs.shape[1])))
X0[0] = xs[0]
plan = TnoPlan(P, objective=objective)
sol = ms.MultiStartSolver(plan, lb, ub, acc=1e-8, max_iter=200)
sol.debug = True
res = sol.solve(X0)
print("success", res.success.sum(), "of", len(res.success), "iters", res.n_iterations, "f range", res.f[res.success].min(), res.f[res.success].max())
