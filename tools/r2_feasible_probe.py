"""Probe: strictly feasible centres of the bench workloads (workloads.max_margin_state) and the feasible fraction of
the candidate ladders around them."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import build_workload
from compas_tno_b200 import workloads as W
from compas_tno_b200.plan import TnoPlan
for wl in ("crossvault_disc10", "crossvault_disc20", "dome_16x20", "crossvault_disc40"):
    P, xs, obj = build_workload(wl, n_states=6)
    plan = TnoPlan(P, objective=obj)
    t = time.time()
    x, s = W.max_margin_state(plan, xs, rows=W.margin_rows(P, plan))
    dt = time.time() - t
    X = W.sample_around(P, x, 4096, seed=1)
    r = plan.evaluate_host(X, ("f", "gmin", "status"))
    feas = (r["gmin"] >= -1e-9) & (r["status"] == 0)
    print("%s: margin %.4g in %.1f s; ladder feasible %d / 4096; f range feasible [%.4g, %.4g]" % (
        wl, s, dt, feas.sum(), r["f"][feas].min() if feas.any() else np.nan, r["f"][feas].max() if feas.any() else np.nan), flush=True)
    plan.close()
