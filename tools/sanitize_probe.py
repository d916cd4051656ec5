"""Small run of every objective / variable combination through both kernel families (a quick end-to-end probe; also the command to put under compute-sanitizer where that tool is available)."""
import os, sys, numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import load_case, rel_err
from compas_tno_b200.plan import TnoPlan
cases = sys.argv[1:] or ["cross6_q_zb_min", "cross14_q_zb_bestfit", "cross14_q_zb_loadpath", "cross14_q_zb_ecomp_nonlinear",
                         "cross14_q_zb_t", "dome_q_zb_lambdh_reac", "cross14_q_xyb_zb_min"]
for name in cases:
    P, objective, xs, exp = load_case(name)
    for kernel in ("onchip", "interleaved"):
        try:
            plan = TnoPlan(P, objective=objective, kernel=kernel)
        except NotImplementedError:
            continue
        got = plan.evaluate_host(np.repeat(xs, 2, axis=0)[:5])
        print(name, kernel, "J err %.1e" % rel_err(got["J"][0], exp["J"][0]), flush=True)
        plan.close()
