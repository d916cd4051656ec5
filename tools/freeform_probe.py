"""Throughput of the free-form fixture (k = 360 independents, 30 supports: nvar = 390, J = 2790 x 390) on the interleaved family."""
import os, sys, numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import load_case
from compas_tno_b200.plan import TnoPlan
P, objective, xs, exp = load_case("freeform_q_zb_min")
plan = TnoPlan(P, objective=objective)
print(plan.info)
names = ("f", "grad", "g", "J", "z", "gmin", "status")
for N in (64, 512):
    rng = np.random.default_rng(1)
    X = xs[0] * (1.0 + 0.02 * rng.uniform(-1, 1, size=(N, xs.shape[1])))
    xd = torch.from_numpy(X).cuda()
    out = {k: torch.empty(plan._shape(k, N), device="cuda", dtype=torch.int32 if k == "status" else torch.float64) for k in names}
    for _ in range(2): plan.evaluate_device(xd, names, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): plan.evaluate_device(xd, names, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    plan.profile(True)
    plan.evaluate_device(xd, names, out=out); torch.cuda.synchronize()
    pr = plan.profile_read(); plan.profile(False)
    print("   " + "  ".join("%s %.2f(%d)" % (k, v[0], v[1]) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0])))
    print("N=%d  %.2f ms  %.0f evals/s  bytes/eval %.2f MB -> %.0f GB/s of outputs, status!=0: %d" % (N, ms, N / ms * 1e3, plan.bytes_per_eval / 1e6, N * plan.bytes_per_eval / ms / 1e6, int((out["status"] != 0).sum())))
    del out
