import sys, os, torch
sys.path.insert(0, os.getcwd())
from bench import build_workload, candidates
from compas_tno_b200 import workloads as W
from compas_tno_b200.plan import TnoPlan
names = ("f", "grad", "g", "J", "z", "gmin", "status")
def run(P, X, obj, kernel, load_dir=None, reps=5):
    plan = TnoPlan(P, objective=obj, kernel=kernel)
    xd = torch.from_numpy(X).cuda()
    ld = torch.from_numpy(load_dir).cuda() if load_dir is not None else None
    out = {k: torch.empty(plan._shape(k, len(X)), device="cuda", dtype=torch.int32 if k == "status" else torch.float64) for k in names}
    for _ in range(2): plan.evaluate_device(xd, names, out=out, load_dir=ld)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): plan.evaluate_device(xd, names, out=out, load_dir=ld)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    plan.close(); del out
    torch.cuda.empty_cache()
    return ms
P, xs, obj = build_workload("dome_16x20", n_states=6)
for N in (1024, 4096, 8192, 16384):
    X = candidates(P, xs, N, seed=3)
    print("dome N=%d  onchip %.3f ms  interleaved %.3f ms" % (N, run(P, X, obj, "onchip"), run(P, X, obj, "interleaved")), flush=True)
Pd, xd_ = W.dome_direction_problem(20, 16, n_states=6)
X, ld = W.direction_candidates(Pd, xd_, 8192, seed=5)
print("dome dir N=8192 onchip %.3f ms interleaved %.3f ms" % (run(Pd, X, "max_load", "onchip", ld), run(Pd, X, "max_load", "interleaved", ld)), flush=True)
