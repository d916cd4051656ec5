"""Fill-reducing ordering (minimum degree vs geometric nested dissection) against step time, interleaved family."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import build_workload, candidates
from compas_tno_b200.plan import TnoPlan
wl = sys.argv[1] if len(sys.argv) > 1 else "crossvault_disc40"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
P, xs, obj = build_workload(wl, n_states=6)
X = torch.from_numpy(candidates(P, xs, N, seed=3)).cuda()
names = ("f", "grad", "g", "J", "z", "gmin", "status")
for kernel in ("interleaved", "auto"):
    for ordering in ("mindeg", "nested"):
        plan = TnoPlan(P, objective=obj, kernel=kernel, ordering=ordering)
        out = {k: torch.empty(plan._shape(k, N), device="cuda", dtype=torch.int32 if k == "status" else torch.float64) for k in names}
        for _ in range(2): plan.evaluate_device(X, names, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): plan.evaluate_device(X, names, out=out)
        b.record(); torch.cuda.synchronize()
        i = plan.info
        print("%s %s %s: nnz_l %d height %d  %.2f ms  %.0f evals/s  family %d" % (wl, kernel, ordering, i["nnz_l"], i["etree_height"], a.elapsed_time(b) / 3, N / a.elapsed_time(b) * 3e3, i["last_kernel"]), flush=True)
        plan.close(); del out; torch.cuda.empty_cache()
