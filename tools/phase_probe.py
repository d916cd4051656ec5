import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import build_workload, candidates
from compas_tno_b200.plan import TnoPlan
P, xs, obj = build_workload(sys.argv[1] if len(sys.argv)>1 else "crossvault_disc20")
plan = TnoPlan(P, objective=obj)
print(plan.info)
for N in (1, 296, 4096):
    X = candidates(P, xs, N, 1)
    xd = torch.from_numpy(X).cuda()
    for _ in range(2):
        r = plan.evaluate_device(xd); torch.cuda.synchronize()
