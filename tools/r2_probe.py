"""Round-2 probe: k_onchip timing of the headline workload (one process per variant: the knobs are read once).
usage: python tools/r2_probe.py [workload] [N]   -- prints ms per call, evals/s and, with TNO_PHASE_CLOCKS=1, the phase clocks."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import build_workload, candidates  # noqa: E402
from compas_tno_b200.plan import TnoPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "crossvault_disc20"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
kernel = os.environ.get("PROBE_KERNEL", "auto")
P, xs, obj = build_workload(wl, n_states=6)
plan = TnoPlan(P, objective=obj, kernel=kernel)
names = ("f", "grad", "g", "J", "z", "gmin", "status")
X = candidates(P, xs, N, 1)
xd = torch.from_numpy(X).cuda()
out = {k: torch.empty(plan._shape(k, N), device="cuda", dtype=torch.int32 if k == "status" else torch.float64) for k in names}
for _ in range(3):
    plan.evaluate_device(xd, names, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
a.record()
for _ in range(reps):
    plan.evaluate_device(xd, names, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
info = plan.info
print("PROBE %s N=%d kernel=%s last=%d ctas/sm=%d smem=%d env={%s}: %.3f ms/call  %.3f M evals/s  bad=%d  fsum=%.12e" % (
    wl, N, kernel, info["last_kernel"], info["onchip_ctas_per_sm"], info["smem_bytes"],
    ",".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("TNO_")), ms, N / ms / 1e3,
    int((out["status"] != 0).sum()), float(out["f"].sum())), flush=True)
