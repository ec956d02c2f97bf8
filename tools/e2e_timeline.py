"""Where identify()'s time goes on the headline workload (1 M-sample G1-12dof log in pinned host memory): CUDA-event / wall
times of the pieces, best of a few repetitions.  Diagnostic -- bench.py's e2e leg is the reported figure."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from system_identification_b200 import ops
from system_identification_b200.identify import identify, _plan_for
from system_identification_b200.sys_identification import SystemIdentification

N = int(os.environ.get("E2E_SAMPLES", 1_000_000))
flat = bench.load_flat()
si = SystemIdentification.from_flat_model(flat)
dm = si.device_model
q, dq, ddq, tau, cnt = bench.host_log(flat, N)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
dev[3] = bench.identifiable_tau(flat, dm, dev, seed=17)
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (q, dq, ddq, tau, cnt)]
pinned[3] = dev[3].cpu().pin_memory()
plan = _plan_for(si, 13, 12, 1e-1, 1e-10, 1000, "constant_pullback")


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best, out

res = {}
res["gram_device_resident_ms"], st = timed(lambda: dm.gram_accumulate(*dev))
res["gram_host_streamed_ms"], _ = timed(lambda: dm.gram_accumulate_host(*pinned))
res["gram_host_streamed_presolve_ms"], _ = timed(lambda: dm.gram_accumulate_host(*pinned, presolve=plan, presolve_samples=131072))
for ps in (32768, 65536, 262144, 524288):
    res["gram_host_streamed_presolve_%d_ms" % ps], _ = timed(lambda: dm.gram_accumulate_host(*pinned, presolve=plan, presolve_samples=ps, chunk=max(131072, ps)))
res["solve_cold_ms"], (x, info) = timed(lambda: plan.solve(st))
res["solve_cold_newton"] = int(info[0]["iterations"])
dm.gram_accumulate_host(*pinned, presolve=plan, presolve_samples=131072)
res["solve_warm_ms"], (x, info) = timed(lambda: plan.solve(st, warm=plan.warm))
res["solve_warm_newton"] = int(info[0]["iterations"])
res["presolve_newton"] = int(plan.presolve_info()["iterations"])
w = plan.warm.cpu().numpy()
res["presolve_kernel_ms_under_load"] = (w[-1] - w[-2]) * 1e-6
res["identify_presolve_ms"], _ = timed(lambda: identify(si, *pinned, sharded=True))
res["identify_cold_ms"], _ = timed(lambda: identify(si, *pinned, sharded=True, presolve=False))
print(json.dumps(res, indent=1))
