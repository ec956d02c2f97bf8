"""identify() on the headline workload (1 M-sample G1-12dof log in pinned host memory) with the second pre-solve stage at
different fractions of the log: wall time (best of 5) and Newton steps of the final solve.  Diagnostic."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from system_identification_b200 import ops
from system_identification_b200 import identify as idm
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
res = {}
for tol, frac in ((None, 0.0), (None, 0.45), (None, 0.6), (1e-7, 0.45), (1e-6, 0.45), (1e-6, 0.55), (1e-6, 0.65), (1e-5, 0.45), (1e-5, 0.6), (1e-4, 0.5), (1e-4, 0.65)):
    idm.PRESOLVE_FIRST_TOL = tol
    idm.PRESOLVE_REFINE_FRACTION = frac
    idm.PRESOLVE_REFINE_MIN_LOG = 900_000 if frac > 0 else 10**12
    best, it = None, None
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = idm.identify(si, *pinned, sharded=True, return_info=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        if best is None or dt < best: best, it = dt, int(out[3]["iterations"])
    res["first_tol_%s_refine_%.2f" % (tol, frac)] = {"identify_ms": round(best, 3), "final_newton_steps": it}
print(json.dumps(res))
