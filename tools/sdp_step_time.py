"""Time of one Newton step of the LMI solve (sdp_alm_kernel, one CTA) on the headline problem: cold solve of the statistics of a
G1-12dof log (default 250 000 samples with ground-truth torques), CUDA events on the launching stream, five repeats after warm-up.
Prints one JSON line: ms per solve, Newton steps, us per step, the objective and |x| (to compare two builds of the solver:
SYSID_B200_LIB=<other .so> python tools/sdp_step_time.py).  Diagnostic -- never a bench number."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from system_identification_b200 import ops
from system_identification_b200.identify import _plan_for
from system_identification_b200.sys_identification import SystemIdentification
N = int(os.environ.get("SDP_STEP_SAMPLES", 250_000))
flat = bench.load_flat(); si = SystemIdentification.from_flat_model(flat); dm = si.device_model
q, dq, ddq, tau, cnt = bench.host_log(flat, N)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
dev[3] = bench.identifiable_tau(flat, dm, dev, seed=17)
st = dm.gram_accumulate(*dev)
plan = _plan_for(si, 13, 12, 1e-1, 1e-10, 1000, "constant_pullback")
for _ in range(2):
    x, info = plan.solve(st)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); x, info = plan.solve(st); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
it = int(info[0]["iterations"])
xs = x[0].cpu().numpy()
print(json.dumps({"lib": os.environ.get("SYSID_B200_LIB", "in-tree"), "samples": N, "solve_ms": [round(m, 3) for m in ms], "newton_steps": it,
                  "us_per_step": round(1e3 * min(ms) / it, 2), "status": int(info[0]["status"]),
                  "objective": float(info[0]["objective"]), "x_norm": float(np.linalg.norm(xs)), "x_head": [float(v) for v in xs[:4]]}))
