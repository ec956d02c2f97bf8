"""G1-29dof (30 bodies, nv = 35, c = 358) on the large-model path: statistics rate, evaluation pass and the LMI fit.
    python tools/g1_29dof_rate.py [samples]   ->  one JSON line (copy to profiles/)"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
from system_identification_b200.sys_identification import SystemIdentification

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", "g1_29dof.json"))
si = SystemIdentification.from_flat_model(flat)
dm = si.device_model
q, dq, ddq, cnt = synth.make_trajectory(flat, N, 7)
tau = synth.synth_tau(flat, N, 3, scale=5.0)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


ms_gram, st = timed(lambda: dm.gram_accumulate(*dev))
L, nd = 30, 29
ms_sdp, (x, info) = timed(lambda: ops.sdp_solve(st, L, nd, flat.phi_prior, flat.ellipsoids, flat.robot_mass))
t0 = time.perf_counter(); phi = si.identify(q, dq, ddq, tau, cnt); t_id = time.perf_counter() - t0
t0 = time.perf_counter(); phi = si.identify(q, dq, ddq, tau, cnt); t_id = min(t_id, time.perf_counter() - t0)
print(json.dumps({"robot": "g1_29dof", "samples": N, "c": 358, "gram_ms": ms_gram, "gram_samples_per_s": N / ms_gram * 1e3,
                  "algorithmic_tflops": N * 4523330 / ms_gram / 1e9, "sdp_ms": ms_sdp, "sdp_newton_steps": int(info[0]["iterations"]),
                  "sdp_status": int(info[0]["status"]), "identify_host_arrays_s": t_id}))
