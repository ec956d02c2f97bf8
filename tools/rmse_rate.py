"""Throughput of the evaluation pass (print_tau_prediction_rmse, sysid_predict_rmse) on one B200."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = {}
for robot in ("g1_12dof", "solo12"):
    flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", robot + ".json"))
    dm = ops.DeviceModel(flat)
    q, dq, ddq, cnt = synth.make_trajectory(flat, N, 7)
    tau = synth.synth_tau(flat, N, 3, scale=10.0)
    dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
    phi = torch.as_tensor(np.asarray(flat.phi_prior, dtype=np.float64))
    ms = []
    for it in range(6):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = dm.predict_rmse(*dev, phi); e1.record(); torch.cuda.synchronize()
        if it >= 2: ms.append(e0.elapsed_time(e1))
    out[robot] = {"ms": float(np.mean(ms)), "Msamples_per_s": N / float(np.mean(ms)) / 1e3, "total": float(r[0])}
print(json.dumps({"workload": f"predict_rmse over {N} samples, one B200", **out}))
