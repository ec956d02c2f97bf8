"""Newton-step counts of the LMI solve under a penalty schedule given by the environment (SYSID_SDP_SIGMA_GROWTH / _CAP / _THRESH,
read once per process by the library): cold solves of G1-12dof (250 000 samples, two torque seeds; 20 000 samples), Spot and Solo-12
problems (the warm-started chain of identify() is covered by bench.py's e2e leg, which reports its final solve).  The first run (no overrides, SWEEP_REF=write) stores its solutions;
later runs print their distance to them.  Diagnostic for the schedule's defaults in csrc/sdp_kernels.cuh."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from system_identification_b200 import ops, synth
from system_identification_b200.model import FlatModel
from system_identification_b200.identify import _plan_for
from system_identification_b200.sys_identification import SystemIdentification

REF = os.path.join(ROOT, "gpurun_out", "sigma_sweep_ref.npz")
out, sols = {}, {}


def timed(plan, st):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    x, info = plan.solve(st)
    e1.record(); torch.cuda.synchronize()
    return x[0].cpu().numpy(), int(info[0]["iterations"]), int(info[0]["refactorizations"]), int(info[0]["status"]), e0.elapsed_time(e1)


def problem(robot, N, seed, traj_seed, scale):
    flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", robot + ".json"))
    si = SystemIdentification.from_flat_model(flat); dm = si.device_model
    q, dq, ddq, cnt = synth.make_trajectory(flat, N, traj_seed)
    tau = synth.synth_tau(flat, N, 11, scale=scale)
    dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
    dev[3] = synth.identifiable_tau_device(flat, dm, dev, seed=seed)
    plan = _plan_for(si, flat.nbodies, flat.joints_dof, 1e-1, 1e-10, 1000, "constant_pullback")
    return dm, dev, plan


for robot, N, seed, tseed, scale in (("g1_12dof", 250_000, 17, synth.SEEDS["g1_1m"], 10.0), ("g1_12dof", 250_000, 18, synth.SEEDS["g1_1m"], 10.0),
                                     ("g1_12dof", 20_000, 23, synth.SEEDS["g1_12dof"], 10.0), ("spot", 20_000, 23, synth.SEEDS["spot"], 10.0),
                                     ("solo12", 20_000, 23, synth.SEEDS["solo12"], 1.0)):
    dm, dev, plan = problem(robot, N, seed, tseed, scale)
    st = dm.gram_accumulate(*dev)
    timed(plan, st)
    x, it, ref, status, ms = timed(plan, st)
    key = f"{robot}_{N}_s{seed}"
    out[key] = {"newton": it, "outer": ref, "status": status, "ms": round(ms, 3)}
    sols[key] = x

if os.environ.get("SWEEP_REF") == "write":
    np.savez(REF, **sols)
elif os.path.exists(REF):
    ref = np.load(REF)
    for k, x in sols.items():
        if k in ref:
            out[k]["rel_to_ref"] = float(np.linalg.norm(x - ref[k]) / np.linalg.norm(ref[k]))
print(json.dumps({"schedule": {k: os.environ.get("SYSID_SDP_SIGMA" + k) for k in ("0", "_GROWTH", "_CAP", "_THRESH")}, "results": out}))
