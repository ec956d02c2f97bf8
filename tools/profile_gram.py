"""Small driver for ncu: a few launches of the fused kernel (and one LMI solve) on a 262 144-sample G1-12dof log."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
N = int(os.environ.get("PROFILE_SAMPLES", 262144))
flat = FlatModel.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "system_identification_b200", "robots", "g1_12dof.json"))
dm = ops.DeviceModel(flat)
q, dq, ddq, cnt = synth.make_trajectory(flat, N, 7)
tau = synth.synth_tau(flat, N, 3, scale=10.0)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
for it in range(5):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); st = dm.gram_accumulate(*dev); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("gram N=%d: %.3f ms  %.2f Msamples/s  %.2f TFLOP/s algorithmic" % (N, ms, N / ms / 1e3, N * 435204 / ms / 1e9))
if os.environ.get("PROFILE_SDP", "1") == "1":
    for it in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); x, info = ops.sdp_solve(st, 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass); e1.record(); torch.cuda.synchronize()
        print("sdp: %.3f ms, newton %d, outer %d, status %d" % (e0.elapsed_time(e1), info[0]["iterations"], info[0]["refactorizations"], info[0]["status"]))
