"""Newton-step counts of the LMI solve on the headline problem (1 M-sample G1 log) for the statistics as they come out of 1, 2, 4
and 8 shards (different summation orders: the problems differ in the last bits only).  Diagnostic for the solver's start / stall
policies (SYSID_SDP_START, SYSID_SDP_STALL_BREAK)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from system_identification_b200 import ops
from system_identification_b200.distributed import shard_bounds
from system_identification_b200.identify import _plan_for
from system_identification_b200.sys_identification import SystemIdentification
N = 1_000_000
flat = bench.load_flat(); si = SystemIdentification.from_flat_model(flat); dm = si.device_model
q, dq, ddq, tau, cnt = bench.host_log(flat, N)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
plan = _plan_for(si, 13, 12, 1e-1, 1e-10, 1000, "constant_pullback")
out = {}
for seed in (17, 18):
    dev[3] = bench.identifiable_tau(flat, dm, dev, seed=seed)
    for R in (1, 2, 4, 8, 3, 5):
        st = torch.zeros(dm.stats_len(True), dtype=torch.float64, device="cuda")
        for r in range(R):
            lo, hi = shard_bounds(N, r, R)
            dm.gram_accumulate(*[a[:, lo:hi] for a in dev], stats=st)
        x, info = plan.solve(st)
        out[f"seed{seed}_shards{R}"] = (int(info[0]["iterations"]), int(info[0]["refactorizations"]), int(info[0]["status"]))
print(json.dumps(out))
