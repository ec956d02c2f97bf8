import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from system_identification_b200 import ops
from system_identification_b200.identify import _plan_for
from system_identification_b200.sys_identification import SystemIdentification
N = 1_000_000
flat = bench.load_flat()
si = SystemIdentification.from_flat_model(flat)
dm = si.device_model
q, dq, ddq, tau, cnt = bench.host_log(flat, N)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
plan = _plan_for(si, 13, 12, 1e-1, 1e-10, 1000, "constant_pullback")
st0 = dm.gram_accumulate(*[a[:, :131072] for a in dev]).clone()
side = torch.cuda.Stream()
def run(with_solve, chunks=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stats = torch.zeros(dm.stats_len(True), dtype=torch.float64, device="cuda")
    n = N // chunks
    ev = None
    for k in range(chunks):
        dm.gram_accumulate(*[a[:, k * n:(k + 1) * n] for a in dev], stats=stats)
        if with_solve and k == 0:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); plan.solve(st0, sync_info=False); e1.record(); ev = (e0, e1)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    return dt, (ev[0].elapsed_time(ev[1]) if ev else 0.0)
for _ in range(2): run(False); run(True)
print("reserve", os.environ.get("SYSID_DEBUG_RESERVE_SMS"), "8 chunks, no solve:", min(run(False)[0] for _ in range(5)))
r = [run(True) for _ in range(5)]
print("8 chunks, concurrent cold solve: total", min(x[0] for x in r), "solve kernel", min(x[1] for x in r))
