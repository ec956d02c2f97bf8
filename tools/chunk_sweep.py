"""Host-streaming entry (sysid_gram_accumulate_host): time of the 1M-sample G1 statistics from pinned host arrays vs chunk size."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
N = 1_000_000
flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", "g1_12dof.json"))
dm = ops.DeviceModel(flat)
q, dq, ddq, cnt = synth.make_trajectory(flat, N, 7)
tau = synth.synth_tau(flat, N, 3, scale=10.0)
pinned = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory() for a in (q, dq, ddq, tau, cnt)]
for chunk in (32768, 65536, 98304, 131072, 196608, 262144, 524288):
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = dm.gram_accumulate_host(*pinned, chunk=chunk)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("chunk %7d: %.2f ms (min %.2f)" % (chunk, 1e3 * np.mean(ts[2:]), 1e3 * min(ts[2:])))
