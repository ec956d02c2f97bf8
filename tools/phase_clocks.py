"""Phase split of the fused kernel (needs a build with SYSID_NVCC_EXTRA=-DSYSID_PHASE_CLOCKS): per-CTA clock64 sums of the
F / C / M phases, read back from the partial workspace.  Diagnostic only -- never a bench number."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
N = int(os.environ.get("PROFILE_SAMPLES", 262144))
robot = os.environ.get("PROFILE_ROBOT", "g1_12dof")
flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", robot + ".json"))
dm = ops.DeviceModel(flat)
q, dq, ddq, cnt = synth.make_trajectory(flat, N, 7)
tau = synth.synth_tau(flat, N, 3, scale=10.0)
dev = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); st = dm.gram_accumulate(*dev); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s gram N=%d: %.3f ms  %.2f Msamples/s" % (robot, N, ms, N / ms / 1e3))
ws = dm._ws[("gram", dev[0].device)].view(torch.float64).cpu().numpy()
per = ws.size // 148
ws = ws[:148 * per].reshape(148, per)
names = os.environ.get("PHASE_NAMES", "qcols,fill,M,stage+sincos,chains,feet,qbuild").split(",")
clk = ws[:, 210 * 64 + 3: 210 * 64 + 3 + len(names)]
tot = clk.sum(1).mean()
spc = N / 148.0
for k, nm in enumerate(names):
    print("phase %s: %.0f clk/CTA  (%.1f%%)  %.0f clk/sample" % (nm, clk[:, k].mean(), 100 * clk[:, k].mean() / tot, clk[:, k].mean() / spc))
print("total %.0f clk/CTA = %.3f ms at 1.965 GHz; %.0f clk/sample" % (tot, tot / 1.965e6, tot / spc))
