import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
from oracle import urdf_tree as ut, dynamics as dy
torch.cuda.init()
for name in ['solo12', 'spot', 'g1_12dof']:
    m = FlatModel.load(f'/root/repo/system_identification_b200/robots/{name}.json')
    t = ut.tree_from_flat(m)
    N = 70
    q, dq, ddq, cnt = synth.make_trajectory(m, N, synth.SEEDS[name])
    tau = synth.synth_tau(m, N, 3)
    dm = ops.DeviceModel(m)
    dq_, ddq_, q_, tau_, cnt_ = [ops.to_device(a) for a in (dq, ddq, q, tau, cnt)]
    Y = dm.regressor_batch(q_, dq_, ddq_).cpu().numpy()
    Yo = np.array([dy.joint_torque_regressor(t, q[:, i], dq[:, i], ddq[:, i]) for i in range(N)])
    print(name, 'Y rel err', np.abs(Y - Yo).max() / np.abs(Yo).max(), 'struct zeros equal', ((Y == 0) == (Yo == 0)).all())
    A, b, P = dm.projected_batch(q_, dq_, ddq_, tau_, cnt_, friction=True, want_P=True)
    A, b, P = A.cpu().numpy(), b.cpu().numpy(), P.cpu().numpy()
    Ao, bo = dy.stacked_system(t, q, dq, ddq, tau, cnt, m.ee_names)
    Po = np.array([dy.null_space_projector(t, q[:, i], cnt[:, i], m.ee_names) for i in range(N)])
    print('  P err', np.abs(P - Po).max(), 'A rel err', np.abs(A.reshape(-1, A.shape[-1]) - Ao).max() / np.abs(Ao).max(), 'b rel', np.abs(b.reshape(-1) - bo).max() / np.abs(bo).max())
    stats = dm.gram_accumulate(q_, dq_, ddq_, tau_, cnt_).cpu().numpy()
    c = Ao.shape[1]
    G = stats[:c * c].reshape(c, c); r = stats[c * c:c * c + c]; s = stats[c * c + c]; n = stats[c * c + c + 1]
    Go, ro, so, no = dy.gram_from_stack(Ao, bo)
    print('  G rel fro', np.linalg.norm(G - Go) / np.linalg.norm(Go), 'r', np.linalg.norm(r - ro) / np.linalg.norm(ro), 's', abs(s - so) / so, 'n', n, no, 'sym', np.abs(G - G.T).max())
    st2 = ops.gram_from_stack(torch.from_numpy(Ao).cuda(), torch.from_numpy(bo).cuda()).cpu().numpy()
    print('  stack-gram rel', np.linalg.norm(st2[:c * c].reshape(c, c) - Go) / np.linalg.norm(Go), np.linalg.norm(st2[c*c:c*c+c] - ro) / np.linalg.norm(ro), st2[-2] / so - 1, st2[-1])
    phi = m.phi_prior.astype(np.float64)
    out = dm.predict_rmse(q_, dq_, ddq_, tau_, cnt_, torch.from_numpy(phi)).cpu().numpy()
    tot, pj = dy.tau_prediction_rmse(t, q, dq, ddq, tau, cnt, phi, m.ee_names)
    print('  rmse total rel', abs(out[0] - tot) / tot, 'per joint rel', np.abs(out[1:] - pj).max() / pj.max())
# throughput quick look
m = FlatModel.load('/root/repo/system_identification_b200/robots/g1_12dof.json')
dm = ops.DeviceModel(m)
N = 1 << 18
q, dq, ddq, cnt = synth.make_trajectory(m, N, 7)
tau = synth.synth_tau(m, N, 3)
q_, dq_, ddq_, tau_, cnt_ = [ops.to_device(a) for a in (q, dq, ddq, tau, cnt)]
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); st = dm.gram_accumulate(q_, dq_, ddq_, tau_, cnt_); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('gram N=%d: %.3f ms  %.2f Msamples/s  %.2f TFLOP/s algorithmic' % (N, ms, N / ms / 1e3, N * 435204 / ms / 1e9))
