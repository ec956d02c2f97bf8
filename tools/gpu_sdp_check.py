"""GPU check of stage 3: ADMM kernel vs the oracle's barrier-Newton solve on the same statistics."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from system_identification_b200.model import FlatModel
from system_identification_b200 import synth, ops
from system_identification_b200.sys_identification import SystemIdentification
from oracle import sdp

def make(name, N):
    m = FlatModel.load(f'/root/repo/system_identification_b200/robots/{name}.json')
    si = SystemIdentification.from_flat_model(m)
    dm = si.device_model
    q, dq, ddq, cnt = synth.make_trajectory(m, N, synth.SEEDS[name])
    dev = [ops.to_device(a) for a in (q, dq, ddq)]
    tau0 = ops.to_device(np.zeros((m.joints_dof, N))); cnt_d = ops.to_device(cnt)
    Y = dm.regressor_batch(*dev).cpu().numpy()
    _, _, P = dm.projected_batch(*dev, tau0, cnt_d, want_P=True)
    rng = np.random.default_rng(5)
    phi_true = m.body_params[1:].reshape(-1) * (1 + 0.15 * rng.standard_normal(10 * m.nbodies))
    sc = 1.0 if name == 'solo12' else 10.0
    bv = rng.uniform(0, 0.02, m.joints_dof) * sc; bc = rng.uniform(0, 0.05, m.joints_dof) * sc
    tau = synth.torques_from_truth(m, Y, P.cpu().numpy(), dq, phi_true, bv, bc, 0.05 * sc, 1)
    return m, si, (q, dq, ddq, tau, cnt)

for name in ['solo12', 'spot', 'g1_12dof']:
    m, si, data = make(name, 1500)
    stats = si.gram(*data)
    sh = stats.cpu().numpy()
    c = m.ncols(True); G = sh[:c*c].reshape(c, c); r = sh[c*c:c*c+c]; s = sh[c*c+c]; n = sh[c*c+c+1]
    for lam, reg in [(0.1, 'constant_pullback'), (1e-3, 'constant_pullback'), (1e-2, 'euclidean')]:
        prob = sdp.build_problem(G, r, s, n, m.nbodies, m.phi_prior, m.robot_mass, m.ellipsoids, m.joints_dof, lambda_reg=lam, reg_type=reg)
        x0 = np.concatenate([m.phi_prior.astype(float), np.ones(2 * m.joints_dof)])
        t0 = time.time()
        try:
            xo, info_o = sdp.solve_barrier(prob, x0)
        except Exception as e:
            print(name, lam, reg, 'oracle failed', e); continue
        t_or = time.time() - t0
        torch.cuda.synchronize(); t0 = time.time()
        x, info = ops.sdp_solve(stats, m.nbodies, m.joints_dof, m.phi_prior, m.ellipsoids, m.robot_mass, lambda_reg=lam, reg_type=reg)
        torch.cuda.synchronize(); t_gpu = time.time() - t0
        x = x[0].cpu().numpy()
        L = m.nbodies
        glob = np.linalg.norm(x - xo) / np.linalg.norm(xo)
        phi_rel = np.linalg.norm(x[:10*L] - xo[:10*L]) / np.linalg.norm(xo[:10*L])
        link = max(np.linalg.norm((x - xo)[10*i:10*i+10]) / np.linalg.norm(xo[10*i:10*i+10]) for i in range(L))
        F = sdp.lmi_values(prob, xo)
        print(f"{name} lam={lam} {reg}: glob {glob:.2e} phi {phi_rel:.2e} maxlink {link:.2e} | gpu {t_gpu*1e3:.2f} ms it={info[0]['iterations']} st={info[0]['status']} refac={info[0]['refactorizations']} rp={info[0]['primal_residual']:.1e} rd={info[0]['dual_residual']:.1e} rho={info[0]['rho']:.3g} obj={info[0]['objective']:.10g} vs {info_o['objective']:.10g} minJ={info[0]['min_eig_J']:.3e} minC={info[0]['min_eig_C']:.3e} mres={info[0]['mass_residual']:.1e} | oracle {t_or:.2f}s minEig {np.linalg.eigvalsh(F).min():.2e}")
    # end-to-end identify
    t0 = time.time(); phi, bv_, bc_, inf = si.identify(*data, return_info=True); torch.cuda.synchronize()
    print('  identify() wall', time.time() - t0, 'rmse prior/ident', si.tau_prediction_rmse(*data, m.phi_prior.astype(float))[0], si.tau_prediction_rmse(*data, phi)[0])
