import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from system_identification_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "solo12"
flat, data = H.small_log(name, 8)
q, dq, ddq, tau, cnt = data
dm = ops.DeviceModel(flat)
t = H.oracle_tree(flat)
for pat in ([1,1,1,1],[1,0,0,1],[0,0,0,0],[1,0,0,0],[0,1,1,1],[0,0,1,0]):
    c2 = np.array(cnt); 
    for k in range(c2.shape[0]): c2[k,:] = pat[k] if k < len(pat) else 0
    d2 = (q, dq, ddq, tau, c2)
    for n in (1, 4, 8):
        dd = tuple(a[:, :n] for a in d2)
        st = dm.gram_accumulate(*(ops.to_device(a) for a in dd)).cpu().numpy()
        G = st[:154*154].reshape(154,154)
        A, b = H.dy.stacked_system(t, *dd, flat.ee_names)
        Go = A.T @ A
        err = np.abs(G - Go)
        i, j = np.unravel_index(err.argmax(), err.shape)
        print(name, pat[:c2.shape[0]], "n=%d" % n, "rel %.2e" % H.rel(G, Go), "max at", (i, j), "G", G[i, j], "Go", Go[i, j])
