"""The structured null-space basis of gram_struct_kernel (6 dense vectors + sparse vectors supported on one leg each), restated in
numpy (oracle/struct_basis.py), spans exactly the null space the reference projects on: Q Q^T == I - pinv(J_c) J_c
(reference src/sys_identification.py:127-135).  CPU test; the kernel itself is compared with the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import helpers as H
from oracle import dynamics as D
from oracle.struct_basis import struct_basis


@pytest.mark.parametrize("name", H.ROBOTS)
def test_structured_basis_equals_pinv_projector(name):
    flat = H.flat_model(name)
    tree = H.oracle_tree(flat)
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, 6000, 11)
    seen = set()
    for i in range(0, 6000, 125):
        Jc, dense, sparse, chains = struct_basis(flat, tree, q[:, i], cnt[:, i])
        vecs = dense + [v for ci in sparse for v in sparse[ci]]
        Q = np.array(vecs).T
        P = np.eye(flat.nv) - (np.linalg.pinv(Jc) @ Jc if Jc.shape[0] else 0.0)
        assert Q.shape[1] == flat.nv - np.linalg.matrix_rank(Jc) if Jc.shape[0] else Q.shape[1] == flat.nv
        assert np.abs(Q.T @ Q - np.eye(Q.shape[1])).max() <= 1e-12          # orthonormal
        assert np.abs(Q @ Q.T - P).max() <= 1e-12
        assert len(dense) == 6                                              # whatever the contact state
        for ci, vs in sparse.items():                                       # a sparse vector lives on the joints of its own leg
            cols = [4 + j for j in chains[ci]]
            for v in vs:
                w = v.copy(); w[cols] = 0.0
                assert np.abs(w).max() == 0.0
        seen.add(int(np.count_nonzero(cnt[:, i])))
    assert len(seen) >= 2                                                   # the trajectory visits several contact states
