"""TEST INFRASTRUCTURE -- CPU restatement of the reference's unconstrained solver.  Only tests/ may
import this.  Parity unpinned against the reference run (src/solver.py imports cvxpy, absent here), but the function is
five numpy calls restated verbatim from reference src/solver.py:32-39, and numpy IS the reference's dependency."""
import numpy as np


def solve_llsq_svd(Y, tau):
    """reference src/solver.py:32-39."""
    U, Sigma, VT = np.linalg.svd(Y, full_matrices=False)
    Sigma_inv = np.linalg.pinv(np.diag(Sigma))
    A_psudo = VT.T @ Sigma_inv @ U.T
    return A_psudo @ tau
