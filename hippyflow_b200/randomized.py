"""Double-pass randomized eigensolvers on the device: hIPPYlib's ``doublePass`` / ``doublePassG``
(called at hippyflow/modeling/PODProjector.py:376, activeSubspaceProjector.py:449-461,556-577,654,
KLEProjector.py:163,177) for operators that live in HBM.

Algorithm (s = 1 at every reference call site):
    doublePass (A, Omega, k):          Q = orth(A Omega);             T = Q^T A Q; eigh; U = Q V
    doublePassG(A, B, Binv, Omega, k): Q = B-orth(Binv A Omega);      T = Q^T A Q; eigh; U = Q V, U^T B U = I
Differences from the column-by-column hIPPYlib code, none of which changes d or span(U):
  * A is applied to all m = k + p columns at once (two DMMA GEMMs + one NCCL allreduce per pass);
  * (B-)orthonormalisation is Cholesky-QR / eig-QR on the Gram matrix instead of MGS (linalg.b_orthonormalize);
  * when the operator exposes ``rayleigh`` the small matrix T = Q^T A Q is formed as a Gram matrix of the
    projected samples (one GEMM less, only an (m x m) allreduce); ``faithful=True`` forces T = (A Q)^T Q.
"""
import numpy as np

from . import _lib as K
from .linalg import b_orthonormalize, top_k_eig
from .multivector import DeviceMultiVector


def _block_apply(A, X):
    Y = DeviceMultiVector(X.tensor().shape[0], X.nvec(), device=X.tensor().device)
    if hasattr(A, "matMvMult"):
        A.matMvMult(X, Y)
    else:  # hippylib MatMvMult fallback: column loop over A.mult
        for j in range(X.nvec()):
            A.mult(X[j], Y[j])
    return Y


def doublePass(A, Omega, k, s=1, faithful=False, info=None):
    """d (k,) descending, U DeviceMultiVector (n, k) with U^T U = I."""
    nvec = Omega.nvec()
    assert k <= nvec
    Q = DeviceMultiVector(Omega)
    for _ in range(s):
        Q = _block_apply(A, Q)
    Qt, _, oinfo = b_orthonormalize(Q.tensor(), None, return_BQ=False)
    Q = DeviceMultiVector(Qt)
    if hasattr(A, "rayleigh") and not faithful:
        T = A.rayleigh(Q, Q)
    else:
        AQ = _block_apply(A, Q)
        T = AQ.dot_mv(Q)
    d, V = top_k_eig(T, k)
    U = DeviceMultiVector(K.dgemm(K.HFB_NN, Q.tensor(), K.to_padded(V, Q.tensor().device)))
    if info is not None:
        info.update(oinfo)
    return d, U


def doublePassG(A, B, Binv, Omega, k, s=1, faithful=False, info=None):
    """Generalised problem A u = lambda B u.  ``B`` is a linalg.CsrMatrix (or any object with
    ``matmat``); ``Binv`` an object with ``solve_block(Y) -> B^-1 Y``; it is not needed when A exposes
    ``solveB_matMvMult`` (A = B C B, so B^-1 A X = C B X without a solve).
    Returns d (k,), U (n, k) with U^T B U = I."""
    nvec = Omega.nvec()
    assert k <= nvec
    Q = DeviceMultiVector(Omega)
    for _ in range(s):
        if hasattr(A, "solveB_matMvMult") and getattr(A, "B", None) is B and (Binv is None or not faithful):
            Y = DeviceMultiVector(Q.tensor().shape[0], nvec, device=Q.tensor().device)
            A.solveB_matMvMult(Q, Y)
            Q = Y
        else:
            Ybar = _block_apply(A, Q)
            Q = DeviceMultiVector(Binv.solve_block(Ybar.tensor()))
    Qt, BQt, oinfo = b_orthonormalize(Q.tensor(), B, return_BQ=True)
    Q, BQ = DeviceMultiVector(Qt), DeviceMultiVector(BQt)
    if hasattr(A, "rayleigh") and not faithful:
        T = A.rayleigh(Q, BQ)
    else:
        AQ = _block_apply(A, Q)
        T = AQ.dot_mv(Q)
    d, V = top_k_eig(T, k)
    U = DeviceMultiVector(K.dgemm(K.HFB_NN, Q.tensor(), K.to_padded(V, Q.tensor().device)))
    if info is not None:
        info.update(oinfo)
    return d, U
