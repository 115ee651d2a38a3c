"""Batched randomized SVD of stored Jacobians on the device.

The reference obtains the low-rank factors of every Jacobian with hIPPYlib's ``accuracyEnhancedSVD(J, Omega, r, s=1)``,
one sample at a time, each operator apply a pair of PDE solves (hippyflow/modeling/activeSubspaceProjector.py:816,1026,
dataGenerator.py:187, hippylibModelWrapper.py:285) and stores U (N, dQ, r), sigma (N, r), V (N, dM, r) in
``Jsvd_data.npz`` (dataGenerator.py:643-655).  On stored Jacobians J (N, dQ, dM) the same algorithm runs for ALL samples at
once; no Python loop over samples, no host LAPACK:

    Y_i  = J_i Omega                         one stacked DMMA GEMM over the (N dQ, dM) array
    Y_i  = J_i (J_i^T Y_i), s times          two strided-batch DMMA GEMMs per power iteration (hfb_dgemm_batched)
    Q_i  = orth(Y_i)                         batched one-sided Jacobi SVD in shared memory (hfb_jacobi_svd_batched;
                                             hIPPYlib: MultiVector.orthogonalize)
    B_i^T = J_i^T Q_i          (dM x l)      strided-batch DMMA GEMM
    H_i  = B_i B_i^T           (l x l)       strided-batch DMMA GEMM (K = dM)
    H_i  = W_i diag(s_i^2) W_i^T             batched Jacobi on the symmetric positive semi-definite H_i
    U_i  = Q_i W_i[:, :k],  sigma_i = s_i[:k],  V_i = B_i^T W_i[:, :k] / sigma_i

hIPPYlib factors B_i^T = Q~ R by MGS and takes the SVD of the small R; the route through H_i = R^T R gives the same
U_i, sigma_i, V_i (up to the sign of each singular pair) and resolves sigma_j to ~eps sigma_1^2 / sigma_j relative, which is
round-off for every singular value a rank-k truncation keeps (sigma_k / sigma_1 > ~1e-4).
"""
import numpy as np
import torch

from .. import _lib as K
from .projection import _dev, stacked_jacobians


def _as_batch(t2, batch, rows):
    """(batch*rows, cols) row-major block -> (batch, rows, cols) strided view."""
    ld = K._ld(t2)
    return t2.as_strided((batch, rows, t2.shape[1]), (rows * ld, ld, 1))


def accuracyEnhancedSVD_batched(J, Omega, k, s=1, device=None, chunk_bytes=6 << 30, return_info=False):
    """U (N, dQ, k), sigma (N, k), V (N, dM, k) with J_i ~ U_i diag(sigma_i) V_i^T for every stored Jacobian
    J (N, dQ, dM); ``Omega`` (dM, l) Gaussian test matrix shared by the samples (activeSubspaceProjector.py:1004-1010,
    "Reusing Omega for each randomized pass"), l >= k, ``s`` power iterations (every reference call site uses s = 1)."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    N, dQ, dM = J.shape
    Om = _dev(Omega, device)
    l = Om.shape[1]
    assert Om.shape[0] == dM and k <= l
    if not (K.jacobi_svd_fits(dQ, l) and K.jacobi_svd_fits(l, l)):
        raise K.HfbError("accuracyEnhancedSVD_batched: a (%d x %d) sketch does not fit the shared-memory Jacobi kernel; "
                         "for observables this large each Jacobian is a big-GEMM problem -- use doublePass on JJT/JTJ" % (dQ, l))
    J2, J3 = stacked_jacobians(J, device)
    U = K.batched_empty(N, dQ, k, device)
    V = K.batched_empty(N, dM, k, device)
    sigma = torch.empty((N, k), dtype=torch.float64, device=device)
    sweeps = torch.empty((N, 2), dtype=torch.int32, device=device)
    ldl = ((l + 15) // 16) * 16
    chunk = int(max(1, min(N, chunk_bytes // max(1, dM * ldl * 8))))
    for i0 in range(0, N, chunk):                                       # chunks bound the (chunk, dM, l) workspace, not a per-sample loop
        i1 = min(N, i0 + chunk)
        c = i1 - i0
        Jc2, Jc3 = J2[i0 * dQ:i1 * dQ], J3[i0:i1]
        Y2 = K.dgemm(K.HFB_NN, Jc2, Om)                                 # (c dQ, l): Y_i = J_i Omega for the whole chunk
        Y3 = _as_batch(Y2, c, dQ)
        Z = K.batched_empty(c, dM, l, device)
        for _ in range(int(s)):
            K.dgemm_batched(K.HFB_TN, Jc3, Y3, out=Z)                   # Z_i = J_i^T Y_i   (dM x l)
            K.dgemm_batched(K.HFB_NN, Jc3, Z, out=Y3)                   # Y_i = J_i Z_i     (dQ x l)
        _, info_q = K.jacobi_svd_batched_(Y3)                           # Y_i <- Q_i (orthonormal columns)
        K.dgemm_batched(K.HFB_TN, Jc3, Y3, out=Z)                       # B_i^T = J_i^T Q_i (dM x l)
        H = K.dgemm_batched(K.HFB_TN, Z, Z)                             # H_i = B_i B_i^T   (l x l)
        s2, info_h = K.jacobi_svd_batched_(H)                           # H_i <- W_i (eigenvectors), s2 = sigma^2 descending
        sig = torch.sqrt(torch.clamp_min(s2[:, :k], 0.0))
        Wk = K.batched_empty(c, l, k, device)
        Wk.copy_(H[:, :, :k])
        K.dgemm_batched(K.HFB_NN, Y3, Wk, out=U[i0:i1])                 # U_i = Q_i W_i[:, :k]
        inv = torch.where(sig > 0, 1.0 / torch.where(sig > 0, sig, torch.ones_like(sig)), torch.zeros_like(sig))
        Wk.mul_(inv.unsqueeze(1))                                       # W_i[:, :k] / sigma_i
        K.dgemm_batched(K.HFB_NN, Z, Wk, out=V[i0:i1])                  # V_i = B_i^T W_i[:, :k] / sigma_i
        sigma[i0:i1].copy_(sig)
        sweeps[i0:i1, 0].copy_(info_q)
        sweeps[i0:i1, 1].copy_(info_h)
    if return_info:
        return U, sigma, V, sweeps
    return U, sigma, V
