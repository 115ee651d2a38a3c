"""Projection of stored training data onto the bases (north star (c)): reduced inputs
(M V)^T m_i, reduced outputs, J_i^T (M Phi), J_i Psi and Phi^T J_i V for all samples at once.  In the
reference these are per-sample column loops of PDE solves (dataGenerator.py:170,177,339); on stored data
they are GEMMs over the stacked sample axis."""
import numpy as np
import torch

from .. import _lib as K
from .operators import _as_device_rows


def _dev(a, device):
    if hasattr(a, "tensor"):
        return a.tensor()
    if K.is_device_tensor(a):
        return a if (a.dim() == 2 and a.stride(1) == 1 and K._ld(a) % 2 == 0 and a.data_ptr() % 16 == 0) else K.to_padded(a, device)
    return K.to_padded(np.asarray(a, dtype=np.float64), device)


def stacked_jacobians(J, device):
    """Stored Jacobians (N, dQ, dM) as (J2, J3): the (N dQ, dM) row-major device matrix (TMA-conforming leading dimension,
    zero-copy when the input already conforms) and its (N, dQ, dM) strided-batch view."""
    N, dQ, dM = J.shape
    J2 = _as_device_rows(J.reshape(N * dQ, dM), device)
    ld = K._ld(J2)
    return J2, J2.as_strided((N, dQ, dM), (dQ * ld, ld, 1))


def project_data(data, encoder, device=None, out=None):
    """(N, r) reduced coordinates: row i = encoder^T data_i, with encoder = M decoder
    (KLEProjector.py:167-168, PODProjector.py:769,830).  ``data`` (N, n) host array or device block."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    X = _as_device_rows(data, device)
    E = _dev(encoder, device)
    return K.dgemm(K.HFB_NN, X, E, out=out)


def jacobian_action(J, Psi, device=None):
    """JPsi (N, dQ, rM): JPsi_i = J_i Psi (dataGenerator.py:177, stacked as at :585)."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    N, dQ, dM = J.shape
    Jd = _as_device_rows(J.reshape(N * dQ, dM), device)
    P = _dev(Psi, device)
    out = K.dgemm(K.HFB_NN, Jd, P)                                  # (N*dQ, rM)
    rM = P.shape[1]
    return out.as_strided((N, dQ, rM), (dQ * out.stride(0), out.stride(0), 1))


def jacobian_transpose_action(J, MPhi, device=None):
    """JstarPhi (N, dM, rQ): JstarPhi_i = J_i^T (M Phi) (dataGenerator.py:170,339, stacked as at :582)."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    _, J3 = stacked_jacobians(J, device)
    E = _dev(MPhi, device)                                          # (dQ, rQ), shared by every sample
    return K.dgemm_batched(K.HFB_TN, J3, E)                         # ONE strided-batch launch: grid = samples x tiles


def reduced_jacobians(J, PhiEnc, V, device=None):
    """(N, rQ, rM): Phi_enc^T J_i V for every sample: one stacked GEMM J_all V, then a batched small GEMM."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    N, dQ, dM = J.shape
    JV = jacobian_action(J, V, device)                              # (N, dQ, rM)
    Phi = _dev(PhiEnc, device)                                      # (dQ, rQ)
    rQ, rM = Phi.shape[1], JV.shape[2]
    out = torch.empty((N, rQ, rM), dtype=torch.float64, device=device)
    K.dgemm_batched_small(K.HFB_TN, Phi.unsqueeze(0), JV, out)
    return out
