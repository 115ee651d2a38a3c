"""Projection-error tests and per-sample Jacobian SVDs on the device.

* ``PriorPreconditionedProjector``: y = U U^T C^-1 x (hippyflow/modeling/priorPreconditionedProjector.py:48-55).
* ``projection_errors``: the rank sweep of KLEProjector.test_errors / PODProjector.test_output_errors
  (KLEProjector.py:200-280, PODProjector.py:392-476) for ALL test samples and ALL ranks with GEMMs:
  relative l2 error ||x - V_r (E_r^T x)|| / ||x||, mean and standard deviation over the samples, averaged over the
  collective the way the reference does (allReduce 'avg' of the mean and of the variance).
* ``jacobian_truncated_svd``: (U, sigma, V) of every stored Jacobian, the arrays ``Jsvd_data.npz`` holds
  (dataGenerator.py:187,643-655; the reference obtains them with hIPPYlib's accuracyEnhancedSVD on the
  matrix-free Jacobian): the same randomized algorithm, batched over all samples (``randomizedSVD.py``).
"""
import numpy as np
import torch

from .. import _lib as K
from ..collectives import NullCollective
from ..multivector import DeviceMultiVector
from .operators import _as_device_rows


class PriorPreconditionedProjector:
    def __init__(self, U, Cinv, my_init_vector=None):
        """U: DeviceMultiVector (n, r); Cinv: object with ``matmat`` (e.g. linalg.CsrMatrix)."""
        self.U = U
        self.Cinv = Cinv
        self.my_init_vector = my_init_vector

    def init_vector(self, x, dim):
        if self.my_init_vector is not None:
            self.my_init_vector(x, dim)
        else:
            x.init(self.U.tensor().shape[0])

    def matMvMult(self, X, Y):
        CinvX = self.Cinv.matmat(X.tensor())
        coef = K.dgemm(K.HFB_TN, self.U.tensor(), CinvX)            # U^T C^-1 X   (r x m)
        K.dgemm(K.HFB_NN, self.U.tensor(), coef, out=Y.tensor())    # U (...)

    def mult(self, x, y):
        self.matMvMult(DeviceMultiVector(x.storage_tensor()), DeviceMultiVector(y.storage_tensor()))


def projection_errors(test_data, decoder, encoder, ranks, collective=None, device=None):
    """(avg_rel_errors, std_rel_errors), one entry per rank in ``ranks`` (sorted ascending like the reference does).
    ``test_data`` (N, n) rows = test samples (local shard); decoder / encoder (n, r_max) arrays or multivectors
    (encoder = M decoder for an M-orthogonal basis, = decoder otherwise)."""
    collective = collective if collective is not None else NullCollective()
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    X = _as_device_rows(test_data, device)
    V = decoder.tensor() if hasattr(decoder, "tensor") else K.to_padded(np.asarray(decoder), device)
    E = encoder.tensor() if hasattr(encoder, "tensor") else K.to_padded(np.asarray(encoder), device)
    ranks = sorted(int(r) for r in ranks)
    assert ranks[-1] <= V.shape[1]
    N, n = X.shape
    coef = K.dgemm(K.HFB_NN, X, E)                                   # (N, r_max) = encoder^T x_i for all samples
    resid = K.padded_empty(N, n, device)
    resid.copy_(X)
    denom = torch.sqrt(K.rowdot(X, X))
    avg, std = [], []
    r_prev = 0
    for r in ranks:
        if r > r_prev:
            inc = K.dgemm(K.HFB_NT, K.to_padded(coef[:, r_prev:r], device), K.to_padded(V[:, r_prev:r], device))
            K.axpby_(-1.0, inc, 1.0, resid)                         # resid -= C[:, r_prev:r] V[:, r_prev:r]^T
            r_prev = r
        rel = (torch.sqrt(K.rowdot(resid, resid)) / denom).cpu().numpy()
        avg.append(collective.allReduce(float(np.mean(rel)), "avg"))
        std.append(np.sqrt(collective.allReduce(float(np.std(rel) ** 2), "avg")))
    return np.array(avg), np.array(std)


def jacobian_truncated_svd(J, rank, device=None, oversampling=10, Omega=None, s=1, seed=1):
    """U (N, dQ, r), sigma (N, r), V (N, dM, r) with J_i ~ U_i diag(sigma_i) V_i^T, sigma descending: the arrays of
    ``Jsvd_data.npz`` (dataGenerator.py:643-655), computed the way the reference does -- hIPPYlib's
    accuracyEnhancedSVD(J, Omega, r, s=1) with r + oversampling Gaussian columns (activeSubspaceProjector.py:1004-1026) --
    for all samples at once (``randomizedSVD.accuracyEnhancedSVD_batched``: strided-batch DMMA GEMMs + batched Jacobi
    kernels, no loop over samples, no host LAPACK)."""
    from .randomizedSVD import accuracyEnhancedSVD_batched
    from .PODProjector import gaussian_omega
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    N, dQ, dM = J.shape
    r = int(rank)
    assert r <= min(dQ, dM)
    if Omega is None:
        Omega = gaussian_omega(dM, min(r + int(oversampling), max(r, min(dQ, dM))), seed, device)
    return accuracyEnhancedSVD_batched(J, Omega, r, s=s, device=device)
