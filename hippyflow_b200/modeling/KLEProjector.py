"""KLE input basis on the device: counterpart of hippyflow/modeling/KLEProjector.py:74-199 for a prior whose
covariance is given by stored parameter draws, C = (1/N) sum_i m_i m_i^T.  The elliptic
solves behind ``prior.Rsolver`` are upstream; the eigensolve structure is the reference's:
'mass' -> doublePassG(M C M, M, Msolver) with encoder = M decoder (:163-168), 'identity' -> doublePass(C)
with encoder = copy of decoder (:175-180)."""
import time

import numpy as np
import torch

from .. import _lib as K
from ..collectives import NullCollective
from ..linalg import CsrMatrix, SampleCovariance
from ..multivector import DeviceMultiVector, mv_to_dense
from ..parameterList import ParameterList
from ..randomized import doublePass, doublePassG
from .operators import SampleCovarianceOperator, SandwichedCovarianceOperator, _as_device_rows
from .PODProjector import _default_device, _to_scipy_csr, gaussian_omega


def KLEParameterList():
    """KLEProjector.py:30-45."""
    parameters = {}
    parameters['error_test_samples'] = [50, 'Number of samples for error test']
    parameters['rank'] = [128, 'Rank of subspace']
    parameters['oversampling'] = [10, 'Oversampling parameter for randomized algorithms']
    parameters['verbose'] = [True, 'Boolean for printing']
    parameters['output_directory'] = ['./data/', 'output directory for saving arrays and plots']
    parameters['plot_label_suffix'] = ['', 'suffix for plot label']
    parameters['save_and_plot'] = [True, 'save and plot or not']
    parameters['input_decoder_name'] = ['KLE_decoder', 'string for naming']
    parameters['omega_seed'] = [1, 'seed of the Gaussian test matrix when none is supplied']
    return ParameterList(parameters)


class SampleCovariancePrior:
    """Prior described by stored draws: ``m_data`` (N_loc, n) local shard (host or device) and the mass
    matrix ``M`` (SciPy CSR or PETSc-backed).  Plays the role of the ``prior`` argument (prior.M, and the
    covariance operator hp.Solver2Operator(prior.Rsolver), KLEProjector.py:103)."""

    def __init__(self, m_data, M, device=None):
        self.device = device if device is not None else _default_device()
        self.m_data = m_data
        self.M_csr = _to_scipy_csr(M)
        self.M = CsrMatrix(self.M_csr, self.device)
        self.Msolver = None  # not needed: B^-1 A = C M for A = M C M


class MassPreconditionedCovarianceOperator(SandwichedCovarianceOperator):
    """M C M (KLEProjector.py:47-69)."""


class KLEProjector:
    def __init__(self, prior, mesh_constructor_comm=None, collective=None, parameters=None):
        self.prior = prior
        self.mesh_constructor_comm = mesh_constructor_comm
        self.collective = collective if collective is not None else NullCollective()
        self.parameters = parameters if parameters is not None else KLEParameterList()
        self.device = prior.device
        Xt = _as_device_rows(prior.m_data, self.device)
        # covariance of the stored draws; the sample shards are averaged over the collective
        self.C = SampleCovarianceOperator(SampleCovariance(Xt), self.collective, 'avg')
        self.d_KLE = None
        self.V_KLE = None
        self.M_orthogonal = None

    def _omega(self, Omega, n):
        m = self.parameters['rank'] + self.parameters['oversampling']
        if Omega is None:
            if self.collective.rank() == 0:
                Omega = gaussian_omega(n, m, self.parameters['omega_seed'], self.device)
            else:
                Omega = DeviceMultiVector(n, m, device=self.device)
            self.collective.bcast(Omega, root=0)
        elif not isinstance(Omega, DeviceMultiVector):
            Omega = DeviceMultiVector.from_dense(Omega, self.device)
        return Omega

    def random_input_projector(self):
        """Random orthonormal basis (KLEProjector.py:114-129)."""
        from ..linalg import b_orthonormalize
        n = self.C.n
        Omega = self._omega(None, n)
        Q, _, _ = b_orthonormalize(Omega.tensor(), None, return_BQ=False)
        return DeviceMultiVector(Q)

    def construct_input_subspace(self, orthogonality='mass', Omega=None, faithful=False):
        t0 = time.time()
        n = self.C.n
        Omega = self._omega(Omega, n)
        rank = self.parameters['rank']
        if orthogonality.lower() == 'mass':
            KLE_Operator = MassPreconditionedCovarianceOperator(self.C, self.prior.M)
            self.d_KLE, self.V_KLE = doublePassG(KLE_Operator, self.prior.M, self.prior.Msolver, Omega, rank, s=1,
                                                 faithful=faithful)
            self.M_orthogonal = True
            kle_decoder = self.V_KLE
            kle_encoder = DeviceMultiVector(self.prior.M.matmat(kle_decoder.tensor()))
        elif orthogonality.lower() == 'prior':
            raise NotImplementedError("orthogonality='prior' needs the SLEPc shift-invert solve with the prior "
                                      "precision R (KLEProjector.py:285-334), which stays upstream")
        elif orthogonality.lower() == 'identity':
            self.d_KLE, self.V_KLE = doublePass(self.C, Omega, rank, s=1, faithful=faithful)
            self.M_orthogonal = False
            kle_decoder = self.V_KLE
            kle_encoder = DeviceMultiVector(kle_decoder)
        else:
            raise ValueError(orthogonality)
        torch.cuda.synchronize(self.device)
        self._subspace_construction_time = time.time() - t0
        if self.parameters['verbose'] and self.collective.rank() == 0:
            print('Construction of input subspace took ', self._subspace_construction_time, 's')
        if self.collective.rank() == 0 and self.parameters['save_and_plot']:
            np.save(self.parameters['output_directory'] + self.parameters['input_decoder_name'], mv_to_dense(self.V_KLE))
            np.save(self.parameters['output_directory'] + 'KLE_d', self.d_KLE)
        return self.d_KLE, kle_decoder, kle_encoder
