"""Active-subspace input/output bases from STORED Jacobians on the device: counterpart of
hippyflow/modeling/activeSubspaceProjector.py:252-673 (eigensolve call sites :447-463, :553-577, :654).
Generating the Jacobians (adjoint PDE solves, :690-1044) is upstream; here ``observable`` carries the local
(N_loc, dQ, dM) array, as the user of ``MeanJTJfromDataOperator`` (operatorWrappers.py:55-121) would hold it."""
import time

import numpy as np
import torch

from .. import _lib as K
from ..collectives import NullCollective
from ..linalg import CsrMatrix, CsrCGSolver, SampleCovariance
from ..multivector import DeviceMultiVector, mv_to_dense
from ..parameterList import ParameterList
from ..randomized import doublePass, doublePassG
from .operators import MeanJJTfromDataOperator, MeanJTJfromDataOperator, SampleCovarianceOperator
from .PODProjector import _default_device, _to_scipy_csr, gaussian_omega


def ActiveSubspaceParameterList():
    """activeSubspaceProjector.py:33-66 (keys that concern the stored-data path keep their defaults)."""
    parameters = {}
    parameters['samples_per_process'] = [64, 'Number of samples per process']
    parameters['jacobian_data_per_process'] = [512, 'Number of samples per process']
    parameters['error_test_samples'] = [50, 'Number of samples for error test']
    parameters['rank'] = [128, 'Rank of subspace']
    parameters['jacobian_rank'] = [128, 'Rank of Jacobians generated']
    parameters['control_jacobian_rank'] = [None, 'Rank of control Jacobians generated']
    parameters['oversampling'] = [10, 'Oversampling parameter for randomized algorithms']
    parameters['double_loop_samples'] = [20, 'Number of samples used in double loop MC approximation']
    parameters['verbose'] = [True, 'Boolean for printing']
    parameters['input_decoder_name'] = ['_input_decoder', 'string for naming']
    parameters['output_decoder_name'] = ['_output_decoder', 'string for naming']
    parameters['initialize_samples'] = [False, 'Boolean for the initialization of samples']
    parameters['serialized_sampling'] = [True, 'Boolean for the serialization of sampling on a process']
    parameters['observable_constructor'] = [None, 'observable constructor function']
    parameters['observable_kwargs'] = [{}, 'kwargs used when instantiating observables']
    parameters['output_directory'] = [None, 'output directory for saving arrays and plots']
    parameters['plot_label_suffix'] = ['', 'suffix for plot label']
    parameters['save_and_plot'] = [True, 'Boolean for saving data and plots (only False for unit testing)']
    parameters['store_Omega'] = [False, 'Boolean for storing Gaussian random matrix (only True for unit testing)']
    parameters['ms_given'] = [False, 'Boolean for passing ms into serialized AS construction']
    parameters['omega_seed'] = [1, 'seed of the Gaussian test matrix when none is supplied']
    return ParameterList(parameters)


class StoredJacobians:
    """``observable`` stand-in: local stored Jacobians J (N_loc, dQ, dM), optional noise precision (dQ, dQ)."""

    def __init__(self, J, noise_cov_inv=None):
        assert len(J.shape) == 3
        self.J = J
        self.noise_cov_inv = noise_cov_inv


class SparsePrior:
    """``prior`` stand-in exposing ``R`` (CSR precision-like SPD matrix on the device) and ``Rsolver`` (block CG),
    the two objects doublePassG receives at activeSubspaceProjector.py:449-450."""

    def __init__(self, R, device=None, rel_tol=1e-13, max_iter=1000, Rsolver=None, on_fail="raise"):
        """``Rsolver``: any object with ``solve_block(Y) -> R^-1 Y`` on device blocks (default: Jacobi-preconditioned
        block CG, which raises when it does not reach ``rel_tol`` within ``max_iter`` iterations -- an unconverged
        R^-1 silently degrades the eigenpairs of doublePassG)."""
        self.device = device if device is not None else _default_device()
        self.R_csr = _to_scipy_csr(R)
        self.R = CsrMatrix(self.R_csr, self.device)
        self.Rsolver = Rsolver if Rsolver is not None else CsrCGSolver(self.R, rel_tol=rel_tol, max_iter=max_iter,
                                                                       on_fail=on_fail)


class LowRankHessian:
    """hIPPYlib ``LowRankHessian`` (the object behind ``prior.Hlr``, used as B and as its own solver at
    activeSubspaceProjector.py:455-459,562-566): H = R + R U diag(d) U^T R with U^T R U = I, and by Sherman-Morrison-
    Woodbury H^-1 = R^-1 - U diag(d / (1 + d)) U^T.  R is a device CSR matrix with a block solver (``SparsePrior``); U a
    (n, r) device block, d (r,).  Exposes the block protocol doublePassG uses (``matmat`` / ``solve_block``) and the
    vector protocol of the reference (``mult`` / ``solve`` / ``init_vector``)."""

    def __init__(self, prior, d, U):
        self.prior = prior
        self.device = prior.device
        self.shape = prior.R.shape
        self.U = U.tensor() if hasattr(U, "tensor") else K.to_padded(np.asarray(U, dtype=np.float64), self.device)
        self.d = torch.as_tensor(np.asarray(d, dtype=np.float64), device=self.device)
        self.RU = prior.R.matmat(self.U)                                   # R U (n, r)

    def init_vector(self, x, dim):
        x.init(self.shape[0])

    def matmat(self, X, out=None):
        Y = self.prior.R.matmat(X, out=out)                                # R X
        coef = K.rowscale(self.d, K.dgemm(K.HFB_TN, self.RU, X))           # diag(d) (R U)^T X   (r, m)
        K.axpby_(1.0, K.dgemm(K.HFB_NN, self.RU, coef), 1.0, Y)            # + R U diag(d) U^T R X
        return Y

    def solve_block(self, Y):
        X = self.prior.Rsolver.solve_block(Y)                              # R^-1 Y
        coef = K.rowscale(self.d / (1.0 + self.d), K.dgemm(K.HFB_TN, self.U, Y))
        K.axpby_(-1.0, K.dgemm(K.HFB_NN, self.U, coef), 1.0, X)            # - U diag(d/(1+d)) U^T Y
        return X

    def mult(self, x, y):
        y.storage_tensor().copy_(self.matmat(_aligned_block(x.storage_tensor())))

    def solve(self, sol, rhs):
        sol.storage_tensor().copy_(self.solve_block(_aligned_block(rhs.storage_tensor())))


def _aligned_block(t):
    if t.data_ptr() % 16 or K._ld(t) % 2:
        return K.to_padded(t, t.device, pad=2 if t.shape[1] == 1 else 16)
    return t


class ActiveSubspaceProjector:
    # largest output dimension for which E[J J^T] is formed as a dense (dQ x dQ) matrix; above it (full-state observable,
    # dQ = n_u) the operator form MeanJJTfromDataOperator is handed to doublePass
    DENSE_OUTPUT_MAX = 4096

    def __init__(self, observable, prior=None, control_distribution=None, mesh_constructor_comm=None,
                 collective=NullCollective(), parameters=None, device=None):
        self.observable = observable
        self.prior = prior
        self.control_distribution = control_distribution
        self.mesh_constructor_comm = mesh_constructor_comm
        self.collective = collective
        self.parameters = parameters if parameters is not None else ActiveSubspaceParameterList()
        self.device = device if device is not None else (prior.device if prior is not None else _default_device())
        self.d_GN = None
        self.V_GN = None
        self.d_NG = None
        self.U_NG = None
        self.Omega_GN = None
        self.Omega_NG = None
        self.prior_preconditioned = None
        self._JTJ = None
        self._JJT = None

    def _operator(self):
        if self._JTJ is None:
            # Average_GN_Hessian = CollectiveOperator(SummedListOperator([JTJ(J_i)]), collective, 'avg') (:427-431)
            self._JTJ = MeanJTJfromDataOperator(self.observable.J, self.prior, self.observable.noise_cov_inv,
                                                device=self.device, collective=self.collective, mpi_op='avg')
        return self._JTJ

    def _omega(self, stored, n):
        m = self.parameters['rank'] + self.parameters['oversampling']
        if stored is not None:
            return stored if isinstance(stored, DeviceMultiVector) else DeviceMultiVector.from_dense(stored, self.device)
        if self.collective.rank() == 0:
            Omega = gaussian_omega(n, m, self.parameters['omega_seed'], self.device)
        else:
            Omega = DeviceMultiVector(n, m, device=self.device)
        self.collective.bcast(Omega, root=0)
        return Omega

    def construct_input_subspace(self, prior_preconditioned=True, name_suffix=None, faithful=False):
        t0 = time.time()
        A = self._operator()
        Omega = self._omega(self.Omega_GN, A.dM)
        if self.parameters['store_Omega']:
            self.Omega_GN = Omega
        rank = self.parameters['rank']
        if prior_preconditioned:
            if self.prior is not None and hasattr(self.prior, "R"):
                self.d_GN, self.V_GN = doublePassG(A, self.prior.R, self.prior.Rsolver, Omega, rank, s=1, faithful=faithful)
                as_decoder = self.V_GN
                as_encoder = DeviceMultiVector(self.prior.R.matmat(as_decoder.tensor()))   # hp.MatMvMult(prior.R, ...) :452-453
            elif self.prior is not None and hasattr(self.prior, "Hlr"):
                # doublePassG(A, prior.Hlr, prior.Hlr, ...) :455-459 -- Hlr is the operator and its own solver
                self.d_GN, self.V_GN = doublePassG(A, self.prior.Hlr, self.prior.Hlr, Omega, rank, s=1, faithful=faithful)
                as_decoder = self.V_GN
                as_encoder = DeviceMultiVector(self.prior.Hlr.matmat(as_decoder.tensor()))
            else:
                raise ValueError("prior_preconditioned=True needs a prior exposing R and Rsolver, or Hlr")
        else:
            self.d_GN, self.V_GN = doublePass(A, Omega, rank, s=1, faithful=faithful)
            as_decoder = self.V_GN
            as_encoder = DeviceMultiVector(as_decoder)
        torch.cuda.synchronize(self.device)
        self.prior_preconditioned = prior_preconditioned
        self._input_subspace_construction_time = time.time() - t0
        if self.parameters['verbose'] and self.collective.rank() == 0:
            print(('Input subspace construction took ' + str(self._input_subspace_construction_time)[:5] + ' s').center(80))
        if self.parameters['save_and_plot'] and self.collective.rank() == 0:
            name = 'AS_' + str(int(self.parameters['samples_per_process'] * self.collective.size()))
            if name_suffix is not None:
                assert type(name_suffix) is str
                name += name_suffix
            np.save(self.parameters['output_directory'] + name + self.parameters['input_decoder_name'], mv_to_dense(self.V_GN))
            np.save(self.parameters['output_directory'] + name + '_d_GN', self.d_GN)
        return self.d_GN, as_decoder, as_encoder

    def construct_output_subspace(self, name_suffix=None, operator_form=None):
        """E[J J^T] (activeSubspaceProjector.py:618-673): doublePass on the sample-averaged J J^T.  For a small output space
        (pointwise observations) the (dQ x dQ) matrix is formed by ONE strided-batch GEMM whose K loop runs over
        (sample, dM); for a large one (full-state observable, dQ = n_u > DENSE_OUTPUT_MAX, or ``operator_form=True``) the
        operator mean_i J_i (J_i^T X) is applied block-wise without ever forming the matrix."""
        t0 = time.time()
        J = self.observable.J
        N, dQ, dM = J.shape
        if getattr(self, "_JJT", None) is None:
            stacked = None
            if self._JTJ is not None:                                            # reuse the device copy of the input subspace
                J2 = self._JTJ._cov.Xt
                stacked = (J2, J2.as_strided((N, dQ, dM), (dQ * K._ld(J2), K._ld(J2), 1)))
            self._JJT = MeanJJTfromDataOperator(J, device=self.device, collective=self.collective, mpi_op='avg',
                                                stacked=stacked)
        if operator_form is None:
            operator_form = dQ > self.DENSE_OUTPUT_MAX
        if operator_form:
            A = self._JJT
        else:
            Cd = self._JJT.dense()

            class _Dense:
                overwrites = True

                def matMvMult(self_, X, Y):
                    K.dgemm(K.HFB_NN, Cd, X.tensor(), out=Y.tensor())

            A = _Dense()

        Omega = self._omega(self.Omega_NG, dQ)
        if self.parameters['store_Omega']:
            self.Omega_NG = Omega
        self.d_NG, self.U_NG = doublePass(A, Omega, min(self.parameters['rank'], dQ), s=1)
        output_decoder = self.U_NG
        output_encoder = DeviceMultiVector(output_decoder)
        torch.cuda.synchronize(self.device)
        self._output_subspace_construction_time = time.time() - t0
        if self.parameters['save_and_plot'] and self.collective.rank() == 0:
            name = 'AS_' + str(int(self.parameters['samples_per_process'] * self.collective.size()))
            if name_suffix is not None:
                name += name_suffix
            np.save(self.parameters['output_directory'] + name + self.parameters['output_decoder_name'], mv_to_dense(self.U_NG))
            np.save(self.parameters['output_directory'] + name + '_d_NG', self.d_NG)
        return self.d_NG, output_decoder, output_encoder
