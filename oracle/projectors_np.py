"""CPU restatement of the reference's projector call structure for the stored-data hot path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function follows the reference file:line it cites and drives the NumPy hIPPYlib restatement in
``hippylib_np`` exactly the way the reference drives hIPPYlib: operator object -> collective wrapper
-> Omega -> doublePass / doublePassG -> encoder.
"""
import numpy as np
import scipy.linalg as la
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import hippylib_np as hp


# ----------------------------------------------------------------------------- collectives
class NullCollective:
    """hippyflow/collectives/collective.py:19-38."""

    def bcast(self, v, root=0):
        return v

    def size(self):
        return 1

    def rank(self):
        return 0

    def allReduce(self, v, op):
        if op.lower() not in ["sum", "avg"]:
            raise NotImplementedError("Unknown operation *{0}* in NullCollective.allReduce".format(op))
        return v


class CollectiveOperator:
    """hippyflow/collectives/collectiveOperator.py:14-55: local apply, then allReduce."""

    def __init__(self, local_op, collective, mpi_op="sum"):
        assert hasattr(local_op, "mult")
        self.local_op, self.collective, self.mpi_op = local_op, collective, mpi_op

    def mult(self, x, y):
        self.local_op.mult(x, y)
        self.collective.allReduce(y, self.mpi_op)

    def transpmult(self, x, y):
        self.local_op.transpmult(x, y)
        self.collective.allReduce(y, self.mpi_op)

    def init_vector(self, x, dim):
        self.local_op.init_vector(x, dim)


class SimulatedRanksOperator:
    """P sample shards applied one after the other, then combined the way
    ``_allReduce_array`` does (collective.py:61-71): SUM over ranks, 'avg' = SUM * (1/size)."""

    def __init__(self, local_ops, mpi_op="avg"):
        self.local_ops, self.mpi_op = local_ops, mpi_op.lower()

    def init_vector(self, x, dim):
        self.local_ops[0].init_vector(x, dim)

    def mult(self, x, y):
        acc = np.zeros(x.size())
        tmp = hp.Vector(np.zeros(x.size()))
        for op in self.local_ops:
            tmp.zero()
            op.mult(x, tmp)
            acc += tmp._a
        if self.mpi_op == "avg":
            acc *= 1.0 / float(len(self.local_ops))
        y.set_local(acc)


# ----------------------------------------------------------------------------- operators
class MeanJTJfromDataOperator:
    """y = mean_i J_i^T [Gamma^-1] J_i x  (hippyflow/modeling/operatorWrappers.py:95-114):
    two contractions over the stored (ndata, r, dM) array, then the mean over samples."""

    def __init__(self, J, noise_cov_inv=None):
        self.J = J
        self.ndata, self.r, self.dM = J.shape
        self.noise_cov_inv = noise_cov_inv

    def init_vector(self, x, dim):
        x.init(self.dM)

    def mult(self, x, y):
        xv = x.get_local()
        JX = self.J @ xv                                    # (ndata, r)   == einsum('ijk,ik->ij', J, tile(x))
        if self.noise_cov_inv is not None:
            JX = JX @ np.asarray(self.noise_cov_inv).T      # einsum('ij,kj->ki', G, JX)
        JTJX = np.einsum("ijk,ij->ik", self.J, JX)          # (ndata, dM)
        y.set_local(np.mean(JTJX, axis=0))

    transpmult = mult


class SandwichedCovarianceOperator:
    """x -> M C M x with C = X X^T / N given by stored samples (rows of ``data``).
    This is ``H_matvec`` of PODProjector.py:750-754 (MX @ (MX.T @ x) / n_data) and the
    M C M structure of KLEProjector.py:66-69."""

    def __init__(self, data, M):
        self.data, self.M = data, M

    def init_vector(self, x, dim):
        x.init(self.M.shape[0])

    def mult(self, x, y):
        Mx = self.M @ x._a
        t = self.data @ Mx
        y.set_local(self.M @ (self.data.T @ t) / self.data.shape[0])


# ----------------------------------------------------------------------------- POD
def weighted_l2_norm_vector(x, W):
    """PODProjector.py:658-661."""
    return np.sqrt(np.einsum("ij,ij->j", W @ x, x))


def pod_from_data(u_data, M_csr, u_rank, shifted=True, method="hep"):
    """Deterministic M-weighted POD: PODProjectorFromData.construct_subspace,
    PODProjector.py:699-852 (shift :732-738; ghep :743-773; inverse_ghep :775-810; hep :812-833)."""
    n_data, dim_u = u_data.shape
    assert u_rank <= n_data
    if shifted:
        u_shift = np.mean(u_data, axis=0)
        u_data = u_data - u_shift
    else:
        u_shift = np.zeros(dim_u)
    X = u_data.T
    if method == "ghep":
        MX = M_csr @ X
        H = spla.LinearOperator(matvec=lambda x: MX @ (MX.T @ x) / n_data, shape=(dim_u, dim_u), dtype=np.float64)
        d, phi = spla.eigsh(H, M=M_csr, k=u_rank)
        d, phi = d[::-1][:u_rank], phi[:, ::-1][:, :u_rank]
        Mphi = M_csr @ phi
    elif method == "inverse_ghep":
        lu = spla.splu(M_csr.tocsc())
        Minv = spla.LinearOperator(shape=M_csr.shape, matvec=lu.solve, dtype=np.float64)
        H = spla.LinearOperator(matvec=lambda x: X @ (X.T @ x) / n_data, shape=(dim_u, dim_u), dtype=np.float64)
        d, Mphi = spla.eigsh(H, k=u_rank, M=Minv, Minv=spla.aslinearoperator(M_csr))
        d, Mphi = d[::-1], Mphi[:, ::-1]
        phi = Minv @ Mphi
    elif method == "hep":
        G = X.T @ M_csr @ X
        s, U = la.eigh(G)
        d = s[::-1][:u_rank] / n_data
        U = U[:, ::-1][:, :u_rank]
        phi = X @ U
        phi = phi / weighted_l2_norm_vector(phi, M_csr)
        Mphi = M_csr @ phi
    else:
        raise ValueError("Unavailable method")
    return d, phi, Mphi, u_shift


def pod_randomized_weighted(u_data, M_csr, u_rank, Omega, shifted=True, ranks=1):
    """The M-weighted randomized POD of the north star: the GHEP of PODProjector.py:750-761
    (A = M X X^T M / N, B = M) solved with doublePassG as at KLEProjector.py:163-168.
    ``ranks`` > 1 shards the samples and averages the shard operators like
    CollectiveOperator(..., 'avg') does (collectiveOperator.py:31-38, collective.py:61-71).
    Returns d (k,), decoder (n,k) with decoder^T M decoder = I, encoder = M decoder, shift."""
    if shifted:
        u_shift = np.mean(u_data, axis=0)
        u_data = u_data - u_shift
    else:
        u_shift = np.zeros(u_data.shape[1])
    if ranks == 1:
        A = CollectiveOperator(SandwichedCovarianceOperator(u_data, M_csr), NullCollective(), "avg")
    else:
        assert u_data.shape[0] % ranks == 0
        A = SimulatedRanksOperator([SandwichedCovarianceOperator(s, M_csr) for s in np.split(u_data, ranks)], "avg")
    B = hp.SparseOperator(M_csr)
    Om = hp.MultiVector.from_dense(Omega)
    d, U = hp.doublePassG(A, B, B, Om, u_rank, s=1)
    enc = hp.MultiVector(U)
    hp.MatMvMult(B, U, enc)
    return d, U.to_dense(), enc.to_dense(), u_shift


def pod_randomized_weighted_blocked(u_data, M_csr, u_rank, Omega, shifted=True):
    """Best-effort CPU evaluation of the same M-weighted double pass with every operator apply done on the whole
    (n x m) block at once (BLAS-3 instead of the reference's column-by-column hp.MatMvMult loop) and the
    M-orthonormalisation done by two eigen-based Cholesky-QR sweeps instead of MGS.  d and span(decoder) agree with
    ``pod_randomized_weighted`` (they do not depend on the choice of M-orthonormal basis of the sketch).  Used as the
    "fair" CPU baseline of SURVEY.md 8(d) next to the faithful column-by-column port."""
    if shifted:
        u_shift = np.mean(u_data, axis=0)
        X = u_data - u_shift
    else:
        u_shift = np.zeros(u_data.shape[1])
        X = u_data
    N = X.shape[0]
    Q = X.T @ (X @ (M_csr @ Omega)) / N                  # M^-1 A Omega = C M Omega
    for _ in range(2):
        G = Q.T @ (M_csr @ Q)
        w, V = np.linalg.eigh(0.5 * (G + G.T))
        keep = w > w.max() * Q.shape[1] * np.finfo(float).eps        # a rank-deficient sketch keeps its range only
        Q = Q @ (V * np.where(keep, 1.0 / np.sqrt(np.where(keep, w, 1.0)), 0.0))
    W = X @ (M_csr @ Q)
    T = W.T @ W / N                                      # Q^T A Q
    dd, VV = np.linalg.eigh(0.5 * (T + T.T))
    d = dd[::-1][:u_rank]
    U = Q @ VV[:, ::-1][:, :u_rank]
    return d, U, M_csr @ U, u_shift


def pod_randomized(u_data, rank, Omega, ranks=1):
    """PODProjector.construct_subspace, PODProjector.py:359-376: LowRankOperator(ones/N_loc, U_loc)
    per rank, CollectiveOperator(..., 'avg'), doublePass(s=1).  No weighting, no shift."""
    shards = np.split(u_data, ranks)
    ops = []
    for s in shards:
        U_loc = hp.MultiVector.from_dense(s.T)
        ops.append(hp.LowRankOperator(np.ones(s.shape[0]) / s.shape[0], U_loc))
    A = CollectiveOperator(ops[0], NullCollective(), "avg") if ranks == 1 else SimulatedRanksOperator(ops, "avg")
    d, U = hp.doublePass(A, hp.MultiVector.from_dense(Omega), rank, s=1)
    return d, U.to_dense()


# ----------------------------------------------------------------------------- active subspace
def as_input_from_jacobians(J, rank, Omega, noise_cov_inv=None, B_csr=None, ranks=1):
    """Input active subspace from stored Jacobians: MeanJTJfromDataOperator
    (operatorWrappers.py:55-121) wrapped like activeSubspaceProjector.py:427-463:
    prior_preconditioned (B_csr given, stands for prior.R / prior.Rsolver) -> doublePassG + encoder = R decoder
    (:447-453); otherwise doublePass and encoder = copy of decoder (:461-463)."""
    if ranks == 1:
        A = CollectiveOperator(MeanJTJfromDataOperator(J, noise_cov_inv), NullCollective(), "avg")
    else:
        A = SimulatedRanksOperator([MeanJTJfromDataOperator(s, noise_cov_inv) for s in np.split(J, ranks)], "avg")
    Om = hp.MultiVector.from_dense(Omega)
    if B_csr is not None:
        B = hp.SparseOperator(B_csr)
        d, V = hp.doublePassG(A, B, B, Om, rank, s=1)
        enc = hp.MultiVector(V)
        hp.MatMvMult(B, V, enc)
    else:
        d, V = hp.doublePass(A, Om, rank, s=1)
        enc = hp.MultiVector(V)
    return d, V.to_dense(), enc.to_dense()


def _b_orthonormalize_blocked(Q, B_csr=None):
    """Two eigen-based Cholesky-QR sweeps in the B inner product (B = I when None) -- the blocked stand-in for
    hIPPYlib's MGS Borthogonalize; directions at round-off level are dropped, not amplified."""
    for _ in range(2):
        G = Q.T @ (Q if B_csr is None else B_csr @ Q)
        w, V = np.linalg.eigh(0.5 * (G + G.T))
        keep = w > w.max() * Q.shape[1] * np.finfo(float).eps
        Q = Q @ (V * np.where(keep, 1.0 / np.sqrt(np.where(keep, w, 1.0)), 0.0))
    return Q


def as_input_from_jacobians_blocked(J, rank, Omega, noise_cov_inv=None, B_csr=None):
    """Blocked (BLAS-3) evaluation of ``as_input_from_jacobians``: the operator mean_i J_i^T [G] J_i
    (operatorWrappers.py:95-114) applied to all columns at once through the stacked (N dQ, dM) matrix, the same
    doublePass / doublePassG structure (activeSubspaceProjector.py:447-463), T = Q^T A Q as a Gram matrix.  d and
    span(V) agree with the column-by-column port (checked in tests/test_oracle_golden.py); used at benchmark shapes
    where the column-by-column port would take hours."""
    N, dQ, dM = J.shape
    J2 = J.reshape(N * dQ, dM)

    def weighted(W):                                        # blockdiag(Gamma^-1) W, sample by sample
        if noise_cov_inv is None:
            return W
        G = np.asarray(noise_cov_inv)
        return np.einsum("ab,ibm->iam", G, W.reshape(N, dQ, -1)).reshape(N * dQ, -1)

    Y = J2.T @ weighted(J2 @ Omega) / N                     # A Omega
    if B_csr is not None:
        lu = spla.splu(B_csr.tocsc())
        Y = lu.solve(Y)                                     # B^-1 A Omega
    Q = _b_orthonormalize_blocked(Y, B_csr)
    W = J2 @ Q
    T = W.T @ weighted(W) / N                               # Q^T A Q
    dd, VV = np.linalg.eigh(0.5 * (T + T.T))
    d = dd[::-1][:rank]
    V = Q @ VV[:, ::-1][:, :rank]
    return d, V, (B_csr @ V if B_csr is not None else V.copy())


def pod_randomized_blocked(u_data, rank, Omega):
    """Blocked evaluation of ``pod_randomized`` (PODProjector.py:359-376 / KLE 'identity', KLEProjector.py:175-180):
    doublePass on C = X^T X / N, no weighting, no shift."""
    N = u_data.shape[0]
    Q = _b_orthonormalize_blocked(u_data.T @ (u_data @ Omega) / N)
    W = u_data @ Q
    dd, VV = np.linalg.eigh(W.T @ W / N)
    return dd[::-1][:rank], Q @ VV[:, ::-1][:, :rank]


def as_output_from_jacobians(J, rank, Omega):
    """Output subspace E[J J^T] (activeSubspaceProjector.py:625-673: JJT operators, 'avg', doublePass)."""
    N = J.shape[0]

    class _JJT:
        def init_vector(self, x, dim):
            x.init(J.shape[1])

        def mult(self, x, y):
            t = np.einsum("ijk,j->ik", J, x._a)
            y.set_local(np.einsum("ijk,ik->j", J, t) / N)

    d, U = hp.doublePass(CollectiveOperator(_JJT(), NullCollective(), "avg"), hp.MultiVector.from_dense(Omega), rank, s=1)
    return d, U.to_dense()


# ----------------------------------------------------------------------------- KLE
def kle_from_samples(m_data, M_csr, rank, Omega, orthogonality="mass", ranks=1):
    """KLEProjector.construct_input_subspace (KLEProjector.py:136-199) with the covariance given by
    stored parameter draws, C = m_data^T m_data / N:
    'mass'     -> doublePassG(M C M, M, Msolver), encoder = M decoder (:163-168);
    'identity' -> doublePass(C), encoder = copy of decoder (:175-180)."""
    if orthogonality.lower() == "mass":
        d, V, E, _ = pod_randomized_weighted(m_data, M_csr, rank, Omega, shifted=False, ranks=ranks)
        return d, V, E
    elif orthogonality.lower() == "identity":
        d, V = pod_randomized(m_data, rank, Omega, ranks=ranks)
        return d, V, V.copy()
    raise ValueError(orthogonality)


# ----------------------------------------------------------------------------- projection of stored data
def project_data(data, encoder):
    """Reduced coordinates of stored samples: row i -> encoder^T data_i, i.e. (M V)^T m_i
    (encoder = M decoder: KLEProjector.py:167-168, PODProjector.py:769,830)."""
    return data @ encoder


def jstar_phi(J, MPhi):
    """JstarPhi_i = J_i^T (M Phi), stacked (N, dM, rQ)  (dataGenerator.py:170,339,582)."""
    return np.matmul(np.swapaxes(J, 1, 2), MPhi)                   # (N, dM, dQ) @ (dQ, rQ): BLAS, not a naive einsum loop


def j_psi(J, Psi):
    """JPsi_i = J_i Psi, stacked (N, dQ, rM)  (dataGenerator.py:177,585)."""
    return np.matmul(J, Psi)


def reduced_jacobians(J, PhiEnc, V):
    """Phi_enc^T J_i V, stacked (N, rQ, rM) (north star: 'Phi^T J V over all samples')."""
    return np.matmul(PhiEnc.T, np.matmul(J, V))                    # Phi^T (J_i V) per sample


# ----------------------------------------------------------------------------- comparison metrics
def principal_angle(U, V, M=None):
    """Largest principal angle between span(U) and span(V) in the M inner product."""
    def orth(A):
        G = A.T @ (A if M is None else M @ A)
        L = np.linalg.cholesky((G + G.T) / 2)
        return np.linalg.solve(L, A.T).T
    Uo, Vo = orth(U), orth(V)
    # sin(theta_max) = || (I - Uo Uo^T M) Vo ||_2 in the M-norm
    R = Vo - Uo @ (Uo.T @ (Vo if M is None else M @ Vo))
    G = R.T @ (R if M is None else M @ R)
    s = np.sqrt(max(np.linalg.eigvalsh((G + G.T) / 2).max(), 0.0))
    return float(np.arcsin(min(s, 1.0)))
