"""NumPy restatement of the hIPPYlib pieces the hot path calls.  TEST INFRASTRUCTURE ONLY.

hIPPYlib is a third-party dependency of the reference that is NOT under /root/reference and is not
installed here; it is pinned only as branch ``matmvmult`` of github.com/hippylib/hippylib
(/root/reference/.travis.yml:15).  This file restates, from the published algorithms, the routines the
reference calls on the path:

* ``doublePass``   -- Halko, Martinsson, Tropp, "Finding structure with randomness" (2011), Alg. 5.3
                      (two-pass symmetric eigensolver), as called at
                      hippyflow/modeling/PODProjector.py:376, activeSubspaceProjector.py:461,568,577,654,
                      KLEProjector.py:177.
* ``doublePassG``  -- Saibaba, Lee, Kitanidis, "Randomized algorithms for generalized Hermitian
                      eigenvalue problems" (2016), two-pass algorithm, as called at
                      activeSubspaceProjector.py:449,455,556,562 and KLEProjector.py:163.
* ``MultiVector``  -- dense collection of vectors with dot_v / dot_mv / reduce / (B-)orthogonalize
                      (modified Gram-Schmidt with re-orthogonalisation).
* ``LowRankOperator``, ``Solver2Operator``, ``MatMvMult``, ``MatMvTranspmult``, ``MvDSmatMult``.

Parity status: "parity unpinned" at the hIPPYlib boundary -- the reference holds no golden vectors
for these routines; what pins this file is (i) the property tests the reference's
own unit tests state (orthogonality 1e-10, eigen-residuals, equal eigenvalues under two operator
orders to 1e-12: hippyflow/test/test_KLEProjector.py:91-129, test_derivativeSubspace.py:83-102) and
(ii) agreement with the reference's deterministic solvers run verbatim (tests/golden/).
"""
import numpy as np


class Vector:
    """Minimal dolfin.Vector stand-in: the handful of methods the reference's operators touch
    (get_local/set_local/axpy/zero/inner/init; hippyflow/modeling/operatorWrappers.py:99,114,
    collectives/collective.py:102-105)."""

    def __init__(self, arr=None):
        if isinstance(arr, Vector):
            self._a = arr._a.copy()
        elif arr is None:
            self._a = np.zeros(0)
        else:
            self._a = arr

    def init(self, n):
        self._a = np.zeros(int(n))

    def get_local(self):
        return self._a.copy()

    def set_local(self, v):
        self._a[:] = v

    def apply(self, mode=""):
        pass

    def zero(self):
        self._a[:] = 0.0

    def axpy(self, alpha, x):
        self._a += alpha * x._a

    def inner(self, x):
        return float(self._a @ x._a)

    def norm(self, kind="l2"):
        return float(np.linalg.norm(self._a))

    def size(self):
        return self._a.shape[0]

    def copy(self):
        return Vector(self._a.copy())

    def __imul__(self, alpha):
        self._a *= alpha
        return self


class MultiVector:
    """k vectors of length n.  Storage is (k, n) C-order so that vector j is contiguous; the
    boundary layout of the reference is the transpose, dense (n, k) with column j = vector j
    (hippyflow/utilities/mv_utilities.py:31-49) -- see ``to_dense`` / ``from_dense``."""

    def __init__(self, arg, nvec=None):
        if isinstance(arg, MultiVector):
            self._d = arg._d.copy()
        elif isinstance(arg, Vector):
            self._d = np.zeros((int(nvec), arg.size()))
        else:
            raise TypeError("MultiVector(Vector, nvec) or MultiVector(MultiVector)")

    @staticmethod
    def from_dense(a):
        mv = MultiVector.__new__(MultiVector)
        mv._d = np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)
        return mv

    def to_dense(self):
        return np.ascontiguousarray(self._d.T)

    def nvec(self):
        return self._d.shape[0]

    def __getitem__(self, j):
        return Vector(self._d[j])

    def zero(self):
        self._d[:] = 0.0

    def swap(self, other):
        self._d, other._d = other._d, self._d

    def dot_v(self, v):
        return self._d @ v._a

    def dot_mv(self, mv):
        # entry (i, j) = <self[i], mv[j]>
        return self._d @ mv._d.T

    def reduce(self, y, alpha):
        y._a += np.asarray(alpha) @ self._d

    def scale(self, j, alpha):
        self._d[j] *= alpha

    def _mgs(self, Bmult):
        """Modified Gram-Schmidt with re-orthogonalisation in the inner product defined by
        ``Bmult`` (None = Euclidean).  Statement of the scheme in SURVEY.md section 3.7: for column
        k, t = ||q_k||_B; repeat { q_k -= sum_{i<k} (q_i^T B q_k) q_i sequentially; tt = ||q_k||_B }
        while tt > 10 eps t and tt < t/10 (then t = tt); a column with tt < 10 eps t is zeroed."""
        k_tot = self.nvec()
        n = self._d.shape[1]
        eps = np.finfo(np.float64).eps
        Bq = np.zeros_like(self._d)
        tmp = Vector(np.zeros(n))

        def apply_B(j):
            if Bmult is None:
                Bq[j] = self._d[j]
            else:
                tmp.zero()
                Bmult(Vector(self._d[j]), tmp)
                Bq[j] = tmp._a

        for k in range(k_tot):
            apply_B(k)
            t = np.sqrt(max(Bq[k] @ self._d[k], 0.0))
            again = True
            while again:
                for i in range(k):
                    s = Bq[i] @ self._d[k]
                    self._d[k] -= s * self._d[i]
                apply_B(k)
                tt = np.sqrt(max(Bq[k] @ self._d[k], 0.0))
                if tt > 10.0 * eps * t and tt < t / 10.0:
                    t = tt
                else:
                    again = False
                    if tt < 10.0 * eps * t:
                        tt = 0.0
            inv = 1.0 / tt if abs(tt) > eps else 0.0
            self._d[k] *= inv
            Bq[k] *= inv

    def orthogonalize(self):
        self._mgs(None)

    def Borthogonalize(self, B):
        self._mgs(B.mult)


def MatMvMult(A, X, Y):
    """Y[j] = A X[j]; uses A.matMvMult when the operator has one, else a column loop
    (this is the dispatch hippyflow/collectives/collectiveOperator.py:68-80 relies on)."""
    if hasattr(A, "matMvMult"):
        A.matMvMult(X, Y)
    else:
        for j in range(X.nvec()):
            A.mult(X[j], Y[j])


def MatMvTranspmult(A, X, Y):
    if hasattr(A, "matMvTranspmult"):
        A.matMvTranspmult(X, Y)
    else:
        for j in range(X.nvec()):
            A.transpmult(X[j], Y[j])


def MvDSmatMult(X, A, Y):
    """Y = X A with X a MultiVector (n x k) and A dense (k x l)."""
    Y._d[:] = np.asarray(A).T @ X._d


class LowRankOperator:
    """y = U diag(d) U^T x  (hippyflow/modeling/PODProjector.py:360 builds it with d = 1/N_loc)."""

    def __init__(self, d, U, init_vector=None):
        self.d = np.asarray(d, dtype=np.float64)
        self.U = U
        self._init_vector = init_vector

    def init_vector(self, x, dim):
        if self._init_vector is not None:
            self._init_vector(x, dim)
        else:
            x.init(self.U._d.shape[1])

    def mult(self, x, y):
        t = self.U.dot_v(x)
        y.zero()
        self.U.reduce(y, self.d * t)

    transpmult = mult


class Solver2Operator:
    """Wrap an object with solve(y, x) as an operator (hippyflow/modeling/KLEProjector.py:103)."""

    def __init__(self, S, mpi_comm=None):
        self.S = S

    def init_vector(self, x, dim):
        self.S.init_vector(x, dim)

    def mult(self, x, y):
        self.S.solve(y, x)


class SparseOperator:
    """SciPy sparse matrix as an operator/solver pair (stands in for dolfin matrices such as
    prior.M / prior.Msolver, hippyflow/modeling/KLEProjector.py:163-168)."""

    def __init__(self, M):
        self.M = M.tocsr()
        self._lu = None

    def init_vector(self, x, dim):
        x.init(self.M.shape[0])

    def mult(self, x, y):
        y.set_local(self.M @ x._a)

    transpmult = mult

    def solve(self, y, x):
        if self._lu is None:
            import scipy.sparse.linalg as spla
            self._lu = spla.splu(self.M.tocsc())
        y.set_local(self._lu.solve(x._a))


def _top_k(T, k):
    d, V = np.linalg.eigh(T)
    perm = np.argsort(d)[::-1][:k]
    return d[perm], V[:, perm]


def doublePass(A, Omega, k, s=1):
    """Two-pass randomized eigensolver for symmetric A (HMT 2011):
    Q = orth(A^s Omega); T = (AQ)^T Q; eigh(T); top-k descending; U = Q V."""
    nvec = Omega.nvec()
    assert k <= nvec
    Q = MultiVector(Omega)
    Y = MultiVector(Omega)
    for _ in range(s):
        Y.zero()
        MatMvMult(A, Q, Y)
        Q.swap(Y)
    Q.orthogonalize()
    AQ = MultiVector(Omega)
    AQ.zero()
    MatMvMult(A, Q, AQ)
    T = AQ.dot_mv(Q)
    d, V = _top_k(T, k)
    U = MultiVector(Omega[0], k)
    MvDSmatMult(Q, V, U)
    return d, U


def doublePassG(A, B, Binv, Omega, k, s=1):
    """Two-pass randomized solver for A u = lambda B u (Saibaba-Lee-Kitanidis 2016):
    Ybar = A Omega; Q = Binv Ybar (s times); B-orthonormalise Q; T = (AQ)^T Q; eigh; U = Q V
    with U^T B U = I."""
    nvec = Omega.nvec()
    assert k <= nvec
    Ybar = MultiVector(Omega)
    Q = MultiVector(Omega)
    Bi = Solver2Operator(Binv)
    for _ in range(s):
        Ybar.zero()
        MatMvMult(A, Q, Ybar)
        MatMvMult(Bi, Ybar, Q)
    Q.Borthogonalize(B)
    AQ = MultiVector(Omega)
    AQ.zero()
    MatMvMult(A, Q, AQ)
    T = AQ.dot_mv(Q)
    d, V = _top_k(T, k)
    U = MultiVector(Omega[0], k)
    MvDSmatMult(Q, V, U)
    return d, U


def accuracyEnhancedSVD(A, Omega, k, s=1):
    """hIPPYlib ``accuracyEnhancedSVD`` (randomizedSVD.py of the pinned hIPPYlib branch; source absent from
    /root/reference -- restated from the published algorithm, Halko/Martinsson/Tropp 2011 Alg. 4.4 + 5.1, as hIPPYlib
    documents it; call sites activeSubspaceProjector.py:816,1026, dataGenerator.py:187 with s = 1).  PARITY UNPINNED at this
    boundary: the reference holds no golden vectors for it; the tests pin it against numpy.linalg.svd instead.

        Y = A Omega;  repeat s times: Z = A^T Y, Y = A Z;  Q = orth(Y) (MGS);
        B^T = A^T Q, orthogonalised in place with R returned: B^T = Q~ R;   R = V^ diag(d) U^^T
        U = Q U^[:, :k],  V = Q~ V^[:, :k],  d[:k].
    ``A`` exposes mult / transpmult / init_vector (dim 0 = range, 1 = domain); Omega is a MultiVector in the domain."""
    nvec = Omega.nvec()
    assert nvec >= k
    y = Vector(np.zeros(0))
    A.init_vector(y, 0)
    Y = MultiVector(y, nvec)
    MatMvMult(A, Omega, Y)
    Z = MultiVector(Omega)
    for _ in range(s):
        MatMvTranspmult(A, Y, Z)
        MatMvMult(A, Z, Y)
    Q = MultiVector(Y)
    Q.orthogonalize()
    BT = MultiVector(Omega)
    MatMvTranspmult(A, Q, BT)
    Bt = BT.to_dense()                                    # (dM, l)
    BT.orthogonalize()
    Qt = BT.to_dense()
    R = Qt.T @ Bt                                         # the triangular factor MGS returns (B^T = Q~ R)
    V_hat, d, U_hatT = np.linalg.svd(R, full_matrices=False)
    U = Q.to_dense() @ U_hatT.T[:, :k]
    V = Qt @ V_hat[:, :k]
    return U, d[:k], V


class DenseOperator:
    """A stored (dQ, dM) Jacobian as a hIPPYlib-style operator (mult = J x, transpmult = J^T y)."""

    def __init__(self, J):
        self.J = np.asarray(J)

    def init_vector(self, x, dim):
        x.init(self.J.shape[0] if dim == 0 else self.J.shape[1])

    def mult(self, x, y):
        y.set_local(self.J @ x.get_local())

    def transpmult(self, x, y):
        y.set_local(self.J.T @ x.get_local())
