"""Sample-averaged operators on the device (layer L3 of SURVEY.md section 1): the duck-typed hIPPYlib
operator protocol -- ``mult(x, y)``, ``transpmult(x, y)``, ``init_vector(x, dim)``, ``matMvMult(X, Y)`` --
implemented with the two-GEMM block apply of ``linalg.SampleCovariance``."""
import numpy as np
import torch

from .. import _lib as K
from ..collectives import NullCollective
from ..linalg import CsrMatrix, SampleCovariance
from ..multivector import DeviceMultiVector, DeviceVector


def _as_device_rows(data, device):
    """(rows, n) float64 array -> padded row-major device block (zero-copy if already conforming)."""
    if K.is_device_tensor(data) and data.dtype == torch.float64 and data.dim() == 2 \
            and data.stride(1) == 1 and K._ld(data) % 2 == 0 and data.data_ptr() % 16 == 0:
        return data
    return K.to_padded(data, device)


class SampleCovarianceOperator:
    """A = avg over ranks of (1/N_loc) sum_i X_i G X_i^T.

    Equivalent of ``CollectiveOperator(hp.LowRankOperator(ones/N_loc, U_loc), collective, 'avg')``
    (PODProjector.py:360-363) and of ``CollectiveOperator(SummedListOperator([JTJ(J_i)]), collective,
    'avg')`` (activeSubspaceProjector.py:427-431): local mean, then allReduce with mpi_op.  The averaging
    assumes every rank holds the same number of samples (comment at activeSubspaceProjector.py:429-430)."""

    overwrites = True   # matMvMult writes every entry of Y (no caller-side zeroing needed)

    def __init__(self, cov, collective=None, mpi_op="avg"):
        self.cov = cov
        self.collective = collective if collective is not None else NullCollective()
        self.mpi_op = mpi_op
        self.n = cov.n
        self.device = cov.Xt.device

    def init_vector(self, x, dim):
        x.init(self.n)

    def matMvMult(self, X, Y):
        _, GW = self.cov.project(X.tensor())
        self.lift_reduced(GW, Y)

    matMvTranspmult = matMvMult

    LIFT_CHUNKS = int(__import__("os").environ.get("HFB_LIFT_CHUNKS", 4))

    def lift_reduced(self, GW, Y, nchunk=None):
        """Y = allReduce_op( (1/N_loc) Xt^T GW ).  With more than one rank the lift GEMM is cut into row blocks of Y and
        the NCCL allreduce of each block is issued asynchronously as soon as its GEMM is queued, so the exchange of
        block i overlaps the GEMM of block i+1 (the 'avg' factor 1/size is folded into the GEMM's alpha)."""
        Yt = Y.tensor()
        n = Yt.shape[0]
        nchunk = self.LIFT_CHUNKS if nchunk is None else nchunk
        scale = 1.0 / self.cov.nsamples
        size = self.collective.size()
        lazy = self.cov.lazy_pending()
        if lazy and (not self.cov.can_lazy(Yt, GW.shape[1]) or self.mpi_op.lower() != "avg"):
            # no spare column in the result block: determine the mean the ordinary way first
            c = K.colsum(self.cov.Xt, 1.0 / self.cov.nsamples)
            self.collective.allReduce(c, "avg")
            self.cov.center = c
            W, GW = self.cov.project(self.cov._lazy_B)
            self.cov._lazy_B = None
            lazy = False
        if size == 1 or not hasattr(self.collective, "allReduce_async") or n < 4096:
            self.cov.lift(GW, out=Yt, scale=scale, weighted=True)
            self.collective.allReduce(Y, self.mpi_op)       # reduces the whole padded block: the mean column travels along
            if lazy:
                self.cov.finish_lazy(Yt, self.collective if size > 1 else None, self.mpi_op)
            return
        if self.mpi_op.lower() == "avg":
            scale /= float(size)
        elif self.mpi_op.lower() != "sum":
            raise NotImplementedError("Unknown operation *{0}*".format(self.mpi_op))
        if Yt.data_ptr() % 16 == 0 and K._ld(Yt) % 2 == 0 and self._lift_peer(GW, Y, scale, lazy):
            return
        self._lift_nccl(GW, Y, Yt, scale, lazy, nchunk)

    def _lift_nccl(self, GW, Y, Yt, scale, lazy, nchunk):
        """Row blocks of the lift, each followed by an asynchronous NCCL allreduce (``scale`` carries the 1/size of 'avg')."""
        n = Yt.shape[0]
        step = max(128, ((n + nchunk - 1) // nchunk + 127) // 128 * 128)     # 128-row multiples keep TMA alignment
        full = Y.storage_tensor()
        works = []
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            self.cov.lift(GW, out=Yt[lo:hi], scale=scale, weighted=True, rows=(lo, hi))
            works.append(self.collective.allReduce_async(full[lo:hi], "sum"))
        for w in works:
            w.wait()
        if lazy:
            # the chunks carried scale / size and were summed: the extra column is the global mean already; 1^T W / N is
            # averaged inside finish_lazy
            self.cov.finish_lazy(Yt, self.collective, "avg")

    PEER_LIFT = __import__("os").environ.get("HFB_PEER_LIFT", "1") != "0"
    PEER_VERIFY_TOL = 1e-11

    def _peer_exchange(self, n, ld, ncols):
        """The collective's NVLink exchange buffers for this block shape (created collectively on first use; None when the
        ranks cannot map each other's memory -- then every rank takes the NCCL route)."""
        if not self.PEER_LIFT or getattr(self.collective, "_backend", None) != "nccl":
            return None
        cache = self.collective.__dict__.setdefault("_peer_exchanges", {})
        key = (n, ld, ncols)
        if key not in cache:
            from ..peer import PeerExchange, release_all
            if len(cache) >= 4:                                   # bound the device memory held by stale shapes
                release_all([e for e in cache.values() if e is not None], self.collective.group)
                cache.clear()
            cache[key] = PeerExchange.create(self.collective.group, self.device, n, ld, ncols)
        return cache[key]

    def _lift_peer(self, GW, Y, scale, lazy):
        """The lift with its allreduce fused into the GEMM epilogue over NVLink peer memory (hippyflow_b200/peer.py).
        Returns False when the route is unavailable.  The reduced sketch arrives in the exchange buffer itself: a block that
        the solver allocated as scratch (``Y.adoptable``) is re-pointed at it, any other block receives a copy.  The first
        exchange of every buffer set is checked against the NCCL route on the same operands; a mismatch disables the peer
        route on all ranks (with a warning)."""
        cov = self.cov
        Yt = Y.tensor()
        n, m = Yt.shape[0], GW.shape[1]
        ld = K._ld(Yt)
        if lazy:
            Wop, ncols = cov._Wext, m + 1
            if GW.data_ptr() != cov._Wext.data_ptr() or ld < m + 1:
                return False
        else:
            Wop, ncols = GW, m
        ex = self._peer_exchange(n, ld, ncols)
        if ex is None:
            return False
        R = ex.lift_allreduce(cov.Xt, Wop, scale)                  # (n, ncols) view of the exchange buffer, leading dimension ld
        if not getattr(ex, "verified", False):
            ref = K.dgemm(K.HFB_TN, cov.Xt, Wop, alpha=scale).contiguous()
            self.collective.allReduce(ref, "sum")
            err = float((ref - R).abs().max() / ref.abs().max().clamp_min(1e-300))
            bad = torch.tensor([0 if err < self.PEER_VERIFY_TOL else 1], dtype=torch.int32, device=self.device)
            self.collective.allReduce(bad, "sum")
            ex.verified, ex.verify_err = True, err
            if int(bad.item()):
                import warnings
                from ..peer import release_all
                warnings.warn("hippyflow_b200: NVLink peer exchange disagrees with the NCCL allreduce (rel. error %.2e); "
                              "peer route disabled" % err)
                self.collective._peer_exchanges[(n, ld, ncols)] = None
                Yt.as_strided((n, ncols), (ld, 1)).copy_(ref)
                release_all([ex], self.collective.group)
                R = None
        if R is not None:
            if getattr(Y, "adoptable", False) and Y.tensor().shape[1] == m:
                Y.adopt(R[:, :m])
            else:
                Yt.as_strided((n, ncols), (ld, 1)).copy_(R)
        Yt = Y.tensor()
        if lazy:
            cov.finish_lazy(Yt, self.collective, "avg")
        elif cov.center is not None:
            cw = K.colsum(GW, 1.0)                                 # - c (1^T W) with the GLOBAL sum of the projections
            self.collective.allReduce(cw, "sum")
            K.rank1_update_(Yt, -scale, cov.center, cw)
        return True

    def mult(self, x, y):
        self.cov.apply(x.storage_tensor(), out=y.storage_tensor())
        self.collective.allReduce(y, self.mpi_op)

    transpmult = mult

    def rayleigh(self, Q, BQ):
        """T = Q^T A Q (m x m, host) as a Gram matrix of the projected samples; only an (m x m) allreduce.
        (BQ is unused: A = C does not involve B.)"""
        return self.rayleigh_device(Q, BQ).cpu().numpy()

    def rayleigh_device(self, Q, BQ):
        """Same, left on the device (asynchronous)."""
        T = self.cov.gram_T(Q.tensor())
        self.collective.allReduce(T, self.mpi_op)
        return T


class SandwichedCovarianceOperator:
    """A = B C B with C a SampleCovarianceOperator and B a sparse SPD matrix on the device:
    ``MassPreconditionedCovarianceOperator`` (KLEProjector.py:47-69, M C M) and ``H_matvec`` of the
    weighted POD (PODProjector.py:750-754, MX (MX)^T / N)."""

    overwrites = True   # matMvMult writes every entry of Y (no caller-side zeroing needed)

    def __init__(self, C, B):
        self.C = C
        self.B = B
        self.n = C.n

    def init_vector(self, x, dim):
        x.init(self.n)

    def matMvMult(self, X, Y):
        BX = DeviceMultiVector(self.B.matmat(X.tensor()))
        CBX = DeviceMultiVector(self.n, X.nvec(), device=X.tensor().device)
        CBX.adoptable = True                    # scratch: the sample operator may hand back its exchange block instead
        self.C.matMvMult(BX, CBX)
        self.B.matmat(CBX.tensor(), out=Y.tensor())

    def mult(self, x, y):
        X = DeviceMultiVector(x.storage_tensor())
        Y = DeviceMultiVector(y.storage_tensor())
        self.matMvMult(X, Y)

    transpmult = mult

    def solveB_matMvMult(self, X, Y):
        """Y = B^-1 A X = C (B X): the generalized range finder needs no solve with B."""
        BX = DeviceMultiVector(self.B.matmat(X.tensor()))
        self.C.matMvMult(BX, Y)

    def rayleigh(self, Q, BQ):
        return self.C.rayleigh(BQ, None)   # Q^T (B C B) Q = (BQ)^T C (BQ)

    def rayleigh_device(self, Q, BQ):
        return self.C.rayleigh_device(BQ, None)


class MeanJTJfromDataOperator:
    """Drop-in for hippyflow/modeling/operatorWrappers.py:55-121: y = mean_i J_i^T [Gamma^-1] J_i x from a
    stored (ndata, r, dM) array.  The reference reads all of J twice per column through two einsums; here
    the array sits in HBM as the (ndata*r, dM) row-major matrix and a block of columns costs two GEMMs."""

    overwrites = True   # matMvMult writes every entry of Y (no caller-side zeroing needed)

    def __init__(self, J, prior=None, noise_cov_inv=None, device=None, collective=None, mpi_op="avg"):
        assert len(J.shape) == 3
        self._J = J
        self.ndata, self.r, self.dM = J.shape
        self._prior = prior
        if noise_cov_inv is not None:
            assert hasattr(noise_cov_inv, "__matmul__")
        self._noise_cov_inv = noise_cov_inv
        if device is None:
            device = J.device if K.is_device_tensor(J) else torch.device("cuda", torch.cuda.current_device())
        J2 = J.reshape(self.ndata * self.r, self.dM)
        self._cov = SampleCovariance(_as_device_rows(J2, device), block=self.r,
                                     noise_cov_inv=None if noise_cov_inv is None else np.asarray(noise_cov_inv))
        self._op = SampleCovarianceOperator(self._cov, collective, mpi_op)
        self.device = device

    @property
    def J(self):
        return self._J

    @property
    def prior(self):
        return self._prior

    @property
    def noise_cov_inv(self):
        return self._noise_cov_inv

    def init_vector(self, x, dim):
        # the reference delegates to prior.R / prior.Hlr and trips over an undefined name
        # (operatorWrappers.py:92); domain and range both have dimension dM
        x.init(self.dM)

    def mult(self, x, y):
        if isinstance(x, DeviceVector):
            self._op.mult(x, y)
        else:  # host vectors with get_local / set_local, as in the reference
            xd = DeviceVector(self.dM, self.device)
            yd = DeviceVector(self.dM, self.device)
            xd.set_local(x.get_local())
            self._op.mult(xd, yd)
            y.set_local(yd.get_local())

    def transpmult(self, x, y):
        return self.mult(x, y)

    def matMvMult(self, X, Y):
        self._op.matMvMult(X, Y)

    def rayleigh(self, Q, BQ):
        return self._op.rayleigh(Q, BQ)

    def rayleigh_device(self, Q, BQ):
        return self._op.rayleigh_device(Q, BQ)


class MeanJJTfromDataOperator:
    """y = mean_i J_i J_i^T x on the OUTPUT space from a stored (ndata, dQ, dM) array: the operator
    ``CollectiveOperator(SummedListOperator([JJT(J_i)], average=True), collective, 'avg')`` of the output active subspace
    (activeSubspaceProjector.py:625-673, jacobian.py:169-193) in stored-data form.  It never forms the (dQ x dQ) matrix, so it
    also serves the full-state observable where dQ = n_u (fullStateObservable.py:18).  A block of m columns costs two
    strided-batch DMMA launches per chunk of samples:
        T_i = J_i^T X          (dM x m) for every sample of the chunk      hfb_dgemm_batched, independent outputs
        Y  += sum_i J_i T_i    (dQ x m), K loop folded over (sample, dM)   hfb_dgemm_batched, HFB_BATCH_REDUCE
    (the chunk bounds the (chunk, dM, m) workspace; it is not a per-sample loop)."""

    overwrites = True

    def __init__(self, J, device=None, collective=None, mpi_op="avg", chunk_bytes=4 << 30, stacked=None):
        """``stacked``: an existing (J2, J3) pair from ``projection.stacked_jacobians`` (shares the device copy another
        operator already made)."""
        from .projection import stacked_jacobians
        assert len(J.shape) == 3
        self.ndata, self.dQ, self.dM = J.shape
        if device is None:
            device = J.device if K.is_device_tensor(J) else torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._J2, self._J3 = stacked if stacked is not None else stacked_jacobians(J, device)
        self.collective = collective if collective is not None else NullCollective()
        self.mpi_op = mpi_op
        self.chunk_bytes = int(chunk_bytes)

    def init_vector(self, x, dim=None):
        x.init(self.dQ)

    def _apply_local(self, Xt, Yt):
        m = Xt.shape[1]
        ldm = ((m + 15) // 16) * 16
        chunk = int(max(1, min(self.ndata, self.chunk_bytes // max(1, self.dM * ldm * 8))))
        T = K.batched_empty(chunk, self.dM, m, self.device)
        for i0 in range(0, self.ndata, chunk):
            i1 = min(self.ndata, i0 + chunk)
            Jc = self._J3[i0:i1]
            Tc = T[:i1 - i0]
            K.dgemm_batched(K.HFB_TN, Jc, Xt, out=Tc)                                        # T_i = J_i^T X
            K.dgemm_batched(K.HFB_NN, Jc, Tc, out=Yt, alpha=1.0 / self.ndata, reduce=True, accumulate=(i0 > 0))
        return Yt

    def matMvMult(self, X, Y):
        self._apply_local(_tma_operand(X.tensor()), Y.tensor())
        self.collective.allReduce(Y, self.mpi_op)

    matMvTranspmult = matMvMult

    def mult(self, x, y):
        xt = _tma_operand(x.storage_tensor())
        yt = K.padded_empty(self.dQ, 1, self.device, pad=2)
        self._apply_local(xt, yt)
        y.storage_tensor().copy_(yt)
        self.collective.allReduce(y, self.mpi_op)

    transpmult = mult

    def dense(self):
        """C = mean_i J_i J_i^T as a (dQ x dQ) device matrix: ONE strided-batch launch with the K loop folded over
        (sample, dM) -- the cheap route when dQ is small (pointwise observations)."""
        C = K.dgemm_batched(K.HFB_NT, self._J3, self._J3, alpha=1.0 / self.ndata, reduce=True)
        self.collective.allReduce(C, self.mpi_op)
        return C


class JTJ:
    """Gauss-Newton Hessian J^T J of ONE stored Jacobian (dQ, dM): the stored-data form of
    hippyflow/modeling/jacobian.py:142-166 (there every apply is an incremental forward and adjoint PDE solve)."""

    overwrites = True   # matMvMult writes every entry of Y (no caller-side zeroing needed)

    def __init__(self, J, device=None):
        if device is None:
            device = J.device if K.is_device_tensor(J) else torch.device("cuda", torch.cuda.current_device())
        self._cov = SampleCovariance(_as_device_rows(J, device), block=J.shape[0])
        self.dM = J.shape[1]

    def init_vector(self, x, dim=0):
        x.init(self.dM)

    def matMvMult(self, X, Y):
        self._cov.apply(X.tensor(), out=Y.tensor(), scale=1.0)

    def mult(self, x, y):
        self._cov.apply(x.storage_tensor(), out=y.storage_tensor(), scale=1.0)

    transpmult = mult


class SummedListOperator:
    """hippyflow/modeling/activeSubspaceProjector.py:69-95: sum (or average) of a list of operators of equal
    dimension.  Works on device vectors and, block-wise, on device multivectors.  One deliberate difference: the
    reference seeds its accumulator with a copy of ``y`` (``temp = dl.Vector(y)``, :84-85), so whatever ``y`` holds on
    entry is added to the sum; here the accumulator starts from zero.  Both agree for the zero-initialised ``y`` every
    caller on the path passes (hIPPYlib's MatMvMult, activeSubspaceProjector.py:214-221)."""

    def __init__(self, operators, communicator=None, average=True):
        assert type(operators) is list
        self.operators = operators
        self.average = average

    def init_vector(self, x, dim=0):
        self.operators[0].init_vector(x, dim)

    def _accumulate(self, apply, x, y, make_tmp):
        acc = make_tmp()
        for op in self.operators:
            apply(op, x, y)
            acc.axpy(1.0, y)
        y.zero()
        y.axpy(1.0 / float(len(self.operators)) if self.average else 1.0, acc)

    def mult(self, x, y):
        self._accumulate(lambda op, a, b: op.mult(a, b), x, y, lambda: DeviceVector(y.size(), y.storage_tensor().device))

    def matMvMult(self, X, Y):
        self._accumulate(lambda op, a, b: op.matMvMult(a, b), X, Y,
                         lambda: DeviceMultiVector(Y.tensor().shape[0], Y.nvec(), device=Y.tensor().device))


def _tma_operand(t):
    """A GEMM operand must start on a 16-byte boundary with an even leading dimension (TMA); a column view of a
    multivector block (``mv[j]``) may not -- stage it into an aligned buffer then."""
    if t.data_ptr() % 16 or K._ld(t) % 2:
        return K.to_padded(t, t.device, pad=2 if t.shape[1] == 1 else 16)
    return t


class JJT:
    """J J^T of ONE stored Jacobian (dQ, dM): the stored-data form of hippyflow/modeling/jacobian.py:169-193
    (``J.transpmult`` then ``J.mult``, each a PDE solve there, each one DMMA GEMM here).  Output-space operator, used by
    the output active subspace E[J J^T] (activeSubspaceProjector.py:576-583)."""

    overwrites = True

    def __init__(self, J, device=None):
        if device is None:
            device = J.device if K.is_device_tensor(J) else torch.device("cuda", torch.cuda.current_device())
        self._J = _as_device_rows(J, device)                      # (dQ, dM)
        self.dQ, self.dM = self._J.shape

    def init_vector(self, x, dim=None):
        x.init(self.dQ)

    def matMvMult(self, X, Y):
        W = K.dgemm(K.HFB_TN, self._J, _tma_operand(X.tensor()))  # J^T X   (dM, m)
        K.dgemm(K.HFB_NN, self._J, W, out=Y.tensor())             # J (J^T X)

    def mult(self, x, y):
        W = K.dgemm(K.HFB_TN, self._J, _tma_operand(x.storage_tensor()))
        K.dgemm(K.HFB_NN, self._J, W, out=y.storage_tensor())

    transpmult = mult


class npToDolfinOperator:
    """Dense matrix as an operator (hippyflow/modeling/operatorWrappers.py:19-52; the name is the reference's): the
    matrix lives in HBM, ``mult`` / ``transpmult`` are one GEMM each on device vectors or multivector blocks."""

    overwrites = True

    def __init__(self, npArray, device=None):
        assert len(npArray.shape) == 2
        if device is None:
            device = npArray.device if K.is_device_tensor(npArray) else torch.device("cuda", torch.cuda.current_device())
        self.matrix = _as_device_rows(npArray, device)
        self.domain_help = None
        self.range_help = None

    def init_vector(self, x, dim):
        if dim == 0:
            x.init(self.matrix.shape[0])
        elif dim == 1:
            x.init(self.matrix.shape[1])
        else:
            raise ValueError("dim must be 0 or 1")

    def mult(self, x, y):
        K.dgemm(K.HFB_NN, self.matrix, _tma_operand(x.storage_tensor()), out=y.storage_tensor())

    def transpmult(self, x, y):
        K.dgemm(K.HFB_TN, self.matrix, _tma_operand(x.storage_tensor()), out=y.storage_tensor())

    def matMvMult(self, X, Y):
        K.dgemm(K.HFB_NN, self.matrix, _tma_operand(X.tensor()), out=Y.tensor())

    def matMvTranspmult(self, X, Y):
        K.dgemm(K.HFB_TN, self.matrix, _tma_operand(X.tensor()), out=Y.tensor())


class LowRankRectangularOperator:
    """A = U diag(s) V^T with U (dQ, r), V (dM, r) device multivectors and s (r,) -- hippyflow/modeling/
    lowRankRectangularOperator.py:19-72, e.g. the truncated SVD of a stored Jacobian (``Jsvd_data.npz``).
    ``mult`` / ``transpmult`` follow the reference step by step (dot_v, elementwise scale, reduce into a zeroed y);
    ``matMvMult`` / ``matMvTranspmult`` do the same for a whole block with two GEMMs and a row scaling."""

    def __init__(self, U, s, V, U_init_vector=None, V_init_vector=None):
        self.U = U if isinstance(U, DeviceMultiVector) else DeviceMultiVector.from_dense(np.asarray(U), _default_dev())
        dev = self.U.tensor().device
        self.V = V if isinstance(V, DeviceMultiVector) else DeviceMultiVector.from_dense(np.asarray(V), dev)
        self.s = np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s, dtype=np.float64).ravel()
        assert self.U.nvec() == self.V.nvec() == self.s.size
        self._s_dev = torch.as_tensor(self.s, device=dev)
        self.U_init_vector = U_init_vector
        self.V_init_vector = V_init_vector

    def init_vector(self, x, dim):
        if dim == 0:
            if self.U_init_vector is not None:
                self.U_init_vector(x)
            else:
                x.init(self.U.tensor().shape[0])
        elif dim == 1:
            if self.V_init_vector is not None:
                self.V_init_vector(x)
            else:
                x.init(self.V.tensor().shape[0])
        else:
            raise ValueError('dim must be 0 or 1')

    def mult(self, x, y):
        """y = U s V^T x  (lowRankRectangularOperator.py:50-57)."""
        Vtx = self.V.dot_v(x)
        sVtx = self.s * Vtx
        y.zero()
        self.U.reduce(y, sVtx)

    def transpmult(self, x, y):
        """y = V s U^T x  (lowRankRectangularOperator.py:59-66)."""
        Utx = self.U.dot_v(x)
        sUtx = self.s * Utx
        y.zero()
        self.V.reduce(y, sUtx)

    def _block(self, left, right, X, Y):
        coef = K.dgemm(K.HFB_TN, right.tensor(), _tma_operand(X.tensor()))        # right^T X   (r, m)
        K.dgemm(K.HFB_NN, left.tensor(), K.rowscale(self._s_dev, coef), out=Y.tensor())

    def matMvMult(self, X, Y):
        self._block(self.U, self.V, X, Y)

    def matMvTranspmult(self, X, Y):
        self._block(self.V, self.U, X, Y)


def _default_dev():
    if not torch.cuda.is_available():
        raise RuntimeError("hippyflow_b200 needs a CUDA device: the hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())
