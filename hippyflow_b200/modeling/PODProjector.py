"""POD output bases on the device: drop-in counterparts of hippyflow/modeling/PODProjector.py.

``PODProjectorFromData`` keeps the reference signature and return values (PODProjector.py:666-852) and adds
``method='randomized'``: the M-weighted double-pass solve of the GHEP written at PODProjector.py:750-761
(A = M X X^T M / N, B = M).  ``PODProjector`` is the collective double-pass path of PODProjector.py:331-390
for snapshots that are already stored (the PDE solves at :343-357 are upstream).
"""
import time

import numpy as np
import scipy.sparse as sp
import torch

from .. import _lib as K
from ..collectives import NullCollective
from ..linalg import CsrMatrix, SampleCovariance
from ..multivector import DeviceMultiVector, mv_to_dense
from ..parameterList import ParameterList
from ..randomized import doublePass, doublePassG
from .operators import SampleCovarianceOperator, SandwichedCovarianceOperator, _as_device_rows


def PODParameterList():
    """PODProjector.py:35-49."""
    parameters = {}
    parameters['sample_per_process'] = [100, 'Number of samples per process']
    parameters['rank'] = [20, 'Rank of POD subspace']
    parameters['oversampling'] = [10, 'Oversampling parameter for randomized algorithms']
    parameters['data_per_process'] = [250, 'Total number of testing and training data to be constructed']
    parameters['verbose'] = [True, 'Boolean for prints']
    parameters['output_directory'] = [None, 'output directory for saving arrays and plots']
    parameters['plot_label_suffix'] = ['', 'suffix for plot label']
    parameters['save_and_plot'] = [True, 'save the projector arrays (plots are out of scope)']
    parameters['omega_seed'] = [1, 'seed of the Gaussian test matrix when none is supplied']
    return ParameterList(parameters)


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("hippyflow_b200 needs a CUDA device: the hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_scipy_csr(M):
    """Accept a SciPy sparse matrix, or a dolfin/PETSc matrix the way the reference does
    (``dl.as_backend_type(M).mat().getValuesCSR()``, PODProjector.py:695-697)."""
    if sp.issparse(M):
        return M.tocsr()
    if hasattr(M, "mat"):
        row, col, val = M.mat().getValuesCSR()
        return sp.csr_matrix((val, col, row))
    if hasattr(M, "getValuesCSR"):
        row, col, val = M.getValuesCSR()
        return sp.csr_matrix((val, col, row))
    raise TypeError("M_output must be a scipy.sparse matrix or expose getValuesCSR()")


def gaussian_omega(n, m, seed, device):
    """Gaussian test matrix (role of hp.parRandom.normal(1., Omega), PODProjector.py:367-372) generated on the
    device by the counter-based generator of hfb_fill_random: the same (seed, n, m) gives the same Omega on
    every rank.  Parity runs pass an explicit Omega instead so that oracle and GPU see identical values."""
    Om = DeviceMultiVector(K.padded_empty(n, m, device))
    K.fill_random_(Om.tensor(), seed)
    return Om


def _upload_chunks(N, max_chunks=8):
    """Row chunks [(lo, hi)] of a pipelined upload.  The GPU works on chunk i while chunk i+1 crosses PCIe, so what remains
    after the LAST byte has landed is the work on the last chunk: the chunks taper (weights 8 ... 8, 4, 2, 1) to make that
    tail small while the early chunks stay large enough for efficient GEMMs."""
    nchunk = int(max(1, min(max_chunks, N // 256)))
    if nchunk < 4:
        return [(i * N // nchunk, (i + 1) * N // nchunk) for i in range(nchunk)]
    w = [8] * (nchunk - 2) + [4, 2, 1]
    tot, acc, cuts = float(sum(w)), 0.0, [0]
    for x in w[:-1]:
        acc += x
        cuts.append(min(N, max(cuts[-1] + 8, int(round(N * acc / tot / 8.0)) * 8)))
    cuts.append(N)
    return [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


_copy_streams = {}


def _copy_stream(dev):
    st = _copy_streams.get(dev.index)
    if st is None:
        st = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    return st


_staging = {}


def _host_copy_threads():
    """Threads for the pageable -> pinned staging copies: the cores this process may run on, shared between the ranks of
    the node (LOCAL_WORLD_SIZE), at most 16."""
    import os
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return int(max(1, min(16, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))))


def _staging_ring(dev, n, nbuf=3, target_bytes=96 << 20):
    """Ring of pinned staging buffers (rows x n) for uploads from pageable host memory, cached per device and row length."""
    key = (dev.index, n)
    ring = _staging.get(key)
    if ring is None:
        rows = int(max(1, target_bytes // (8 * n)))
        ring = []
        for _ in range(nbuf):
            ev = torch.cuda.Event()
            ev.record()
            ring.append((torch.empty((rows, n), dtype=torch.float64, pin_memory=True), ev))
        _staging.clear()                                                     # one row length at a time: do not hoard pinned memory
        _staging[key] = ring
    return ring


def to_host(t):
    """Device block -> NumPy array through a pinned staging tensor (torch's caching host allocator reuses the
    pinned blocks once earlier results are garbage-collected)."""
    h = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def weighted_l2_norm_vector(x, W):
    """sqrt(diag(x^T W x)) per column (PODProjector.py:658-661); x a device block, W a CsrMatrix."""
    Wx = W.matmat(x)
    return torch.sqrt(K.coldot(Wx, x))


class PODProjectorFromData:
    """M-weighted POD from a stored (n_data, dim_u) snapshot array."""

    def __init__(self, Vh=None, M_output=None, device=None):
        """``Vh`` is kept for signature compatibility (PODProjector.py:670); the weighting matrix must be
        given as ``M_output`` (SciPy CSR, or a PETSc-backed matrix) because FEniCS assembly is upstream."""
        self.Vh = Vh
        if M_output is None:
            raise ValueError("M_output is required: assembling the mass matrix from Vh needs FEniCS (upstream)")
        self.M = M_output
        self.M_csr = _to_scipy_csr(M_output)
        self.device = device if device is not None else _default_device()
        self._Md = None
        self.timings = {}
        self.shift_route = None     # how the last construct_subspace applied the mean shift

    @property
    def M_device(self):
        if self._Md is None:
            self._Md = CsrMatrix(self.M_csr, self.device)
        return self._Md

    # (|mean| / rms fluctuation)^2 above which the implicit mean shift (round-off ~ eps * ratio^(1/2)) is replaced by the
    # explicit subtraction of the reference
    IMPLICIT_SHIFT_MAX_RATIO = 1.0e6

    def construct_subspace(self, u_data, u_rank, shifted=True, method='hep', verify=False,
                           oversampling=10, Omega=None, collective=None, return_device=False, faithful=False,
                           overwrite_data=False, pipelined_upload=True, implicit_shift=True, lazy_mean=True):
        """Same contract as PODProjector.py:699-852: returns (d, phi, Mphi, u_shift) as NumPy arrays of shape
        (r,), (n, r), (n, r), (n,).  ``u_data`` may be a NumPy array or a float64 CUDA tensor (rows = samples;
        with a collective, the local shard).  Extra keywords (randomized method only): ``oversampling``,
        ``Omega`` ((n, r+p) array or DeviceMultiVector fed to the solver), ``collective`` (sample-parallel),
        ``implicit_shift`` (apply the mean shift of :732-738 inside the products, (X - 1 u^T) B = X B - 1 (u^T B), instead
        of rewriting the stored snapshots; falls back to the explicit shift when the mean dominates the fluctuations),
        ``overwrite_data`` (allow an explicit shift to be applied in place to a device-resident ``u_data``)."""
        n_data, dim_u = u_data.shape
        collective = collective if collective is not None else NullCollective()
        n_total = n_data * collective.size()
        assert u_rank <= n_total, "number of samples needs to be greater than rank of projector"
        if method not in ('hep', 'ghep', 'inverse_ghep', 'randomized'):
            raise ValueError("Unavailable method")
        dev = self.device
        Md = self.M_device
        t0 = time.time()
        owns = not (K.is_device_tensor(u_data))
        pre = None          # pass-1 products formed while the snapshots were uploaded
        center = None       # vector still to be subtracted (implicitly) from every stored row
        if owns and method == 'randomized' and pipelined_upload:
            # host input: stream the snapshots to the device in row chunks and overlap the upload with the sample mean
            # and the whole range-finding pass  Y = X~^T X~ (M Omega)
            Omega = self._resolve_omega(Omega, dim_u, u_rank + oversampling)
            Xt, u_shift_d, center, pre = self._upload_pipelined(u_data, Md, Omega, shifted, collective)
            self.shift_route = 'pipelined' if shifted else 'none'
        else:
            Xt = _as_device_rows(u_data, dev)
            m_cols = u_rank + oversampling
            if shifted and method == 'randomized' and implicit_shift and lazy_mean and m_cols % 16 != 0:
                # the mean comes out of the range-finding pass itself (one extra column in the lift GEMM, see
                # linalg.SampleCovariance): no separate sweep over the stored snapshots
                center = "lazy"
                u_shift_d = None
                self.shift_route = 'implicit'
            elif shifted:
                # u_shift = mean over ALL samples (np.mean(u_data, axis=0), PODProjector.py:733), then X - shift
                u_shift_d = K.colsum(Xt, 1.0 / n_data)
                collective.allReduce(u_shift_d, 'avg')
                if method == 'randomized' and implicit_shift:
                    center = u_shift_d
                    self.shift_route = 'implicit'
                else:
                    Xt = self._shift_explicit(Xt, u_shift_d, owns or overwrite_data)
                    self.shift_route = 'explicit'
            else:
                u_shift_d = torch.zeros(dim_u, dtype=torch.float64, device=dev)
                self.shift_route = 'none'
        self.timings['upload_shift'] = time.time() - t0

        t1 = time.time()
        if method == 'randomized':
            d, phi_d, Mphi_d, ratio = self._randomized(Xt, Md, u_rank, oversampling, Omega, collective, faithful, center, pre)
            worst = None        # (|mean| / rms fluctuation)^2 summed over the ranks: ONE scalar exchange serves both checks
            if isinstance(center, str):
                # lazily centred solve: the mean is known now; its first apply lost ~eps * ratio digits, so redo the solve
                # with the known center when the mean is not small against the fluctuations
                center = u_shift_d = self._last_cov.center
                worst = float(collective.allReduce(ratio.reshape(1).clone(), 'sum'))
                if worst > SampleCovariance.LAZY_MAX_RATIO * collective.size():
                    self.shift_route = 'implicit (lazy mean redone)'
                    d, phi_d, Mphi_d, ratio = self._randomized(Xt, Md, u_rank, oversampling, Omega, collective, faithful, center, None)
                    worst = None
            if ratio is not None:
                if worst is None:
                    worst = float(collective.allReduce(ratio.reshape(1).clone(), 'sum'))
                if worst > self.IMPLICIT_SHIFT_MAX_RATIO * collective.size():
                    # the mean dominates the fluctuations: redo with the data shifted explicitly, as the reference does
                    Xt = self._shift_explicit(Xt, center, owns or overwrite_data)
                    center = None
                    self.shift_route = 'explicit-fallback'
                    d, phi_d, Mphi_d, _ = self._randomized(Xt, Md, u_rank, oversampling, Omega, collective, faithful, None, None)
        else:
            if collective.size() != 1:
                raise NotImplementedError("method='%s' is serial like the reference (PODProjector.py:683); "
                                          "use method='randomized' with a collective" % method)
            d, phi_d, Mphi_d = self._snapshot_eig(Xt, Md, u_rank, method)
        torch.cuda.synchronize(dev)
        self.timings['eigensolve'] = time.time() - t1

        if verify:
            if center is not None:
                Xt = self._shift_explicit(Xt, center, owns)
            r = u_rank - 1 if shifted else u_rank
            G = K.dgemm(K.HFB_TN, phi_d[:, :r], Mphi_d[:, :r]).cpu().numpy()
            print(f"Basis-Projector Orthogonality error: {np.linalg.norm(G - np.eye(r))}")
            coef = K.dgemm(K.HFB_NN, Xt, Mphi_d[:, :r])                       # (N, r)
            rec = K.dgemm(K.HFB_NT, coef, phi_d[:, :r])                       # (N, n)
            K.axpby_(1.0, Xt, -1.0, rec)
            err = torch.sqrt(K.coldot(Md.matmat_rows(rec).t().contiguous(), rec.t().contiguous())) \
                if n_data <= 512 else None
            if err is not None:
                nrm = torch.sqrt(K.coldot(Md.matmat_rows(Xt).t().contiguous(), Xt.t().contiguous()))
                rel = (err / nrm).cpu().numpy()
                print(f"Mean reconstruction error: {np.mean(rel):.3e}")
                print(f"Max reconstruction error: {np.max(rel):.3e}")
        if return_device:
            return d, phi_d, Mphi_d, u_shift_d
        return self._results_to_host(d, phi_d, Mphi_d, u_shift_d)

    @staticmethod
    def _shift_explicit(Xt, shift, in_place):
        """X - 1 shift^T written out (PODProjector.py:734); copies first unless the buffer may be overwritten."""
        if not in_place:
            Xc = K.padded_empty(Xt.shape[0], Xt.shape[1], Xt.device)
            Xc.copy_(Xt)
            Xt = Xc
        K.subtract_row_(Xt, shift)
        return Xt

    def _results_to_host(self, d, phi_d, Mphi_d, u_shift_d):
        """Device results -> NumPy through pinned staging buffers; the two (n x r) blocks travel back to back on the
        copy stream."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        cs = _copy_stream(dev)
        cs.wait_stream(main)
        hosts = []
        with torch.cuda.stream(cs):
            for t in (phi_d, Mphi_d, u_shift_d):
                h = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                t.record_stream(cs)
                hosts.append(h)
        cs.synchronize()
        return d, hosts[0].numpy(), hosts[1].numpy(), hosts[2].numpy()

    # ---------------------------------------------------------------- randomized GHEP (north star (a))
    def _resolve_omega(self, Omega, n, m):
        if Omega is None:
            return gaussian_omega(n, m, 1, self.device)
        if not isinstance(Omega, DeviceMultiVector):
            return DeviceMultiVector.from_dense(Omega, self.device)
        return Omega

    def _upload_pipelined(self, u_host, Md, Omega, shifted, collective, max_chunks=8):
        """Host snapshots -> device in row chunks on a copy stream.  As each chunk X_c lands the main stream (i) removes a
        PROVISIONAL mean p (the mean of the first chunk) from it in place, (ii) adds its column sums to the running mean,
        (iii) computes its rows of the projection W' = X' (M Omega) and (iv) accumulates its share of the lift
        Y0 += X_c'^T W_c' / N, so the PCIe transfer hides the whole range-finding pass.  With the true mean known at the
        end, delta = mean - p is small (O(sigma / sqrt(chunk))), and the exactly shifted products follow from rank-one
        corrections without cancellation:
            W~ = W' - 1 (delta^T B),      Y = Y0 - mean' (delta^T B)^T - delta (1^T W' / N - delta^T B)^T,
        where mean' = mean of the stored rows X'.  The stored rows keep the residual ``delta``; later products subtract it
        implicitly (SampleCovariance(center=delta)).  Returns (Xt', u_shift, delta or None, (W~, Y_local))."""
        dev = self.device
        src = u_host if isinstance(u_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(u_host, dtype=np.float64))
        N, n = src.shape
        m = Omega.nvec()
        Xt = K.padded_empty(N, n, dev)
        B = Md.matmat(Omega.tensor())                                        # M Omega, ready before the data arrives
        W = K.padded_empty(N, m, dev)
        Y = K.padded_empty(n, m, dev)
        bounds = _upload_chunks(N, max_chunks)
        main = torch.cuda.current_stream(dev)
        copy_stream = _copy_stream(dev)
        copy_stream.wait_stream(main)
        colsum = None
        prov = None

        def consume(i, lo, hi, ev):
            """Main-stream work on rows lo..hi once their copy (event ev) has landed."""
            nonlocal colsum, prov
            main.wait_event(ev)
            Xc = Xt[lo:hi]
            if shifted:
                if i == 0:
                    prov = K.colsum(Xc, 1.0 / (hi - lo))
                K.subtract_row_(Xc, prov)
                part = K.colsum(Xc, 1.0)
                colsum = part if colsum is None else colsum.add_(part)
            K.dgemm(K.HFB_NN, Xc, B, out=W[lo:hi])
            K.dgemm(K.HFB_TN, Xc, W[lo:hi], out=Y, alpha=1.0 / N, accumulate=(i > 0))

        if src.is_pinned():
            events = []
            with torch.cuda.stream(copy_stream):
                for lo, hi in bounds:
                    Xt[lo:hi].copy_(src[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    events.append(ev)
            for i, ((lo, hi), ev) in enumerate(zip(bounds, events)):
                consume(i, lo, hi, ev)
        else:
            # pageable input (a plain NumPy array, the documented drop-in call): a cudaMemcpy from pageable memory is staged
            # by the driver at ~10 GB/s.  Instead the rows go through a small ring of pinned buffers filled by torch's
            # multi-threaded host copy, so the host copy of piece j+1 overlaps the DMA of piece j and the GEMMs of the
            # previous chunk.
            ring = _staging_ring(dev, n)
            rows_per = ring[0][0].shape[0]
            # the staging copies are worth this rank's share of the host cores (torchrun exports OMP_NUM_THREADS=1, which
            # would leave torch's copy_ single-threaded; hfb_host_copy runs its own threads and streams around the cache)
            threads = _host_copy_threads()
            j = 0
            for i, (lo, hi) in enumerate(bounds):
                for s0 in range(lo, hi, rows_per):
                    s1 = min(hi, s0 + rows_per)
                    buf, free = ring[j % len(ring)]
                    free.synchronize()                                       # its previous DMA has finished
                    K.host_copy_(buf[:s1 - s0], src[s0:s1], threads)
                    with torch.cuda.stream(copy_stream):
                        Xt[s0:s1].copy_(buf[:s1 - s0], non_blocking=True)
                        free.record(copy_stream)
                    j += 1
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                consume(i, lo, hi, ev)
        Xt.record_stream(copy_stream)
        if not shifted:
            return Xt, torch.zeros(n, dtype=torch.float64, device=dev), None, (W, Y)
        mean_p = colsum.mul_(1.0 / N)                                        # mean of the stored rows X' (local)
        u_shift = prov + mean_p                                              # local mean of the original rows
        collective.allReduce(u_shift, 'avg')                                 # global mean (equal shard sizes)
        delta = u_shift - prov                                               # what the stored rows still carry
        sb = K.colsum(B, 1.0, weights=delta)                                 # delta^T B
        cw = K.colsum(W, 1.0 / N)                                            # 1^T W' / N
        K.subtract_row_(W, sb)
        K.rank1_update_(Y, -1.0, mean_p, sb)
        K.rank1_update_(Y, -1.0, delta, cw - sb)
        return Xt, u_shift, delta, (W, Y)

    def _randomized(self, Xt, Md, u_rank, oversampling, Omega, collective, faithful, center=None, pre=None):
        n = Xt.shape[1]
        m = u_rank + oversampling
        Omega = self._resolve_omega(Omega, n, m)
        assert Omega.nvec() >= u_rank
        cov = SampleCovariance(Xt, center=center)
        self._last_cov = cov
        C = SampleCovarianceOperator(cov, collective, 'avg')
        A = SandwichedCovarianceOperator(C, Md)
        self.info = {}
        Q0 = None
        if pre is not None:
            # range finder already formed during the upload: Q0 = avg_g (1/N_loc) X~_g^T W~_g  (= C M Omega)
            Q0 = DeviceMultiVector(pre[1])
            collective.allReduce(Q0, 'avg')
        d, U = doublePassG(A, Md, None, Omega, u_rank, s=1, faithful=faithful, info=self.info, Q0=Q0)
        Mphi = Md.matmat(U.tensor())
        return d, U.tensor(), Mphi, cov.center_ratio

    # ---------------------------------------------------------------- method of snapshots ('hep' :812-833)
    def _snapshot_eig(self, Xt, Md, u_rank, method):
        """'hep' is the reference's method of snapshots.  'ghep' (:743-773) and 'inverse_ghep' (:775-810) pose the
        SAME eigenproblem, (1/N) X X^T M phi = d phi with phi^T M phi = I, to ARPACK in generalized mode; their
        eigenpairs coincide with 'hep' (the reference's three outputs agree to ~1e-12 on the golden fixture).
        Here they are solved to machine precision by a dense factorisation instead of Lanczos: through the
        (N x N) snapshot Gram matrix when n_data <= dim_u, through the dense (n x n) pencil (H, M) otherwise."""
        n_data, n = Xt.shape
        if method != 'hep' and n_data > n:
            return self._dense_pencil_eig(Xt, Md, u_rank)
        Zt = Md.matmat_rows(Xt)                                   # (M X)^T, sample-major
        G = K.dgemm(K.HFB_NT, Xt, Zt, symmetric=True)             # X^T M X  (N x N): upper tiles + mirror
        del Zt
        Gs = 0.5 * (G + G.t())
        if n_data <= 1024:
            s, Uh = np.linalg.eigh(Gs.cpu().numpy())
            Uh = torch.as_tensor(Uh, device=self.device)
        else:
            s, Uh = torch.linalg.eigh(Gs.contiguous())
            s = s.cpu().numpy()
        d = s[::-1][:u_rank] / n_data
        Usel = K.to_padded(torch.flip(Uh, dims=[1])[:, :u_rank], self.device)
        phi = K.dgemm(K.HFB_TN, Xt, Usel)                         # X U  (n x r)
        nrm = weighted_l2_norm_vector(phi, Md)
        K.colscale_(phi, 1.0 / nrm)
        Mphi = Md.matmat(phi)
        return np.ascontiguousarray(d), phi, Mphi


    def _dense_pencil_eig(self, Xt, Md, u_rank):
        import scipy.linalg as sla
        n_data, n = Xt.shape
        if n > 4096:
            raise NotImplementedError("ghep / inverse_ghep with n_data > dim_u > 4096: use method='randomized'")
        Zt = Md.matmat_rows(Xt)                                   # (M X)^T
        H = K.dgemm(K.HFB_TN, Zt, Zt, alpha=1.0 / n_data).cpu().numpy()   # M X X^T M / N  (n x n)
        w, V = sla.eigh(0.5 * (H + H.T), self.M_csr.toarray())    # V^T M V = I
        d = np.ascontiguousarray(w[::-1][:u_rank])
        phi = K.to_padded(np.ascontiguousarray(V[:, ::-1][:, :u_rank]), self.device)
        return d, phi, Md.matmat(phi)


class StoredSnapshots:
    """Stand-in for the ``observable`` argument of PODProjector when the snapshots q_i = B u(m_i) are already
    stored: ``snapshots`` is the local (N_loc, n) array (host or device)."""

    def __init__(self, snapshots):
        self.snapshots = snapshots


class PODProjector:
    """Collective double-pass POD of stored snapshots (PODProjector.py:52-390, eigensolve part :359-384)."""

    def __init__(self, observable, prior=None, control_distribution=None, mesh_constructor_comm=None,
                 collective=None, parameters=None, device=None):
        self.observable = observable
        self.prior = prior
        self.control_distribution = control_distribution
        self.mesh_constructor_comm = mesh_constructor_comm
        self.collective = collective if collective is not None else NullCollective()
        self.parameters = parameters if parameters is not None else PODParameterList()
        self.device = device if device is not None else _default_device()
        self.d = None
        self.U_MV = None
        self.Omega = None

    def construct_subspace(self, Omega=None):
        t0 = time.time()
        snaps = self.observable.snapshots
        Xt = _as_device_rows(snaps, self.device)
        n = Xt.shape[1]
        # LocalPODOperator = LowRankOperator(ones/N_loc, LocalObservables); Global = CollectiveOperator(.., 'avg')
        A = SampleCovarianceOperator(SampleCovariance(Xt), self.collective, 'avg')
        m = self.parameters['rank'] + self.parameters['oversampling']
        if Omega is None:
            # rank 0 draws, everybody else zero, then bcast (PODProjector.py:367-374)
            if self.collective.rank() == 0:
                Omega = gaussian_omega(n, m, self.parameters['omega_seed'], self.device)
            else:
                Omega = DeviceMultiVector(n, m, device=self.device)
            self.collective.bcast(Omega, root=0)
        elif not isinstance(Omega, DeviceMultiVector):
            Omega = DeviceMultiVector.from_dense(Omega, self.device)
        self.Omega = Omega
        self.d, self.U_MV = doublePass(A, Omega, self.parameters['rank'], s=1)
        torch.cuda.synchronize(self.device)
        self._subspace_construction_time = time.time() - t0
        if self.parameters['verbose'] and self.collective.rank() == 0:
            print('Construction of POD subspace took ', self._subspace_construction_time, 's')
        if self.parameters['save_and_plot'] and self.collective.rank() == 0 and self.parameters['output_directory'] is not None:
            np.save(self.parameters['output_directory'] + 'POD_projector', mv_to_dense(self.U_MV))
            np.save(self.parameters['output_directory'] + 'POD_d', self.d)
        return self.d, self.U_MV
