"""Device-side building blocks of the sample-averaged randomized eigensolves.

Everything numerical here is a call into libhfb200.so (``_lib``); the only host arithmetic is the
O(m^3) work on (m x m) matrices (Cholesky / eigh of Gram matrices and of T), which the reference also
does on the host with NumPy (hIPPYlib ``doublePass``: ``np.linalg.eigh(T)``).

Layouts ("operand layouts are fixed by the reference"):
  * stored samples ``Xt``: (R, n) row-major, rows = samples -- ``u_data`` (N, n) of
    PODProjector.py:726, ``m_data``, or stored Jacobians (N, dQ, dM) viewed as (N*dQ, dM)
    (operatorWrappers.py:62-64);
  * sketches / bases: (n, m) row-major, column j = vector j (mv_utilities.py:31-49).
"""
import os

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack as _lapack
import torch

from . import _lib as K


class CsrMatrix:
    """Sparse SPD weight matrix (mass matrix M, prior precision R) resident on the device as int32 CSR,
    the format the reference exports (PODProjector.py:695-697)."""

    WIDE_DEFAULT = "frag"       # kernel 'auto' uses over the cluster plan (r02: 5 / 3 / 2 column groups per warp by width)
    # narrower blocks go to the generic L1-panel kernel (r02, m = 74: fragment kernel 0.129 ms = 2601 GB/s vs 0.222 ms)
    CLUSTER_MIN_COLS = int(__import__("os").environ.get("HFB_SPMM_MIN_COLS", 32))
    RUNS_MIN_COLS = 65          # 'auto': run-staged FMA kernel from this width (two column pairs per lane) up to 384 columns

    def __init__(self, M_csr, device, cluster_rows=True):
        M = M_csr.tocsr()
        if not M.has_canonical_format:      # the cluster kernels scatter entries by (row, column): no duplicates allowed
            M = M.copy()
            M.sum_duplicates()
        self.shape = M.shape
        self.nnz = M.nnz
        self.device = device
        self.rowptr = torch.as_tensor(np.asarray(M.indptr, dtype=np.int32), device=device)
        self.colind = torch.as_tensor(np.asarray(M.indices, dtype=np.int32), device=device)
        self.val = torch.as_tensor(np.asarray(M.data, dtype=np.float64), device=device)
        # SpMM plan (one-time host preprocessing, _build_plan): clusters of <= 16 mesh-neighbouring rows touching <= 32 (48
        # for dense rows) distinct columns; the cluster kernels stage exactly those rows of B in shared memory
        self.order = None
        self.plan = None
        import os
        # SpMM kernel for blocks of >= 32 columns over the cluster plan; HFB_SPMM_IMPL overrides for tuning runs and tests:
        #   "auto"  (default) "runs" for 65 <= m <= 384, else "frag"
        #   "runs"  run-staged FMA kernel: runs of consecutive B rows staged by one TMA copy each, CSR-order FMAs (m <= 384)
        #   "frag"  dense cluster block as host-packed DMMA A-fragment records, whole B rows staged by cp.async
        #   "ring"  the same records through resident CTAs with producer warps + a ring of cluster buffers (m <= 384)
        #   "dmma"  same arithmetic, records decoded in the kernel, double-buffered 64-column panels
        # (five further variants were measured and retired: profiles/r01_spmm_variants.md, tools/experiments/spmm_variants/)
        self.impl = os.environ.get("HFB_SPMM_IMPL", "auto")
        if cluster_rows and M.shape[0] == M.shape[1] and M.shape[0] >= 4096:
            try:
                self.plan = self._build_plan(M, device)
                self.order = self.plan["order"]      # the L1-panel kernel (narrow blocks) walks the same clusters
            except K.HfbError:
                self.plan = None          # e.g. a row with more entries than the column budget: the generic kernel handles it

    @staticmethod
    def _build_plan(M, device, max_rows=None, max_cols=None):
        import os
        # (16, 32) measured best on B200 for the cluster kernels (cfg2, m = 266; profiles/r01_spmm_variants.md: DMMA
        # fragment kernel 0.271 ms, (12, 32) 0.283, (8, 24) 0.286, (8, 20) 0.326): one cluster = two DMMA row halves x
        # eight k-steps, 71 KB of staged rows, 3 CTAs per SM
        max_rows = int(os.environ.get("HFB_SPMM_ROWS", 16)) if max_rows is None else max_rows
        if max_cols is None:
            # 7-point P1 stencils fill (16, 32) clusters; denser rows (9-point, P2: 9-20 entries) would shrink them to 2-3 rows
            # under 32 columns, so they get the DMMA kernels' largest column budget instead
            max_cols = int(os.environ.get("HFB_SPMM_COLS", 32 if M.nnz <= 8 * M.shape[0] else 48))
        order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, max_rows, max_cols)
        ncl = cptr.size - 1
        # caps of the plan (largest cluster: rows, runs of consecutive columns, distinct columns, entries), one O(nnz) host pass
        caps = K.csr_runs_measure(M.indptr, M.indices, order, cptr)
        assert caps["max_brow"] <= max_cols
        plan = {"nclusters": int(ncl), "order": torch.as_tensor(np.ascontiguousarray(order.astype(np.int32)), device=device),
                "max_rows": caps["max_rows"], "max_cols_cap": caps["max_brow"], "max_entries": caps["max_entries"],
                "max_runs": caps["max_runs"], "_host": (M.indptr, M.indices, M.data, order, cptr)}
        return plan

    @staticmethod
    def _panel_blobs(plan, device):
        """Per-cluster records of the panel DMMA kernel, packed on first use."""
        if "blobs" not in plan:
            indptr, indices, data, order, cptr = plan["_host"]
            plan["blobs"] = torch.as_tensor(K.csr_pack_clusters(indptr, indices, data, order, cptr, plan["max_rows"],
                                                                plan["max_cols_cap"], plan["max_entries"]), device=device)
        return plan

    _tma_blobs = _panel_blobs       # round-1 name

    @staticmethod
    def _frag_blobs(plan, device):
        """Fragment records of the whole-row DMMA kernel, packed on first use."""
        if "fblobs" not in plan:
            indptr, indices, data, order, cptr = plan["_host"]
            plan["fblobs"] = torch.as_tensor(K.csr_pack_clusters_frag(indptr, indices, data, order, cptr, plan["max_rows"],
                                                                      plan["max_cols_cap"]), device=device)
        return plan

    @staticmethod
    def _runs_blobs(plan, device):
        """Run records of the run-staged FMA kernel, packed on first use.  Returns the record plan, or None when the
        cluster plan does not fit the kernel (> 32 runs of consecutive columns in a cluster)."""
        if "rblobs" not in plan:
            indptr, indices, data, order, cptr = plan["_host"]
            try:
                blobs, caps = K.csr_pack_clusters_runs(indptr, indices, data, order, cptr)
                caps["blobs"] = torch.as_tensor(blobs, device=device)
                caps["nclusters"] = plan["nclusters"]
                plan["rblobs"] = caps
            except K.HfbError:
                plan["rblobs"] = None
        return plan["rblobs"]

    def matmat(self, B, out=None):
        """out (n, m) = M @ B for a dense row-major (n, m) block."""
        if K.TIMING is None:
            return self._matmat(B, out)[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # bench.py's roofline leg
        e0.record()
        out, kernel = self._matmat(B, out)
        e1.record()
        K.TIMING.append((("spmm", kernel, self.shape[0], B.shape[1]), e0, e1))
        return out

    def _matmat(self, B, out):
        m = B.shape[1]
        wide = K._ld(B) >= m + (m & 1)                      # the padding column of an odd width may be read
        if self.plan is not None and m >= self.CLUSTER_MIN_COLS and wide and B.data_ptr() % 16 == 0 and K._ld(B) % 2 == 0 and \
                self.plan["max_rows"] <= 16 and self.plan["max_cols_cap"] <= 48 and \
                (out is None or (out.data_ptr() % 16 == 0 and K._ld(out) % 2 == 0)):
            import os
            impl = self.impl
            if impl == "auto":
                # measured on B200 (profiles/r02_spmm_runs.md): run-staged FMA kernel for 65 <= m <= 384 (m = 266: 0.217 ms vs
                # 0.241 ms ring / 0.271 ms frag; m = 138: 0.145 ms vs 0.165 ms frag; m = 74: 0.122 ms vs 0.129 ms frag);
                # narrower blocks (one column pair per lane) stay with the per-cluster fragment kernel
                impl = os.environ.get("HFB_SPMM_WIDE", "runs" if self.RUNS_MIN_COLS <= m <= 384 else self.WIDE_DEFAULT)
            if impl == "runs":
                rplan = self._runs_blobs(self.plan, self.device) if m <= 384 else None
                if rplan is not None and K.csr_spmm_runs_slots(rplan, m, K._ld(B)) > 0:
                    return K.csr_spmm_runs(rplan, B, out), "csr_spmm_runs_kernel"
                impl = "ring" if 192 <= m <= 384 else "frag"          # wide pitch / many runs: the fragment kernels
            if impl == "ring" and 8 < self.plan["max_rows"] <= 16 and m <= 384:
                return K.csr_spmm_dmma_ring(self._frag_blobs(self.plan, self.device), B, out), "csr_spmm_ring_kernel"
            if impl in ("frag", "ring"):
                return K.csr_spmm_dmma_frag(self._frag_blobs(self.plan, self.device), B, out,
                                            int(os.environ.get("HFB_SPMM_FRAG_W", 0))), "csr_spmm_dmma_frag_kernel"
            if impl == "dmma":
                return K.csr_spmm_dmma(self._panel_blobs(self.plan, self.device), B, out), "csr_spmm_dmma_kernel"
            raise K.HfbError("unknown HFB_SPMM_IMPL '%s' (auto | runs | frag | ring | dmma)" % impl)
        return K.csr_spmm(self.rowptr, self.colind, self.val, B, out, order=self.order), "csr_spmm_panel_kernel"

    def matmat_rows(self, X, out=None):
        """out (N, n): row i = M @ X[i]  (= (M X^T)^T for sample-major X)."""
        return K.csr_spmm_rows(self.rowptr, self.colind, self.val, X, out)

    def spmm_bytes(self, m):
        """Algorithmic bytes of one SpMM with m columns."""
        n = self.shape[0]
        return self.nnz * 12 + (n + 1) * 4 + 2 * n * m * 8


class SampleCovariance:
    """Local shard of the sample-averaged operator  C = (1/N) sum_i X_i X_i^T  (optionally with a
    per-sample weight Gamma^-1 between the factors), held as the stacked row-major array ``Xt``.

    One object serves all three projectors:
      POD  : X_i = u_i                      (hp.LowRankOperator(ones/N, U), PODProjector.py:360)
      AS   : X_i = J_i^T, Gamma^-1 optional (MeanJTJfromDataOperator.mult, operatorWrappers.py:95-114)
      KLE  : X_i = m_i                      (sample covariance in place of prior.Rsolver, KLEProjector.py:103)
    ``apply`` is the two-GEMM form  Y = Xt^T (G (Xt B)) / N_loc  of those column-by-column loops.
    """

    def __init__(self, Xt, block=1, noise_cov_inv=None, center=None):
        """``center``: optional device vector c (n,).  The operator then acts on the rows x_i - c WITHOUT modifying
        ``Xt``: (X - 1 c^T) B = X B - 1 (c^T B) and (X - 1 c^T)^T W = X^T W - c (1^T W).  This stands in for the explicit
        ``u_data - u_shift`` of PODProjector.py:734 (saves a read+write sweep over the stored snapshots and leaves the
        caller's array untouched).  Round-off grows like eps * |c| / |x_i - c|; ``center_ratio`` reports
        (|c| / rms|x_i - c|)^2 (estimated through the first projection) so that callers can fall back to an explicit shift."""
        self.Xt = K._req(Xt, "Xt")
        self.rows, self.n = Xt.shape
        self.block = int(block)  # rows per sample (dQ for Jacobians, 1 for snapshots)
        assert self.rows % self.block == 0
        self.nsamples = self.rows // self.block
        self.noise_cov_inv = None
        if noise_cov_inv is not None:
            G = torch.as_tensor(np.asarray(noise_cov_inv, dtype=np.float64), device=Xt.device)
            assert tuple(G.shape) == (self.block, self.block)
            self.noise_cov_inv = G.contiguous()
        self._W = None
        self.center = None
        self.center_ratio = None        # device scalar, set by the first ``project`` when a center is present
        # center="lazy": the sample mean is NOT known yet.  It is obtained for free from the first lift, which multiplies
        # with [W | 1] instead of W (one more column inside the same GEMM tiles: X^T 1 = N * mean), and the first apply is
        # centred afterwards through  (X - 1 c^T)^T (X - 1 c^T) B = X^T (X B) - c (1^T X B)  -- this removes the separate
        # 8.6 GB column-sum sweep over the stored snapshots.  Round-off of that first apply grows like eps * |c|^2 / var
        # (``center_ratio``), so callers redo it with a known center when the ratio is large (LAZY_MAX_RATIO).
        self.lazy_center = False
        self._lazy_B = None
        if isinstance(center, str):
            assert center == "lazy" and self.noise_cov_inv is None and self.block == 1
            self.lazy_center = True
        elif center is not None:
            assert self.noise_cov_inv is None and self.block == 1, "implicit centering is for snapshot rows"
            self.center = center.contiguous()

    LAZY_MAX_RATIO = 1.0e3      # (|mean| / rms fluctuation)^2 above which the lazily centred first apply is redone

    def _wbuf(self, m):
        if self._W is None or self._W.shape[1] != m:
            self._Wext = K.padded_empty(self.rows, m + 1, self.Xt.device)     # one spare column: the ones column of a lazy lift
            self._Wext[:, m].fill_(1.0)
            self._W = self._Wext[:, :m]
            self._W2 = K.padded_empty(self.rows, m, self.Xt.device) if self.noise_cov_inv is not None else None
        return self._W

    def lazy_pending(self):
        """True while the center of a lazily centred operator has not been determined yet."""
        return self.lazy_center and self.center is None

    def project(self, B):
        """W (R, m) = Xt @ B  (and Gamma^-1 applied per sample block when present).  Returns (W, GW)
        where GW = blockdiag(Gamma^-1) W (or W itself)."""
        m = B.shape[1]
        W = self._wbuf(m)
        if B.data_ptr() % 16 or K._ld(B) % 2:
            B = K.to_padded(B, B.device)      # e.g. one column of a multivector block: stage into a TMA-aligned buffer
        K.dgemm(K.HFB_NN, self.Xt, B, out=W)
        if self.lazy_pending():
            self._lazy_B = B                      # centring of this apply happens after the lift (finish_lazy)
            return W, W
        if self.center is not None:
            sb = K.colsum(B, 1.0, weights=self.center)                                # c^T B  (m,), one sweep over B
            K.subtract_row_(W, sb)
            if self.center_ratio is None:
                self.center_ratio = self.rows * torch.dot(sb, sb) / torch.clamp_min(K.coldot(W, W).sum(), 1e-300)
        if self.noise_cov_inv is None:
            return W, W
        q = self.block
        W3 = W.as_strided((self.nsamples, q, m), (q * W.stride(0), W.stride(0), 1))
        GW3 = self._W2.as_strided((self.nsamples, q, m), (q * self._W2.stride(0), self._W2.stride(0), 1))
        K.dgemm_batched_small(K.HFB_NN, self.noise_cov_inv.unsqueeze(0), W3, GW3)
        return W, self._W2

    def apply(self, B, out=None, scale=None):
        """out (n, m) = scale * Xt^T G Xt B with scale = 1/nsamples by default (the local 'average'
        of SummedListOperator(average=True) / LowRankOperator(ones/N_loc))."""
        if self.lazy_pending():            # a direct apply cannot carry the extra column: determine the mean the ordinary way
            self.center = K.colsum(self.Xt, 1.0 / self.nsamples)
        _, GW = self.project(B)
        return self.lift(GW, out=out, scale=scale, weighted=True)

    def lift(self, W, out=None, scale=None, weighted=False, rows=None):
        """out = scale * Xt^T [G] W for an already computed projection W = Xt B (second half of ``apply``).
        ``weighted``: W already carries Gamma^-1.  ``rows = (lo, hi)``: only rows lo..hi of the result (a column block of
        Xt), used by the chunked lift that overlaps the sketch exchange."""
        GW = W
        if self.noise_cov_inv is not None and not weighted:
            q, m = self.block, W.shape[1]
            GWb = K.padded_empty(self.rows, m, self.Xt.device)
            K.dgemm_batched_small(K.HFB_NN, self.noise_cov_inv.unsqueeze(0),
                                  W.as_strided((self.nsamples, q, m), (q * W.stride(0), W.stride(0), 1)),
                                  GWb.as_strided((self.nsamples, q, m), (q * GWb.stride(0), GWb.stride(0), 1)))
            GW = GWb
        if scale is None:
            scale = 1.0 / self.nsamples
        lo, hi = (0, self.n) if rows is None else rows
        if self.lazy_pending():
            # raw block scale * X^T [W | 1]: the extra column (scale * X^T 1 = local mean) lands in the padding of ``out``
            m = GW.shape[1]
            assert out is not None and K._ld(out) >= m + 1 and GW.data_ptr() == self._Wext.data_ptr()
            ext = out.as_strided((out.shape[0], m + 1), (K._ld(out), 1))
            K.dgemm(K.HFB_TN, self.Xt if rows is None else self.Xt[:, lo:hi], self._Wext, out=ext, alpha=scale)
            return out
        out = K.dgemm(K.HFB_TN, self.Xt if rows is None else self.Xt[:, lo:hi], GW, out=out, alpha=scale)
        if self.center is not None:
            K.rank1_update_(out, -scale, self.center[lo:hi], K.colsum(GW, 1.0))       # - c (1^T W)
        return out

    def can_lazy(self, Y, m):
        """The lazy lift needs one spare column in the padding of the result block."""
        return self.lazy_pending() and K._ld(Y) >= m + 1

    def finish_lazy(self, Y, collective=None, mpi_op="avg"):
        """After the (all-reduced) raw lift of a lazily centred operator: read the mean from the extra column of ``Y``,
        apply the rank-one correction  Y -= c (1^T W / N)  and fix the center for all later products.  With a collective the
        extra column has been averaged together with the block ('avg' over equal shards = global mean); the (m,) vector
        1^T W / N is averaged here."""
        m = self._W.shape[1]
        ext = Y.as_strided((Y.shape[0], m + 1), (K._ld(Y), 1))
        c = ext[:, m].contiguous()                                           # global mean of the stored rows
        cw = K.colsum(self._W, 1.0 / self.nsamples)                          # 1^T W / N_loc
        if collective is not None:
            collective.allReduce(cw, mpi_op)
        K.rank1_update_(Y, -1.0, c, cw)
        self.center = c
        # (|c| / rms |x_i - c|)^2 through this projection: sum |W - 1 sb^T|^2 = sum |W|^2 - 2 sb.(1^T W) + rows |sb|^2
        sb = K.colsum(self._lazy_B, 1.0, weights=c)
        ww = K.coldot(self._W, self._W).sum() - 2.0 * self.nsamples * torch.dot(sb, cw) + self.rows * torch.dot(sb, sb)
        if collective is not None:
            pass                                                             # local estimate; callers sum it over the ranks
        self.center_ratio = self.rows * torch.dot(sb, sb) / torch.clamp_min(ww, 1e-300)
        self._lazy_B = None
        return c

    def gram_T(self, B, scale=None):
        """T_local (m, m) = scale * (Xt B)^T G (Xt B): the Rayleigh quotient B^T C B without forming C B
        (SURVEY.md 7 'T = Q^T A Q shortcut')."""
        assert not self.lazy_pending(), "the Rayleigh quotient needs the center (apply the operator first)"
        W, GW = self.project(B)
        if scale is None:
            scale = 1.0 / self.nsamples
        return K.dgemm(K.HFB_TN, W, GW, alpha=scale, symmetric=True)

    def flops_apply(self, m):
        return 4.0 * self.rows * self.n * m

    def flops_project(self, m):
        return 2.0 * self.rows * self.n * m


def _sym(G):
    return 0.5 * (G + G.T)


def sym_gram(Y, Z, alpha=1.0):
    """G (m, m) = alpha * Y^T Z for a product known to be symmetric (Z = B Y with B symmetric), on the DMMA kernel.
    The kernel's tiles have 128 rows; when m is a little above a multiple of 128 (m = 266: 10 columns over) the last
    row of tiles would be almost empty, so the leading (m0 x m0) block (m0 = multiple of 128) is computed with the
    symmetric flag (upper tiles + mirror) and the thin border G[:, m0:] by a second narrow GEMM, then mirrored:
    61 k instead of 87 k tile entries per k-step at m = 266."""
    m = Y.shape[1]
    r = m % 128
    if m <= 128 or r == 0 or r > 32 or Y.data_ptr() % 16 or Z.data_ptr() % 16 or K._ld(Y) % 2 or K._ld(Z) % 2:
        return K.dgemm(K.HFB_TN, Y, Z, alpha=alpha, symmetric=True)
    m0 = m - r
    G = K.padded_empty(m, m, Y.device)
    K.dgemm(K.HFB_TN, Y[:, :m0], Z[:, :m0], out=G[:m0, :m0], alpha=alpha, symmetric=True)
    K.dgemm(K.HFB_TN, Y, Z[:, m0:], out=G[:, m0:], alpha=alpha)
    G[m0:, :m0].copy_(G[:m0, m0:].t())
    return G


_side_streams = {}


def _side_stream(dev):
    st = _side_streams.get(dev.index)
    if st is None:
        st = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    return st


def _device_chol_enabled():
    import os
    return os.environ.get("HFB_DEVICE_CHOL", "1") != "0"


def b_orthonormalize_device(Y, Bmat=None, return_BQ=True):
    """Optimistic, host-free form of the two-pass Cholesky-QR for a well-conditioned sketch (the common case): every step
    is queued on the stream and NOTHING is read back --
        Z = B Y, G = Y^T Z, S1 = D^-1 chol(D^-1 G D^-1)^-1 (hfb_chol_inverse), Q1 = Y S1,
        Z1 = B Q1, G1 = Q1^T Z1, S2 = chol(G1)^-1 (hfb_chol_inverse, no scaling),
    and the clean-up factor S2 stays on the device for the caller to fold into the small matrices (T = S2^T T1 S2,
    U = Q1 (S2 V)).  ``info["pending"]`` carries the two device status vectors; the caller checks them when it fetches T
    (one synchronisation for everything) and falls back to the host-controlled ``b_orthonormalize`` on the untouched sketch
    ``Y`` when pass 1 needed a shift or was too ill-conditioned for a folded clean-up (cond * eps * m >= 1e-4)."""
    Z = Bmat.matmat(Y) if Bmat is not None else Y
    G = sym_gram(Y, Z)
    S1, stat1 = K.chol_inverse(G, scale_columns=True)
    Q1 = K.dgemm(K.HFB_NN, Y, S1, b_upper=True)                          # S1 is upper triangular: TRMM
    if Bmat is not None:
        Z1 = Bmat.matmat(Q1, out=Z)
    else:
        Z1 = Q1
    G1 = sym_gram(Q1, Z1)
    # the clean-up factor is not needed before T1 exists: its one-CTA kernel runs on a side stream, on one SM, while the
    # Rayleigh-quotient GEMM the caller queues next keeps the other 147 busy
    main = torch.cuda.current_stream(Y.device)
    side = _side_stream(Y.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        S2, stat2 = K.chol_inverse(G1, scale_columns=False)
    for t in (S2, stat2):
        t.record_stream(main)
    G1.record_stream(side)
    info = {"passes": 1, "shifted": 0, "cond": [], "route": "device",
            "pending": {"stat1": stat1, "stat2": stat2, "S2": S2, "sketch": Y, "side": side}}
    return Q1, (Z1 if return_BQ else None), info


def pending_ok(stats, m):
    """Host check of the device status vectors of ``b_orthonormalize_device`` (stats: (2, 8) NumPy array)."""
    eps = np.finfo(np.float64).eps
    s1, s2 = stats
    ok1 = s1[0] == 0.0 and s1[1] == 0.0 and np.isfinite(s1[2]) and s1[2] * eps * m < 1e-4
    ok2 = s2[0] == 0.0 and s2[1] == 0.0 and np.isfinite(s2[2]) and s2[2] < 4.0
    return bool(ok1 and ok2)


def b_orthonormalize(Y, Bmat=None, max_passes=5, return_BQ=True, defer_last=False, device_chol=None):
    """Orthonormalise the columns of the sketch Y (n, m) in the inner product of the sparse SPD matrix
    ``Bmat`` (None = Euclidean): the role of MultiVector.Borthogonalize / orthogonalize inside hIPPYlib's
    doublePassG / doublePass.

    hIPPYlib does column-by-column modified Gram-Schmidt (level-1 BLAS, one B-apply per column).  Here the
    same subspace is orthonormalised with GEMM-shaped passes (shifted Cholesky-QR, repeated): Z = B Y (SpMM),
    G = Y^T Z (DMMA, split-K), host Cholesky of the column-scaled (m x m) Gram matrix, Y <- Y R^-1 (DMMA).
    When the scaled Gram matrix is too ill-conditioned for a plain Cholesky factorisation a diagonal shift is
    added (the first pass then only pre-conditions the sketch; no direction is discarded) and passes repeat
    until the Gram matrix is the identity to round-off -- three passes for cond(Y) up to ~1e15, two for a
    well-conditioned sketch.  Columns that are exactly zero stay zero, as in hIPPYlib's MGS.  Eigenvalues d and
    span(U) of the eigensolve do not depend on which B-orthonormal basis of span(Y) is used.
    Returns (Q, BQ, info); Y's storage may be reused as scratch.

    ``defer_last=True`` (used by the double-pass drivers): when the first pass was well conditioned
    (cond(G) * eps * m < 1e-4, so Q1 is B-orthonormal to ~1e-7 or better) the clean-up pass is not applied to the
    (n x m) block.  Q1, Z1 = B Q1 and the device matrix G1 = Q1^T Z1 (``info["gram"]``) are returned instead; the
    caller folds the clean-up factor S2 = chol(G1)^-1 into the small matrices (T = S2^T T1 S2, U = Q1 (S2 V)),
    which is the same algebra as Q = Q1 S2 without the (n x m) update GEMM, one SpMM and one host round trip on
    the critical path."""
    n, m = Y.shape
    if device_chol is None:
        device_chol = _device_chol_enabled()
    if defer_last and device_chol and m <= K.CHOL_INVERSE_MAX and n >= m:
        return b_orthonormalize_device(Y, Bmat, return_BQ)
    spare = None
    eps = np.finfo(np.float64).eps
    info = {"passes": 0, "shifted": 0, "cond": [], "route": "host"}
    Z = None
    eye = np.eye(m)
    for it in range(max_passes):
        Z = Bmat.matmat(Y, out=Z) if Bmat is not None else Y
        G = _sym(sym_gram(Y, Z).cpu().numpy())
        d = np.sqrt(np.maximum(np.diag(G), 0.0))
        dead = d <= 0.0
        dinv = np.where(dead, 0.0, 1.0 / np.where(dead, 1.0, d))
        Gs = G * np.outer(dinv, dinv)
        Gs[dead, dead] = 1.0
        dev_from_I = np.abs(Gs - eye).max() if not dead.all() else 0.0
        if it > 0 and dev_from_I < 32 * eps and np.abs(d[~dead] - 1.0).max(initial=0.0) < 32 * eps:
            break  # B-orthonormal to round-off: nothing left to apply
        R = None
        shift = 0.0
        for attempt in range(8):
            Rt, fail = _lapack.dpotrf(Gs + shift * eye if shift else Gs, lower=0, clean=1, overwrite_a=0)
            if fail == 0:
                rd = np.abs(np.diag(Rt))
                cond = (rd.max() / rd.min()) ** 2
                if np.isfinite(cond) and (cond < 1e13 or shift > 0.0):
                    R = Rt
                    break
            shift = 100.0 * m * eps if shift == 0.0 else shift * 100.0
        if R is None:
            raise K.HfbError("b_orthonormalize: Gram matrix could not be factorised")
        if shift > 0.0:
            info["shifted"] += 1
        info["cond"].append(float(cond))
        S, fail = _lapack.dtrtri(R, lower=0)                     # R^-1 (upper triangular), then undo the column scaling
        if fail != 0:
            raise K.HfbError("b_orthonormalize: singular triangular factor")
        S *= dinv[:, None]
        spare = K.dgemm(K.HFB_NN, Y, K.to_padded(S, Y.device), out=spare, b_upper=True)    # S upper triangular: TRMM
        Y, spare = spare, Y                                      # ping-pong instead of copying the (n x m) block back
        info["passes"] += 1
        if defer_last and it == 0 and shift == 0.0 and cond * eps * m < 1e-4:
            Z = Bmat.matmat(Y, out=Z) if Bmat is not None else Y
            info["gram"] = sym_gram(Y, Z)                                          # device (m x m); fetched by the caller, asynchronously
            return Y, (Z if return_BQ else None), info
        if shift == 0.0 and it >= 1 and cond < 4.0:
            # the previous pass already left cond(G) ~ 1, so this pass is accurate to round-off
            Z = None
            break
    Q = Y
    BQ = None
    if return_BQ:
        BQ = Bmat.matmat(Q) if Bmat is not None else Q
    return Q, BQ, info


def cleanup_factor(G1):
    """S2 = chol(G1)^-1 (upper triangular) for the Gram matrix G1 ~ I of a nearly B-orthonormal basis:
    Q = Q1 S2 is B-orthonormal to round-off."""
    R, fail = _lapack.dpotrf(_sym(np.asarray(G1)), lower=0, clean=1, overwrite_a=0)
    if fail != 0:
        raise K.HfbError("cleanup_factor: Gram matrix of the pre-orthonormalised basis is not positive definite")
    S, fail = _lapack.dtrtri(R, lower=0)
    if fail != 0:
        raise K.HfbError("cleanup_factor: singular triangular factor")
    return S


_blas_ctl = None


def _eigh_threads():
    """(controller, threads) for the host eigensolve, or (None, 0) when the BLAS pool is already wide enough.  torchrun
    exports OMP_NUM_THREADS=1, which leaves LAPACK dsyevd single-threaded on the critical path of every rank (measured on the
    B200 host, m = 266: 4.25 ms with 1 thread, 3.84 ms with 4, 3.7 ms with 6-12; tools/lapack_threads_probe.py); the solve is
    worth this rank's share of the host cores for its few milliseconds.  Never more threads than cores: oversubscribed
    OpenBLAS spins for ~1 s per call."""
    global _blas_ctl
    if _blas_ctl is None:
        try:
            from threadpoolctl import ThreadpoolController
            ctl = ThreadpoolController()
            cur = max([i.get("num_threads", 1) for i in ctl.info() if i.get("user_api") == "blas"] or [0])
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            share = cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            want = int(os.environ.get("HFB_EIGH_THREADS", min(6, share)))
            _blas_ctl = (ctl, want) if (cur and want > cur) else (None, 0)
        except Exception:                                   # noqa: BLE001 -- threadpoolctl missing: keep the pool as it is
            _blas_ctl = (None, 0)
    return _blas_ctl


def top_k_eig(T, k):
    """eigh of the small symmetric matrix T on the host, top-k descending
    (hIPPYlib doublePass: np.linalg.eigh(T), sort descending, keep k)."""
    T = _sym(np.asarray(T))
    ctl, threads = _eigh_threads()
    if ctl is not None:
        with ctl.limit(limits=threads, user_api="blas"):
            d, V, fail = _lapack.dsyevd(T, compute_v=1, lower=1)
    else:
        d, V, fail = _lapack.dsyevd(T, compute_v=1, lower=1)
    if fail != 0:
        d, V = np.linalg.eigh(T)
    perm = np.argsort(d)[::-1][:k]
    return d[perm], np.ascontiguousarray(V[:, perm])


class CsrCGSolver:
    """Block conjugate gradients with a Jacobi preconditioner for a sparse SPD matrix on the device: the role
    of ``prior.Msolver`` / ``prior.Rsolver`` inside ``doublePassG`` (KLEProjector.py:163,
    activeSubspaceProjector.py:449) when the weight matrix is available as CSR.  All m right-hand sides advance
    together (SpMM + per-column dots/axpys); each column has its own step lengths.  Intended for mass-matrix-like
    (well conditioned) B; PDE-operator priors keep their upstream solvers."""

    def __init__(self, Bmat, rel_tol=1e-13, max_iter=1000, check_every=8, on_fail="raise"):
        """``on_fail``: what to do when ``max_iter`` is reached with a column still above ``rel_tol``:
        "raise" (HfbError; an inexact B^-1 silently degrades the eigenpairs of doublePassG) or "warn"
        (RuntimeWarning).  ``final_rel_res`` holds the largest relative residual of the last solve either way."""
        self.B = Bmat
        self.rel_tol, self.max_iter, self.check_every = rel_tol, max_iter, check_every
        assert on_fail in ("raise", "warn")
        self.on_fail = on_fail
        self.final_rel_res = None
        n = Bmat.shape[0]
        dev = Bmat.device
        # diagonal of B from the CSR arrays
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (Bmat.rowptr[1:] - Bmat.rowptr[:-1]).long())
        diag = torch.zeros(n, dtype=torch.float64, device=dev)
        mask = rows == Bmat.colind.long()
        diag[rows[mask]] = Bmat.val[mask]
        self.dinv = 1.0 / diag
        self.iterations = 0

    def solve_block(self, Y):
        n, m = Y.shape
        dev = Y.device
        X = K.padded_zeros(n, m, dev)
        R = K.padded_empty(n, m, dev)
        R.copy_(Y)
        Z = K.rowscale(self.dinv, R)
        P = K.padded_empty(n, m, dev)
        P.copy_(Z)
        AP = K.padded_empty(n, m, dev)
        rz = K.coldot(R, Z)
        r0 = torch.sqrt(K.coldot(R, R))
        r0 = torch.where(r0 > 0, r0, torch.ones_like(r0))
        for it in range(self.max_iter):
            self.B.matmat(P, out=AP)
            pAp = K.coldot(P, AP)
            alpha = torch.where(pAp > 0, rz / pAp, torch.zeros_like(rz))
            K.axpby_cols_(alpha, P, None, X)
            K.axpby_cols_(-alpha, AP, None, R)
            if (it + 1) % self.check_every == 0:
                rel = float((torch.sqrt(K.coldot(R, R)) / r0).max())
                if rel < self.rel_tol:
                    self.iterations = it + 1
                    self.final_rel_res = rel
                    break
            K.rowscale(self.dinv, R, out=Z)
            rz_new = K.coldot(R, Z)
            beta = torch.where(rz > 0, rz_new / rz, torch.zeros_like(rz))
            rz = rz_new
            K.axpby_cols_(None, Z, beta, P)
        else:
            self.iterations = self.max_iter
            self.final_rel_res = float((torch.sqrt(K.coldot(R, R)) / r0).max())
            if not self.final_rel_res < self.rel_tol:
                msg = ("CsrCGSolver: %d iterations reached with relative residual %.3e > rel_tol %.1e; the Jacobi-"
                       "preconditioned block CG is meant for mass-matrix-like (well conditioned) B -- pass a better "
                       "solver (any object with solve_block) for an ill-conditioned precision matrix"
                       % (self.max_iter, self.final_rel_res, self.rel_tol))
                if self.on_fail == "raise":
                    raise K.HfbError(msg)
                import warnings
                warnings.warn(msg, RuntimeWarning)
        return X

    def solve(self, x, b):
        """hippylib solver signature: solve(x, b) writes B^-1 b into x (vectors)."""
        x.storage_tensor().copy_(self.solve_block(b.storage_tensor()))

    def init_vector(self, x, dim):
        x.init(self.B.shape[0])
