"""GPU: each C-ABI kernel against a plain torch fp64 reference of the same op (bit-level agreement is not
expected for floating-point sums; tolerances are stated per test)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def K(cuda_device):
    from hippyflow_b200 import _lib
    _lib.lib()
    return _lib


SHAPES = [(100, 25, 289), (289, 25, 100), (513, 266, 1000), (130, 138, 77), (64, 74, 6400), (1000, 210, 333),
          (257, 4, 50), (300, 300, 300), (1, 1, 1), (16, 8, 16), (129, 137, 17)]


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_dgemm_matches_torch(K, cuda_device, layout, shape, splits):
    M, N, Kd = shape
    g = torch.Generator(device="cpu").manual_seed(M * 1000003 + N * 1009 + Kd)
    A = torch.randn(M, Kd, dtype=torch.float64, generator=g).to(cuda_device)
    B = torch.randn(Kd, N, dtype=torch.float64, generator=g).to(cuda_device)
    ref = 0.5 * (A @ B)
    Ad = K.to_padded(A.t().contiguous() if layout == K.HFB_TN else A, cuda_device)
    Bd = K.to_padded(B.t().contiguous() if layout == K.HFB_NT else B, cuda_device)
    C = K.dgemm(layout, Ad, Bd, alpha=0.5, splits=splits)
    assert rel(C, ref) < 1e-13          # fp64 GEMM, K <= 6400: a few ulp * sqrt(K)


def test_dgemm_split_k_is_bitwise_reproducible(K, cuda_device):
    A = K.padded_empty(256, 50000, cuda_device).normal_()
    B = K.padded_empty(50000, 74, cuda_device).normal_()
    C1 = K.dgemm(K.HFB_NN, A, B).clone()
    C2 = K.dgemm(K.HFB_NN, A, B).clone()
    assert torch.equal(C1, C2)
    assert K.lib().hfb_dgemm_auto_splits(0, 256, 74, 50000) > 1


def test_dgemm_linearity_at_config2_width(K, cuda_device):
    """Size-independent property at the benchmark width m = 266: (A)(B1 + B2) = A B1 + A B2 to round-off."""
    n, R, m = 40000, 512, 266
    A = K.padded_empty(R, n, cuda_device).normal_()
    B1 = K.padded_empty(n, m, cuda_device).normal_()
    B2 = K.padded_empty(n, m, cuda_device).normal_()
    Bs = K.padded_empty(n, m, cuda_device)
    Bs.copy_(B1 + B2)
    lhs = K.dgemm(K.HFB_NN, A, Bs).clone()
    rhs = K.dgemm(K.HFB_NN, A, B1).clone() + K.dgemm(K.HFB_NN, A, B2)
    assert rel(lhs, rhs) < 1e-13
    Y = K.dgemm(K.HFB_TN, A, lhs)                      # (n x m) = A^T W
    assert rel(Y, A.t() @ lhs) < 1e-13


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_dgemm_accumulate_flag(K, cuda_device, layout, splits):
    """HFB_GEMM_ACCUMULATE: C += alpha op(A) op(B), in the epilogue (one split) and in the split-K reduce (several);
    the chunked lift Y += X_c^T W_c of the pipelined upload relies on it."""
    for (M, N, Kd) in ((300, 266, 700), (129, 25, 4100), (64, 138, 33)):
        g = torch.Generator(device="cpu").manual_seed(M + N + Kd + splits)
        A = torch.randn(M, Kd, dtype=torch.float64, generator=g).to(cuda_device)
        B = torch.randn(Kd, N, dtype=torch.float64, generator=g).to(cuda_device)
        C0 = torch.randn(M, N, dtype=torch.float64, generator=g).to(cuda_device)
        Ad = K.to_padded(A.t().contiguous() if layout == K.HFB_TN else A, cuda_device)
        Bd = K.to_padded(B.t().contiguous() if layout == K.HFB_NT else B, cuda_device)
        C = K.to_padded(C0, cuda_device)
        K.dgemm(layout, Ad, Bd, out=C, alpha=-0.25, splits=splits, accumulate=True)
        ref = C0 - 0.25 * (A @ B)
        assert rel(C, ref) < 1e-13
    with pytest.raises(K.HfbError):
        K.dgemm(layout, Ad, Bd, alpha=1.0, accumulate=True)              # accumulate needs an existing out


def test_rank1_update_many_rows(K, cuda_device):
    """More than 65535 * 16 rows: the launch is split over several grids."""
    n, m = 1_100_003, 5
    Y = K.padded_zeros(n, m, cuda_device)
    x = torch.arange(n, dtype=torch.float64, device=cuda_device) * 1e-6
    y = torch.tensor([1.0, -2.0, 0.5, 3.0, 7.0], dtype=torch.float64, device=cuda_device)
    K.rank1_update_(Y, 2.0, x, y)
    assert torch.equal(Y, 2.0 * torch.outer(x, y))


def test_dgemm_rejects_bad_operands(K, cuda_device):
    A = torch.randn(8, 9, dtype=torch.float64, device=cuda_device)     # odd leading dimension
    B = K.padded_empty(9, 4, cuda_device).normal_()
    with pytest.raises(K.HfbError):
        K.dgemm(K.HFB_NN, A, B)
    with pytest.raises(K.HfbError):
        K.dgemm(K.HFB_NN, K.padded_empty(8, 10, cuda_device), B)         # inner dimensions differ


@pytest.mark.parametrize("m", [1, 25, 64, 138, 266, 300])
def test_csr_spmm(K, cuda_device, m):
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    M = syn.p1_mass_matrix(23, 17)
    n = M.shape[0]
    Md = CsrMatrix(M, cuda_device)
    B = np.random.default_rng(m).standard_normal((n, m))
    out = Md.matmat(K.to_padded(B, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(out, M @ B, rtol=1e-13, atol=1e-16)
    X = np.random.default_rng(m + 1).standard_normal((37, n))
    outr = Md.matmat_rows(K.to_padded(X, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(outr, (M @ X.T).T, rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("impl", ["dmma", "frag", "ring", "runs"])
@pytest.mark.parametrize("m", [33, 96, 137, 138, 266, 300, 383, 511, 1100])
def test_csr_spmm_clustered_kernels(K, cuda_device, impl, m):
    """The cluster-plan kernels (panel records / fragment records) on a mesh matrix large enough to get a
    plan (n >= 4096), odd and even widths, widths that need 1..4 column chunks; padding columns stay untouched."""
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    M = syn.p1_mass_matrix(70, 63)
    n = M.shape[0]
    Md = CsrMatrix(M, cuda_device)
    assert Md.plan is not None
    Md.impl = impl
    B = np.random.default_rng(m).standard_normal((n, m))
    out = K.padded_empty(n, m, cuda_device)
    full = out.as_strided((n, K._ld(out)), (K._ld(out), 1))
    full.fill_(7.0)
    Md.matmat(K.to_padded(B, cuda_device), out=out)
    np.testing.assert_allclose(out.cpu().numpy(), M @ B, rtol=1e-13, atol=1e-16)
    if K._ld(out) > m:
        assert bool((full[:, m:] == 7.0).all())
    again = Md.matmat(K.to_padded(B, cuda_device))
    assert torch.equal(again, out)                 # fixed summation order: bitwise reproducible


@pytest.mark.parametrize("mode", ["panel64", "panel128"])
@pytest.mark.parametrize("caps", [(8, 16), (8, 20), (8, 32), (8, 40), (12, 24), (16, 24), (16, 32), (16, 48), (5, 20)])
@pytest.mark.parametrize("m", [10, 74, 138, 266, 267, 600])
def test_csr_spmm_dmma_cluster_caps(K, cuda_device, caps, m, mode, monkeypatch):
    """Every instantiation of the cluster-dense DMMA kernel (1 or 2 row halves, 4..12 k-steps; double-buffered 64- and
    128-column panels) against SciPy: narrow blocks, partial
    last panels / chunks, odd widths; padding columns stay untouched; run-to-run bitwise equal."""
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    monkeypatch.setenv("HFB_SPMM_DMMA_NT", "2" if mode == "panel128" else "1")
    M = syn.p1_mass_matrix(41, 37).tocsr()
    n = M.shape[0]
    plan = CsrMatrix._tma_blobs(CsrMatrix._build_plan(M, cuda_device, max_rows=caps[0], max_cols=caps[1]), cuda_device)
    B = np.random.default_rng(m).standard_normal((n, m))
    out = K.padded_empty(n, m, cuda_device)
    full = out.as_strided((n, K._ld(out)), (K._ld(out), 1))
    full.fill_(7.0)
    K.csr_spmm_dmma(plan, K.to_padded(B, cuda_device), out)
    np.testing.assert_allclose(out.cpu().numpy(), M @ B, rtol=1e-13, atol=1e-16)
    if K._ld(out) > m:
        assert bool((full[:, m:] == 7.0).all())
    assert torch.equal(K.csr_spmm_dmma(plan, K.to_padded(B, cuda_device)), out)


@pytest.mark.parametrize("chunk", [0, 64, 144])
@pytest.mark.parametrize("caps", [(8, 16), (8, 24), (8, 32), (8, 40), (12, 24), (16, 24), (16, 32), (16, 48), (5, 20)])
@pytest.mark.parametrize("m", [10, 74, 138, 266, 267, 330, 600])
def test_csr_spmm_dmma_frag_cluster_caps(K, cuda_device, caps, m, chunk):
    """Fragment-record DMMA kernel: every (row halves, k-steps) instantiation, whole-row and chunked staging, widths that
    need 1..5 column groups per warp and several chunks, odd widths; padding untouched; bitwise reproducible."""
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    M = syn.p1_mass_matrix(41, 37).tocsr()
    n = M.shape[0]
    plan = CsrMatrix._frag_blobs(CsrMatrix._build_plan(M, cuda_device, max_rows=caps[0], max_cols=caps[1]), cuda_device)
    B = np.random.default_rng(m).standard_normal((n, m))
    out = K.padded_empty(n, m, cuda_device)
    full = out.as_strided((n, K._ld(out)), (K._ld(out), 1))
    full.fill_(7.0)
    K.csr_spmm_dmma_frag(plan, K.to_padded(B, cuda_device), out, chunk_cols=chunk)
    np.testing.assert_allclose(out.cpu().numpy(), M @ B, rtol=1e-13, atol=1e-16)
    if K._ld(out) > m:
        assert bool((full[:, m:] == 7.0).all())
    assert torch.equal(K.csr_spmm_dmma_frag(plan, K.to_padded(B, cuda_device), chunk_cols=chunk), out)     # bitwise reproducible


@pytest.mark.parametrize("caps", [(16, 24), (16, 32), (16, 48), (12, 24), (9, 20)])
@pytest.mark.parametrize("m", [10, 74, 138, 266, 267, 330, 384])
def test_csr_spmm_ring_vs_scipy_and_frag(K, cuda_device, caps, m):
    """Ring-pipelined kernel: every k-step instantiation, widths with 1..3 tiles per warp and a partial last tile, a mesh
    large enough that every resident CTA walks several clusters through its 2-8 slot ring; bitwise equal to the per-cluster
    fragment kernel (same summation order) and to itself run to run; padding columns stay untouched."""
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    M = syn.p1_mass_matrix(120, 100).tocsr()
    n = M.shape[0]
    plan = CsrMatrix._frag_blobs(CsrMatrix._build_plan(M, cuda_device, max_rows=caps[0], max_cols=caps[1]), cuda_device)
    assert plan["nclusters"] > 4 * 148
    B = np.random.default_rng(m).standard_normal((n, m))
    Bd = K.to_padded(B, cuda_device)
    out = K.padded_empty(n, m, cuda_device)
    full = out.as_strided((n, K._ld(out)), (K._ld(out), 1))
    full.fill_(7.0)
    try:
        K.csr_spmm_dmma_ring(plan, Bd, out)
    except K.HfbError as e:                      # two ring slots of (48 columns x 336) do not fit shared memory: refused, not wrong
        assert "unsupported" in str(e) and caps[1] > 32 and m > 300
        return
    np.testing.assert_allclose(out.cpu().numpy(), M @ B, rtol=1e-13, atol=1e-16)
    if K._ld(out) > m:
        assert bool((full[:, m:] == 7.0).all())
    assert torch.equal(K.csr_spmm_dmma_frag(plan, Bd), out)
    assert torch.equal(K.csr_spmm_dmma_ring(plan, Bd), out)


@pytest.mark.parametrize("caps", [(16, 32), (16, 24), (12, 24), (9, 20)])
@pytest.mark.parametrize("m", [10, 63, 74, 138, 139, 200, 266, 267, 330, 384])
def test_csr_spmm_runs_vs_scipy(K, cuda_device, caps, m, monkeypatch):
    """Run-staged FMA kernel: every column-pairs-per-lane instantiation, odd widths, a mesh large enough that every
    resident CTA walks several clusters through its 2-8 slot ring; bitwise reproducible and independent of the ring
    depth; padding columns stay untouched."""
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    M = syn.p1_mass_matrix(120, 100).tocsr()
    n = M.shape[0]
    plan = CsrMatrix._build_plan(M, cuda_device, max_rows=caps[0], max_cols=caps[1])
    rplan = CsrMatrix._runs_blobs(plan, cuda_device)
    assert rplan is not None and rplan["nclusters"] > 4 * 148
    B = np.random.default_rng(m).standard_normal((n, m))
    Bd = K.to_padded(B, cuda_device)
    out = K.padded_empty(n, m, cuda_device)
    full = out.as_strided((n, K._ld(out)), (K._ld(out), 1))
    full.fill_(7.0)
    assert K.csr_spmm_runs_slots(rplan, m, K._ld(Bd)) >= 2
    K.csr_spmm_runs(rplan, Bd, out)
    np.testing.assert_allclose(out.cpu().numpy(), M @ B, rtol=1e-13, atol=1e-16)
    if K._ld(out) > m:
        assert bool((full[:, m:] == 7.0).all())
    assert torch.equal(K.csr_spmm_runs(rplan, Bd), out)
    monkeypatch.setenv("HFB_RUNS_SLOTS", "2")
    assert torch.equal(K.csr_spmm_runs(rplan, Bd), out)          # same sums whatever the schedule


def test_csr_spmm_runs_irregular_matrix_views_and_fallback(K, cuda_device):
    """Ragged pattern with empty rows; a result block whose rows are only 16-byte aligned; a B block with a wide pitch
    (column view of a wider array) staged at that pitch; CsrMatrix falls back to the fragment kernel when the pitch is too
    wide for the ring or the plan has too many runs."""
    import scipy.sparse as sp
    from hippyflow_b200.linalg import CsrMatrix
    rng = np.random.default_rng(3)
    n = 6000
    A = sp.random(n, n, density=4.0 / n, random_state=7, format="csr")
    A = (A + A.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 17, n - 1):
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    Md = CsrMatrix(A, cuda_device, cluster_rows=False)
    Md.plan = CsrMatrix._build_plan(A, cuda_device, max_rows=16, max_cols=32)
    Md.order = Md.plan["order"]
    B = rng.standard_normal((n, 138))
    for impl in ("runs",):
        Md.impl = impl
        rplan = CsrMatrix._runs_blobs(Md.plan, cuda_device)
        assert rplan is not None
        Bd = K.to_padded(B, cuda_device)
        out = K.csr_spmm_runs(rplan, Bd).cpu().numpy()
        np.testing.assert_allclose(out, A @ B, rtol=1e-12, atol=1e-14)
        assert np.all(out[[0, 17, n - 1]] == 0.0)
        wide = torch.zeros((n, 146), dtype=torch.float64, device=cuda_device)      # ld = 146: even, not a multiple of 4
        view = wide[:, 2:140]
        K.csr_spmm_runs(rplan, Bd, view)
        np.testing.assert_allclose(view.cpu().numpy(), A @ B, rtol=1e-12, atol=1e-14)
        assert bool((wide[:, :2] == 0).all()) and bool((wide[:, 140:] == 0).all())
        # B as a column view of a wider block: pitch 160 is staged as is; the LAST rows' copies stop at the view's width
        Bw = torch.full((n, 160), float("nan"), dtype=torch.float64, device=cuda_device)
        Bw[:, 16:154] = torch.as_tensor(B, device=cuda_device)
        np.testing.assert_allclose(K.csr_spmm_runs(rplan, Bw[:, 16:154]).cpu().numpy(), A @ B, rtol=1e-12, atol=1e-14)
        # pitch 4096: no two slots fit -> CsrMatrix uses the fragment kernel instead
        Bx = torch.zeros((n, 4096), dtype=torch.float64, device=cuda_device)
        Bx[:, :138] = torch.as_tensor(B, device=cuda_device)
        assert K.csr_spmm_runs_slots(rplan, 138, 4096) == 0
        np.testing.assert_allclose(Md.matmat(Bx[:, :138]).cpu().numpy(), A @ B, rtol=1e-12, atol=1e-14)
    # more than 32 runs in a cluster: no run records, 'runs' falls back
    Md.plan = CsrMatrix._build_plan(A, cuda_device, max_rows=16, max_cols=48)
    Md.impl = "runs"
    np.testing.assert_allclose(Md.matmat(K.to_padded(B, cuda_device)).cpu().numpy(), A @ B, rtol=1e-12, atol=1e-14)


def test_csr_spmm_dmma_frag_irregular_and_unaligned_rows(K, cuda_device):
    """Ragged pattern with empty rows; result block whose rows are only 16-byte aligned (128-bit store path)."""
    import scipy.sparse as sp
    from hippyflow_b200.linalg import CsrMatrix
    rng = np.random.default_rng(3)
    n = 6000
    A = sp.random(n, n, density=4.0 / n, random_state=7, format="csr")
    A = (A + A.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 17, n - 1):
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    plan = CsrMatrix._frag_blobs(CsrMatrix._build_plan(A, cuda_device, max_rows=16, max_cols=48), cuda_device)
    B = rng.standard_normal((n, 138))
    Bd = K.to_padded(B, cuda_device)
    out = K.csr_spmm_dmma_frag(plan, Bd).cpu().numpy()
    np.testing.assert_allclose(out, A @ B, rtol=1e-12, atol=1e-14)
    assert np.all(out[[0, 17, n - 1]] == 0.0)
    wide = torch.zeros((n, 146), dtype=torch.float64, device=cuda_device)      # ld = 146: even, not a multiple of 4
    view = wide[:, 2:140]                                                     # base 16-byte aligned, rows not 32-byte aligned
    K.csr_spmm_dmma_frag(plan, Bd, view)
    np.testing.assert_allclose(view.cpu().numpy(), A @ B, rtol=1e-12, atol=1e-14)
    assert bool((wide[:, :2] == 0).all()) and bool((wide[:, 140:] == 0).all())


def test_csr_spmm_dmma_irregular_matrix_and_limits(K, cuda_device):
    """Ragged non-mesh pattern with empty rows through the DMMA kernel; caps beyond its register budget are refused."""
    import scipy.sparse as sp
    from hippyflow_b200.linalg import CsrMatrix
    rng = np.random.default_rng(3)
    n = 6000
    A = sp.random(n, n, density=4.0 / n, random_state=7, format="csr")
    A = (A + A.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 17, n - 1):
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    plan = CsrMatrix._tma_blobs(CsrMatrix._build_plan(A, cuda_device, max_rows=16, max_cols=48), cuda_device)
    B = rng.standard_normal((n, 138))
    out = K.csr_spmm_dmma(plan, K.to_padded(B, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(out, A @ B, rtol=1e-12, atol=1e-14)
    assert np.all(out[[0, 17, n - 1]] == 0.0)
    big = CsrMatrix._tma_blobs(CsrMatrix._build_plan(A, cuda_device, max_rows=32, max_cols=64), cuda_device)
    with pytest.raises(K.HfbError):
        K.csr_spmm_dmma(big, K.to_padded(B, cuda_device))


@pytest.mark.parametrize("impl", ["dmma", "frag", "ring", "runs", "auto"])
def test_csr_spmm_clustered_irregular_matrix(K, cuda_device, impl):
    """Non-mesh sparsity: random symmetric pattern with ragged rows (1..20 entries) plus a few empty rows."""
    import scipy.sparse as sp
    from hippyflow_b200.linalg import CsrMatrix
    rng = np.random.default_rng(3)
    n = 6000
    A = sp.random(n, n, density=4.0 / n, random_state=7, format="csr")
    A = (A + A.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 17, n - 1):
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    Md = CsrMatrix(A, cuda_device, cluster_rows=False)
    Md.plan = CsrMatrix._build_plan(A, cuda_device, max_rows=16, max_cols=64)    # rows have up to ~20 entries
    Md.order = Md.plan["order"]
    Md.impl = impl
    B = rng.standard_normal((n, 138))
    out = Md.matmat(K.to_padded(B, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(out, A @ B, rtol=1e-12, atol=1e-14)
    assert np.all(out[[0, 17, n - 1]] == 0.0)


def test_column_kernels(K, cuda_device):
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((1000, 37)), rng.standard_normal((1000, 37))
    Xd, Yd = K.to_padded(X, cuda_device), K.to_padded(Y, cuda_device)
    np.testing.assert_allclose(K.coldot(Xd, Yd).cpu().numpy(), np.einsum("ij,ij->j", X, Y), rtol=1e-13)
    np.testing.assert_allclose(K.colsum(Xd, 0.25).cpu().numpy(), 0.25 * X.sum(0), rtol=1e-12, atol=1e-14)
    w = rng.standard_normal(1000)
    np.testing.assert_allclose(K.colsum(Xd, 0.5, weights=torch.as_tensor(w, device=cuda_device)).cpu().numpy(),
                               0.5 * (w @ X), rtol=1e-12, atol=1e-13)
    s = rng.standard_normal(37)
    sd = torch.as_tensor(s, device=cuda_device)
    Z = Xd.clone()
    K.rank1_update_(Z, -0.75, torch.as_tensor(w, device=cuda_device), sd)
    np.testing.assert_allclose(Z.cpu().numpy(), X - 0.75 * np.outer(w, s), rtol=1e-14, atol=1e-15)
    Z = Xd.clone()
    K.colscale_(Z, sd)
    np.testing.assert_allclose(Z.cpu().numpy(), X * s, rtol=1e-15)
    Z = Xd.clone()
    K.subtract_row_(Z, sd)
    np.testing.assert_allclose(Z.cpu().numpy(), X - s, rtol=1e-15)
    Z = Yd.clone()
    K.axpby_(2.0, Xd, -0.5, Z)
    np.testing.assert_allclose(Z.cpu().numpy(), 2 * X - 0.5 * Y, rtol=1e-15, atol=1e-15)
    Z = Yd.clone()
    K.axpby_cols_(sd, Xd, None, Z)
    np.testing.assert_allclose(Z.cpu().numpy(), X * s + Y, rtol=1e-14, atol=1e-15)
    r = rng.standard_normal(1000)
    out = K.rowscale(torch.as_tensor(r, device=cuda_device), Xd)
    np.testing.assert_allclose(out.cpu().numpy(), X * r[:, None], rtol=1e-15)


def test_batched_small_gemm(K, cuda_device):
    rng = np.random.default_rng(1)
    G = rng.standard_normal((10, 10))
    W = rng.standard_normal((7, 10, 33))
    out = torch.empty((7, 10, 33), dtype=torch.float64, device=cuda_device)
    K.dgemm_batched_small(K.HFB_NN, torch.as_tensor(G, device=cuda_device).unsqueeze(0),
                          torch.as_tensor(W, device=cuda_device), out)
    np.testing.assert_allclose(out.cpu().numpy(), np.einsum("ab,ibc->iac", G, W), rtol=1e-13, atol=1e-14)
    Phi = rng.standard_normal((10, 6))
    out2 = torch.empty((7, 6, 33), dtype=torch.float64, device=cuda_device)
    K.dgemm_batched_small(K.HFB_TN, torch.as_tensor(Phi, device=cuda_device).unsqueeze(0),
                          torch.as_tensor(W, device=cuda_device), out2)
    np.testing.assert_allclose(out2.cpu().numpy(), np.einsum("qa,iqc->iac", Phi, W), rtol=1e-13, atol=1e-14)


def test_fill_random_is_sharding_independent(K, cuda_device):
    full = K.padded_empty(64, 101, cuda_device)
    K.fill_random_(full, seed=11)
    part = K.padded_empty(16, 101, cuda_device)
    K.fill_random_(part, seed=11, row_offset=32)
    assert torch.equal(part, full[32:48])
    big = K.padded_empty(4096, 256, cuda_device)
    K.fill_random_(big, seed=3)
    assert abs(float(big.mean())) < 5e-3 and abs(float(big.std()) - 1.0) < 5e-3


def test_block_cg(K, cuda_device):
    from hippyflow_b200 import synthetic as syn
    from hippyflow_b200.linalg import CsrCGSolver, CsrMatrix
    M = syn.p1_mass_matrix(20)
    Md = CsrMatrix(M, cuda_device)
    Y = np.random.default_rng(5).standard_normal((M.shape[0], 9))
    X = CsrCGSolver(Md).solve_block(K.to_padded(Y, cuda_device)).cpu().numpy()
    assert np.linalg.norm(M @ X - Y) / np.linalg.norm(Y) < 1e-12


@pytest.mark.parametrize("layout,n,kd", [(2, 300, 777), (1, 266, 5000), (2, 129, 64), (1, 138, 20000), (0, 256, 256)])
def test_dgemm_symmetric_result_flag(K, cuda_device, layout, n, kd):
    """HFB_GEMM_SYMMETRIC: only tiles on/above the diagonal are computed, the rest is mirrored."""
    g = torch.Generator(device="cpu").manual_seed(n + kd)
    X = torch.randn(n, kd, dtype=torch.float64, generator=g).to(cuda_device)
    if layout == 2:        # X X^T
        A, B, ref = K.to_padded(X, cuda_device), K.to_padded(X, cuda_device), X @ X.t()
    elif layout == 1:      # (X^T)^T X^T = X X^T with A = B = X^T stored (kd, n)
        Xt = K.to_padded(X.t().contiguous(), cuda_device)
        A, B, ref = Xt, Xt, X @ X.t()
    else:                  # NN with a symmetric product: S S (S symmetric)
        S = X[:, :n] + X[:, :n].t()
        A, B, ref = K.to_padded(S, cuda_device), K.to_padded(S, cuda_device), S @ S
    for splits in (0, 1, 4):
        C = K.dgemm(layout, A, B, symmetric=True, splits=splits)
        assert float((C - ref).norm() / ref.norm()) < 1e-13
        assert torch.equal(C, C.t())


def test_dgemm_random_shape_fuzz(K, cuda_device):
    """Random ragged shapes / leading dimensions / split counts across all layouts (fixed seed)."""
    rng = np.random.default_rng(1234)
    for trial in range(60):
        layout = int(rng.integers(0, 3))
        M = int(rng.integers(1, 700))
        N = int(rng.choice([1, 2, 7, 15, 25, 64, 74, 137, 138, 210, 266, 300, 511]))
        Kd = int(rng.integers(1, 3000))
        splits = int(rng.choice([0, 0, 1, 2, 5]))
        A = torch.as_tensor(rng.standard_normal((M, Kd)), device=cuda_device)
        B = torch.as_tensor(rng.standard_normal((Kd, N)), device=cuda_device)
        ref = A @ B
        pad_a, pad_b = int(rng.choice([2, 16])), int(rng.choice([2, 16]))
        Ad = K.to_padded(A.t().contiguous() if layout == K.HFB_TN else A, cuda_device, pad=pad_a)
        Bd = K.to_padded(B.t().contiguous() if layout == K.HFB_NT else B, cuda_device, pad=pad_b)
        out = K.padded_empty(M, N, cuda_device, pad=int(rng.choice([1, 2, 16])))
        out.fill_(float("nan"))
        K.dgemm(layout, Ad, Bd, out=out, splits=splits)
        err = float((out - ref).norm() / ref.norm())
        assert err < 1e-13, (trial, layout, M, N, Kd, splits, err)


def test_peer_exchange_emulated_ranks_on_one_gpu():
    """The fused lift + reduce-scatter epilogue (hfb_dgemm_peer), the flag barrier, the fixed-order slot reduction and the
    gather of hippyflow_b200/peer.py with 2-4 emulated ranks on ONE device (every 'peer' pointer local): bitwise equal to the
    rank-ordered sum of plain lifts, with and without pipeline chunks.  Runs in a subprocess because a barrier time-out
    traps the kernel (and the context with it)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "peer_selftest.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_SELFTEST_OK" in r.stdout, r.stdout[-3000:]
