"""GPU: the strided-batch DMMA GEMM, the device Cholesky-QR factor, the batched Jacobi SVD and the per-sample operations
built on them (batched randomized SVD of stored Jacobians, output active subspace in operator form, prior.Hlr branch).
Kernel checks compare with a plain torch / NumPy fp64 evaluation of the same operation; algorithm checks with the oracle."""
import numpy as np
import pytest
import torch

from hippyflow_b200 import synthetic as syn
from oracle import hippylib_np as hnp
from oracle import projectors_np as P
from conftest import subspace_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda_device):
    import hippyflow_b200 as hf
    from hippyflow_b200 import _lib as K
    K.lib()
    return hf, K, cuda_device


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def stack(K, dev, arr):
    """NumPy (batch, rows, cols) -> TMA-conforming device stack."""
    t = K.batched_empty(arr.shape[0], arr.shape[1], arr.shape[2], dev)
    t.copy_(torch.as_tensor(arr))
    return t


# ------------------------------------------------------------------ strided-batch GEMM
@pytest.mark.parametrize("layout", ["NN", "TN", "NT"])
@pytest.mark.parametrize("shape", [(5, 100, 70, 33), (3, 200, 128, 200), (9, 37, 266, 100), (2, 130, 18, 1000), (1, 64, 64, 64)])
@pytest.mark.parametrize("shared", ["none", "A", "B"])
def test_dgemm_batched_independent_vs_torch(env, layout, shape, shared):
    hf, K, dev = env
    batch, M, N, Kd = shape
    rng = np.random.default_rng(hash((layout, shape, shared)) % 2**32)
    a_shape = (Kd, M) if layout == "TN" else (M, Kd)
    b_shape = (N, Kd) if layout == "NT" else (Kd, N)
    A = rng.standard_normal((1 if shared == "A" else batch,) + a_shape)
    B = rng.standard_normal((1 if shared == "B" else batch,) + b_shape)
    Ad = K.to_padded(A[0], dev) if shared == "A" else stack(K, dev, A)
    Bd = K.to_padded(B[0], dev) if shared == "B" else stack(K, dev, B)
    opA = np.swapaxes(A, 1, 2) if layout == "TN" else A
    opB = np.swapaxes(B, 1, 2) if layout == "NT" else B
    ref = 0.5 * np.matmul(opA, opB)
    lay = {"NN": K.HFB_NN, "TN": K.HFB_TN, "NT": K.HFB_NT}[layout]
    out = K.dgemm_batched(lay, Ad, Bd, alpha=0.5)
    assert tuple(out.shape) == (batch, M, N)
    assert rel(out.cpu().numpy(), ref) < 1e-13
    # accumulate into the same outputs
    K.dgemm_batched(lay, Ad, Bd, out=out, alpha=0.5, accumulate=True)
    assert rel(out.cpu().numpy(), 2 * ref) < 1e-13


@pytest.mark.parametrize("layout", ["NN", "TN", "NT"])
@pytest.mark.parametrize("shape", [(7, 100, 100, 1000), (64, 30, 45, 121), (3, 200, 17, 4100), (300, 16, 16, 50)])
def test_dgemm_batched_reduce_vs_torch(env, layout, shape):
    """C = sum_b op(A_b) op(B_b): the K loop runs over (sample, k); a K tail (K % 16 != 0) must read zeros, never the next
    sample's rows."""
    hf, K, dev = env
    batch, M, N, Kd = shape
    rng = np.random.default_rng(hash((layout, shape)) % 2**32)
    a_shape = (Kd, M) if layout == "TN" else (M, Kd)
    b_shape = (N, Kd) if layout == "NT" else (Kd, N)
    A, B = rng.standard_normal((batch,) + a_shape), rng.standard_normal((batch,) + b_shape)
    opA = np.swapaxes(A, 1, 2) if layout == "TN" else A
    opB = np.swapaxes(B, 1, 2) if layout == "NT" else B
    ref = np.matmul(opA, opB).sum(0) / batch
    lay = {"NN": K.HFB_NN, "TN": K.HFB_TN, "NT": K.HFB_NT}[layout]
    Ad, Bd = stack(K, dev, A), stack(K, dev, B)
    out = K.dgemm_batched(lay, Ad, Bd, alpha=1.0 / batch, reduce=True)
    assert rel(out.cpu().numpy(), ref) < 1e-13
    out2 = K.dgemm_batched(lay, Ad, Bd, alpha=1.0 / batch, reduce=True)
    assert torch.equal(out, out2)                                            # deterministic split-K over the folded range
    K.dgemm_batched(lay, Ad, Bd, out=out, alpha=1.0 / batch, reduce=True, accumulate=True)
    assert rel(out.cpu().numpy(), 2 * ref) < 1e-13


def test_dgemm_batched_on_jacobian_view_and_argument_checks(env):
    hf, K, dev = env
    rng = np.random.default_rng(5)
    J = rng.standard_normal((6, 100, 333))                                   # dM odd: the stacked copy gets a padded ld
    J2, J3 = hf.stacked_jacobians(J, dev)
    assert tuple(J3.shape) == (6, 100, 333) and J3.stride(0) == 100 * J3.stride(1)
    E = rng.standard_normal((100, 20))
    out = K.dgemm_batched(K.HFB_TN, J3, K.to_padded(E, dev))
    assert rel(out.cpu().numpy(), np.einsum("iqm,qr->imr", J, E)) < 1e-13
    with pytest.raises(K.HfbError):
        K.dgemm_batched(K.HFB_NN, J3, K.to_padded(E, dev))                   # inner dimensions differ
    with pytest.raises(K.HfbError):
        K.dgemm_batched(K.HFB_TN, J3, stack(K, dev, rng.standard_normal((5, 100, 20))))   # batch sizes differ


@pytest.mark.parametrize("shape", [(1000, 266), (300, 138), (64, 17), (5000, 522)])
@pytest.mark.parametrize("layout", ["NN", "TN"])
def test_dgemm_upper_triangular_B_skips_zero_blocks(env, shape, layout):
    """HFB_GEMM_B_UPPER (the TRMM Q = Y S of Cholesky-QR): same result as the full product when B is upper triangular;
    garbage below the diagonal past the tile boundary is never read."""
    hf, K, dev = env
    M, N = shape
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((N, M) if layout == "TN" else (M, N))
    S = np.triu(rng.standard_normal((N, N)))
    lay = K.HFB_TN if layout == "TN" else K.HFB_NN
    ref = (A.T if layout == "TN" else A) @ S
    out = K.dgemm(lay, K.to_padded(A, dev), K.to_padded(S, dev), b_upper=True)
    assert rel(out.cpu().numpy(), ref) < 1e-13
    full = K.dgemm(lay, K.to_padded(A, dev), K.to_padded(S, dev))
    assert rel(out.cpu().numpy(), full.cpu().numpy()) < 1e-14
    with pytest.raises(K.HfbError):
        K.dgemm(lay, K.to_padded(A, dev), K.to_padded(S[:, : N - 1], dev), b_upper=True)


# ------------------------------------------------------------------ device Cholesky-QR factor
@pytest.mark.parametrize("m", [1, 7, 8, 9, 64, 138, 266, 513, 1024])
def test_chol_inverse_vs_numpy(env, m):
    hf, K, dev = env
    rng = np.random.default_rng(m)
    Y = rng.standard_normal((m + 50, m)) * (10.0 ** rng.uniform(-3, 3, size=m))[None, :]   # badly scaled columns
    G = Y.T @ Y
    S, stat = K.chol_inverse(K.to_padded(G, dev), scale_columns=True)
    S, stat = S.cpu().numpy(), stat.cpu().numpy()
    assert stat[0] == 0 and stat[1] == 0 and stat[7] == m and stat[3] == 1
    assert np.array_equal(S, np.triu(S))
    assert np.abs(S.T @ G @ S - np.eye(m)).max() < 1e-9 * max(1.0, stat[2])
    d = np.sqrt(np.diag(G))
    R = np.linalg.cholesky(G / np.outer(d, d)).T
    Sref = np.linalg.inv(R) / d[:, None]
    assert rel(S, Sref) < 1e-10 * max(1.0, stat[2] ** 0.5)
    rd = np.diag(R)
    np.testing.assert_allclose(stat[2], (rd.max() / rd.min()) ** 2, rtol=1e-8)
    np.testing.assert_allclose(stat[4], np.abs(G / np.outer(d, d) - np.eye(m)).max(), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(stat[5], np.abs(d - 1).max(), rtol=1e-12)
    # without column scaling (clean-up factor of a nearly orthonormal basis)
    Q = np.linalg.qr(rng.standard_normal((m + 50, m)))[0] @ (np.eye(m) + 1e-6 * rng.standard_normal((m, m)))
    G1 = Q.T @ Q
    S2, st2 = K.chol_inverse(K.to_padded(G1, dev), scale_columns=False)
    S2, st2 = S2.cpu().numpy(), st2.cpu().numpy()
    assert st2[0] == 0 and st2[1] == 0 and st2[2] < 1.01
    assert rel(S2, np.linalg.inv(np.linalg.cholesky(G1).T)) < 1e-13


def test_chol_inverse_shift_dead_column_and_failure_flags(env):
    hf, K, dev = env
    rng = np.random.default_rng(0)
    m = 40
    Y = rng.standard_normal((200, 5)) @ rng.standard_normal((5, m))          # rank 5: Cholesky needs the shift
    G = Y.T @ Y
    S, stat = K.chol_inverse(K.to_padded(G, dev))
    stat = stat.cpu().numpy()
    assert stat[0] == 0 and stat[1] > 0 and stat[3] >= 2
    assert np.all(np.isfinite(S.cpu().numpy()))
    G2 = np.eye(m)
    G2[3, :] = G2[:, 3] = 0.0                                                # a zero column stays zero
    S2, st2 = K.chol_inverse(K.to_padded(G2, dev))
    S2, st2 = S2.cpu().numpy(), st2.cpu().numpy()
    assert st2[0] == 0 and st2[6] == 1 and np.all(S2[3] == 0) and np.all(S2[:, 3] == 0)
    np.testing.assert_allclose(np.delete(np.delete(S2, 3, 0), 3, 1), np.eye(m - 1), atol=1e-15)
    G3 = np.full((m, m), np.nan)
    _, st3 = K.chol_inverse(K.to_padded(G3, dev))
    assert st3.cpu().numpy()[0] == 1                                         # reported, not raised: the caller checks the status


def test_device_cholesky_route_equals_host_route(env, monkeypatch):
    """The optimistic device path (no read-back until T) and the host-controlled LAPACK path give the same eigenpairs."""
    hf, K, dev = env
    M = syn.p1_mass_matrix(64)
    n = M.shape[0]
    u = syn.snapshots(n, 600, r0=96, seed=5)
    Om = syn.gaussian_omega(n, 74, seed=6)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("HFB_DEVICE_CHOL", flag)
        d, phi, Mphi, _ = proj.construct_subspace(u.copy(), 64, shifted=True, method="randomized", Omega=Om)
        out[flag] = (d, phi, proj.info.get("route"))
    assert out["1"][2] == "device" and out["0"][2] == "host"
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=1e-11)
    k = int(np.sum(out["0"][0] / out["0"][0][0] > 1e-5))
    assert subspace_angle(out["1"][1][:, :k], out["0"][1][:, :k], M) < 1e-9


# ------------------------------------------------------------------ batched Jacobi SVD
@pytest.mark.parametrize("shape", [(3, 100, 60), (2, 200, 138), (5, 30, 42), (4, 64, 1), (150, 24, 16)])
def test_jacobi_svd_batched_vs_numpy(env, shape):
    hf, K, dev = env
    batch, rows, cols = shape
    rng = np.random.default_rng(rows * cols)
    A = rng.standard_normal(shape) * (np.arange(1, cols + 1) ** -2.0)[None, None, :]
    if cols >= 3:
        A[0, :, cols // 2] = 0.0                                             # an exactly zero column
    Ad = stack(K, dev, A)
    sig, info = K.jacobi_svd_batched_(Ad)
    U, sig, info = Ad.cpu().numpy(), sig.cpu().numpy(), info.cpu().numpy()
    assert np.all(info > 0) or cols == 1, info                             # a single column needs no rotation (0 sweeps)
    for b in range(batch):
        u0, s0, _ = np.linalg.svd(A[b], full_matrices=False)
        r = min(rows, cols)
        np.testing.assert_allclose(sig[b][:r], s0, rtol=1e-12, atol=1e-14 * s0[0])
        live = sig[b] > 1e-12 * sig[b][0]
        G = U[b][:, live].T @ U[b][:, live]
        assert np.abs(G - np.eye(live.sum())).max() < 1e-12
        assert np.all(U[b][:, sig[b] == 0] == 0)
        # span: projector onto the leading singular vectors agrees with NumPy's
        kk = min(5, r)
        assert subspace_angle(U[b][:, :kk], u0[:, :kk]) < 1e-9


def test_jacobi_svd_batched_symmetric_psd_gives_eigenpairs(env):
    hf, K, dev = env
    rng = np.random.default_rng(2)
    B = rng.standard_normal((6, 138, 500)) * (np.arange(1, 139) ** -1.0)[None, :, None]
    H = np.matmul(B, np.swapaxes(B, 1, 2))
    Hd = stack(K, dev, H)
    lam, info = K.jacobi_svd_batched_(Hd)
    W, lam = Hd.cpu().numpy(), lam.cpu().numpy()
    for b in range(6):
        w0 = np.linalg.eigvalsh(H[b])[::-1]
        np.testing.assert_allclose(lam[b], w0, rtol=1e-10, atol=1e-14 * w0[0])
        assert rel(W[b] * lam[b] @ W[b].T, H[b]) < 1e-12


# ------------------------------------------------------------------ batched randomized SVD of stored Jacobians
@pytest.mark.parametrize("s", [0, 1])
def test_accuracy_enhanced_svd_batched_vs_oracle(env, s):
    hf, K, dev = env
    N, dQ, dM, k, l = 10, 100, 4096, 40, 50
    J = syn.jacobians(N, dQ, dM, r0=64, decay=1.0, seed=3)
    Om = syn.gaussian_omega(dM, l, seed=4)
    U, sig, V, sweeps = hf.accuracyEnhancedSVD_batched(J, Om, k, s=s, device=dev, return_info=True)
    U, sig, V = U.cpu().numpy(), sig.cpu().numpy(), V.cpu().numpy()
    assert np.all(sweeps.cpu().numpy() > 0)
    assert U.shape == (N, dQ, k) and sig.shape == (N, k) and V.shape == (N, dM, k)
    for i in range(N):
        U0, d0, V0 = hnp.accuracyEnhancedSVD(hnp.DenseOperator(J[i]), hnp.MultiVector.from_dense(Om), k, s=s)
        np.testing.assert_allclose(sig[i], d0, rtol=1e-9)
        assert rel((U[i] * sig[i]) @ V[i].T, (U0 * d0) @ V0.T) < 1e-9       # same rank-k approximation (signs cancel)
        np.testing.assert_allclose(U[i].T @ U[i], np.eye(k), atol=1e-10)
        np.testing.assert_allclose(V[i].T @ V[i], np.eye(k), atol=1e-8)
    if s == 1:     # with one power iteration the leading singular values are those of J_i to high accuracy
        s_true = np.linalg.svd(J[0], compute_uv=False)
        np.testing.assert_allclose(sig[0][:10], s_true[:10], rtol=1e-4)


def test_jacobian_truncated_svd_writes_the_jsvd_contract(env, golden_jtj):
    hf, K, dev = env
    J = golden_jtj["J"][:6]                                          # (6, 100, 121)
    Om = syn.gaussian_omega(121, 20, seed=3)
    U, s, V = hf.jacobian_truncated_svd(J, 10, dev, Omega=Om)
    U, s, V = U.cpu().numpy(), s.cpu().numpy(), V.cpu().numpy()
    assert U.shape == (6, 100, 10) and s.shape == (6, 10) and V.shape == (6, 121, 10)
    for i in range(6):
        U0, d0, V0 = hnp.accuracyEnhancedSVD(hnp.DenseOperator(J[i]), hnp.MultiVector.from_dense(Om), 10, s=1)
        np.testing.assert_allclose(s[i], d0, rtol=1e-9)
        assert rel((U[i] * s[i]) @ V[i].T, (U0 * d0) @ V0.T) < 1e-9
        s0 = np.linalg.svd(J[i], compute_uv=False)
        np.testing.assert_allclose(s[i][:3], s0[:3], rtol=1e-4)          # the randomized factors track the true SVD
        assert np.all(np.diff(s[i]) <= 0)
        np.testing.assert_allclose(V[i].T @ V[i], np.eye(10), atol=1e-8)
        np.testing.assert_allclose(U[i].T @ U[i], np.eye(10), atol=1e-10)
    U2, s2, V2 = hf.jacobian_truncated_svd(J, 10, dev)                   # Omega drawn on the device (seeded)
    assert tuple(s2.shape) == (6, 10) and bool(torch.all(s2[:, :-1] >= s2[:, 1:]))


# ------------------------------------------------------------------ output active subspace: dense and operator form
@pytest.mark.parametrize("shape", [(16, 100, 121), (8, 3000, 64)])
def test_output_subspace_operator_form_equals_dense_and_oracle(env, shape):
    """(8, 3000, 64) is the full-state case in miniature: dQ = n_u >> dM (activeSubspaceProjector.py:625-673,
    fullStateObservable.py:18)."""
    hf, K, dev = env
    N, dQ, dM = shape
    J = syn.jacobians(N, dQ, dM, r0=min(32, dM), seed=13)
    rank = 20
    Om = syn.gaussian_omega(dQ, rank + 10, seed=9)
    d0, U0 = P.as_output_from_jacobians(J, rank, Om)
    k = int(np.sum(d0 / d0[0] > 1e-5))
    res = {}
    for form in (False, True):
        params = hf.ActiveSubspaceParameterList()
        params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = rank, 10, False, False
        proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J), None, parameters=params, device=dev)
        proj.Omega_NG = Om
        d, dec, enc = proj.construct_output_subspace(operator_form=form)
        np.testing.assert_allclose(d[:k], d0[:k], rtol=1e-10)
        assert subspace_angle(hf.mv_to_dense(dec)[:, :k], U0[:, :k]) < 1e-8
        res[form] = d
    np.testing.assert_allclose(res[True][:k], res[False][:k], rtol=1e-11)
    # the operator itself, one vector and a block, against einsum
    op = hf.MeanJJTfromDataOperator(J, device=dev, chunk_bytes=3 * dM * 32 * 8)      # forces several chunks
    X = hf.DeviceMultiVector.from_dense(Om, dev)
    Y = hf.DeviceMultiVector(dQ, Om.shape[1], device=dev)
    op.matMvMult(X, Y)
    ref = np.einsum("iqm,imk->qk", J, np.einsum("iqm,qk->imk", J, Om)) / N
    assert rel(Y.to_dense(), ref) < 1e-12
    x, y = hf.DeviceVector(dQ, dev), hf.DeviceVector(dQ, dev)
    x.set_local(Om[:, 0])
    op.mult(x, y)
    assert rel(y.get_local(), ref[:, 0]) < 1e-12


# ------------------------------------------------------------------ prior.Hlr branch of the input subspace
def test_input_subspace_with_low_rank_hessian_prior_vs_oracle(env, golden_jtj, golden_dp):
    """doublePassG(A, prior.Hlr, prior.Hlr, Omega, rank) (activeSubspaceProjector.py:455-459): Hlr = R + R U D U^T R is the
    weighting operator AND its own solver."""
    import scipy.sparse as sp
    hf, K, dev = env
    J, Om = golden_jtj["J"], golden_dp["Omega_as"]
    R = syn.p1_mass_matrix(10)
    n = R.shape[0]
    rng = np.random.default_rng(1)
    A0 = rng.standard_normal((n, 6))
    L = np.linalg.cholesky(A0.T @ (R @ A0))
    U = np.linalg.solve(L, A0.T).T                                   # U^T R U = I
    dl = np.array([5.0, 3.0, 2.0, 1.0, 0.5, 0.1])
    H = R.toarray() + (R @ U) @ np.diag(dl) @ (R @ U).T

    class Prior:
        pass

    base = hf.SparsePrior(R, device=dev)
    prior = Prior()
    prior.device = dev
    prior.Hlr = hf.LowRankHessian(base, dl, hf.DeviceMultiVector.from_dense(U, dev))
    # the operator and its solver against the dense matrix
    X = rng.standard_normal((n, 7))
    Xd = K.to_padded(X, dev)
    assert rel(prior.Hlr.matmat(Xd).cpu().numpy(), H @ X) < 1e-13
    assert rel(prior.Hlr.solve_block(Xd).cpu().numpy(), np.linalg.solve(H, X)) < 1e-11
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 32, 10, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J), prior, parameters=params, device=dev)
    proj.Omega_GN = Om[:, :42]
    d, dec, enc = proj.construct_input_subspace(prior_preconditioned=True)
    Hs = sp.csr_matrix(H)
    d0, V0, E0 = P.as_input_from_jacobians(J, 32, Om[:, :42], B_csr=Hs)
    k = int(np.sum(d0 / d0[0] > 1e-5))
    np.testing.assert_allclose(d[:k], d0[:k], rtol=1e-10)
    V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
    assert subspace_angle(V[:, :k], V0[:, :k], Hs) < 1e-8
    assert rel(E, H @ V) < 1e-12
    np.testing.assert_allclose(V.T @ E, np.eye(32), atol=1e-9)
