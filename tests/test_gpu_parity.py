"""GPU parity tests proper: the CUDA path (through the public projector API and the C ABI) against the CPU
oracle on the same seeded inputs and the same Omega, against the committed golden vectors, and through
size-independent properties at benchmark widths.

Tolerances (BASELINE.json north_star, fp64): eigenvalues relative 1e-10; largest principal angle between the
leading subspaces < 1e-8 (in the M inner product where weighted); projected data relative Frobenius 1e-12.
"""
import numpy as np
import pytest
import torch

from hippyflow_b200 import synthetic as syn
from oracle import projectors_np as P
from conftest import subspace_angle

pytestmark = pytest.mark.gpu

EIG_RTOL = 1e-10
ANGLE_TOL = 1e-8
PROJ_RTOL = 1e-12


@pytest.fixture(scope="module")
def hf(cuda_device):
    import hippyflow_b200 as hf
    from hippyflow_b200 import _lib
    _lib.lib()
    return hf


def leading(d, floor=1e-5):
    """Number of leading modes whose eigenvalue ratio to the first is above `floor`: the subspace-angle
    criterion is only meaningful away from round-off-level eigenvalues."""
    return int(np.sum(d / d[0] > floor))


# ------------------------------------------------------------------ POD, M-weighted randomized (config 1)
@pytest.mark.parametrize("shifted", [True, False])
def test_pod_randomized_weighted_vs_oracle(hf, cuda_device, golden_pod, golden_dp, shifted):
    M = syn.p1_mass_matrix(int(golden_pod["nx"]))
    u, Om = golden_pod["u_data"], golden_dp["Omega"]
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 15, shifted=shifted, method="randomized", Omega=Om)
    d0, U0, E0, s0 = P.pod_randomized_weighted(u, M, 15, Om, shifted=shifted)
    np.testing.assert_allclose(d, d0, rtol=EIG_RTOL)
    k = leading(d0)
    assert subspace_angle(phi[:, :k], U0[:, :k], M) < ANGLE_TOL
    np.testing.assert_allclose(shift, s0, rtol=0, atol=1e-14)
    assert phi.shape == (289, 15) and Mphi.shape == (289, 15) and shift.shape == (289,) and d.shape == (15,)
    # reference property tests (test_PODProjector.py:161-174)
    I = np.eye(15)
    assert np.linalg.norm(I - phi.T @ Mphi) / np.linalg.norm(I) < 1e-8
    assert np.linalg.norm(M @ phi - Mphi) / np.linalg.norm(Mphi) < 1e-8
    if not shifted:   # golden: reference CollectiveOperator(NullCollective) + doublePassG
        np.testing.assert_allclose(d, golden_dp["d_w"], rtol=EIG_RTOL)
        assert subspace_angle(phi[:, :k], golden_dp["U_w"][:, :k], M) < ANGLE_TOL


def test_pod_randomized_faithful_equals_shortcut(hf, cuda_device, golden_pod, golden_dp):
    """T = (AQ)^T Q (hIPPYlib's form) and T = W^T W / N give the same eigenpairs."""
    M = syn.p1_mass_matrix(int(golden_pod["nx"]))
    u, Om = golden_pod["u_data"], golden_dp["Omega"]
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d1, phi1, _, _ = proj.construct_subspace(u.copy(), 15, shifted=False, method="randomized", Omega=Om)
    d2, phi2, _, _ = proj.construct_subspace(u.copy(), 15, shifted=False, method="randomized", Omega=Om, faithful=True)
    np.testing.assert_allclose(d1, d2, rtol=EIG_RTOL)
    k = leading(d1)
    assert subspace_angle(phi1[:, :k], phi2[:, :k], M) < ANGLE_TOL


# ------------------------------------------------------------------ POD, deterministic 'hep' vs reference verbatim
@pytest.mark.parametrize("method", ["hep", "ghep", "inverse_ghep"])
@pytest.mark.parametrize("shifted", [True, False])
def test_pod_hep_vs_reference_golden(hf, cuda_device, golden_pod, shifted, method):
    g = golden_pod
    M = syn.p1_mass_matrix(int(g["nx"]))
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(g["u_data"].copy(), 15, shifted=shifted, method=method)
    key = f"{method}_{int(shifted)}"
    np.testing.assert_allclose(d, g[key + "_d"], rtol=1e-9)
    np.testing.assert_allclose(shift, g[key + "_shift"], rtol=0, atol=1e-14)
    assert subspace_angle(phi[:, :10], g[key + "_phi"][:, :10], M) < 1e-7
    rv = 14 if shifted else 15
    I = np.eye(rv)
    assert np.linalg.norm(I - phi[:, :rv].T @ Mphi[:, :rv]) / np.linalg.norm(I) < 1e-8
    assert np.linalg.norm(M @ phi - Mphi) / np.linalg.norm(Mphi) < 1e-8
    # eigen-relation C M phi_i = d_i phi_i (test_PODProjector.py:188-208, tolerance 1e-2)
    X = g["u_data"] - (shift if shifted else 0.0)
    C = X.T @ X / X.shape[0]
    for i in range(rv):
        lhs = C @ (M @ phi[:, i])
        assert np.linalg.norm(lhs - d[i] * phi[:, i]) / np.linalg.norm(d[i] * phi[:, i]) < 1e-2


def test_pod_ghep_dense_pencil_when_more_samples_than_dofs(hf, cuda_device):
    """n_data > dim_u: the (n x n) pencil route; checked against the restated reference (ARPACK 'ghep')."""
    M = syn.p1_mass_matrix(6)                                     # 49 dofs
    u = syn.snapshots(49, 200, r0=30, seed=4)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 8, shifted=True, method="ghep")
    d0, phi0, Mphi0, s0 = P.pod_from_data(u.copy(), M, 8, shifted=True, method="ghep")
    np.testing.assert_allclose(d, d0, rtol=1e-9)
    assert subspace_angle(phi[:, :6], phi0[:, :6], M) < 1e-7
    np.testing.assert_allclose(phi.T @ Mphi, np.eye(8), atol=1e-10)


def test_pod_unavailable_method_and_rank_check(hf, cuda_device, golden_pod):
    M = syn.p1_mass_matrix(int(golden_pod["nx"]))
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    with pytest.raises(ValueError, match="Unavailable method"):
        proj.construct_subspace(golden_pod["u_data"], 15, method="nope")
    with pytest.raises(AssertionError):
        proj.construct_subspace(golden_pod["u_data"][:10], 15, method="randomized")


# ------------------------------------------------------------------ PODProjector (doublePass, collective 'avg')
def test_pod_doublepass_vs_golden(hf, cuda_device, golden_pod, golden_dp):
    u, Om = golden_pod["u_data"], golden_dp["Omega"]
    params = hf.PODParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 15, 10, False, False
    proj = hf.PODProjector(hf.StoredSnapshots(u), collective=hf.NullCollective(), parameters=params, device=cuda_device)
    d, U = proj.construct_subspace(Omega=Om)
    np.testing.assert_allclose(d, golden_dp["d_pod"], rtol=EIG_RTOL)
    Ud = hf.mv_to_dense(U)
    k = leading(golden_dp["d_pod"])
    assert subspace_angle(Ud[:, :k], golden_dp["U_pod"][:, :k]) < ANGLE_TOL
    np.testing.assert_allclose(Ud.T @ Ud, np.eye(15), atol=1e-12)


# ------------------------------------------------------------------ active subspace from stored Jacobians
def test_meanjtj_operator_vs_reference_golden(hf, cuda_device, golden_jtj):
    from oracle.hippylib_np import Vector
    g = golden_jtj
    for G, ys in ((None, g["ys"]), (g["noise_cov_inv"], g["ysG"])):
        op = hf.MeanJTJfromDataOperator(g["J"], None, G, device=cuda_device)
        assert (op.ndata, op.r, op.dM) == (64, 100, 121)
        for x, yref in zip(g["xs"], ys):
            y = Vector(np.zeros(121))
            op.mult(Vector(x.copy()), y)
            assert np.linalg.norm(y.get_local() - yref) / np.linalg.norm(yref) < PROJ_RTOL
            y2 = Vector(np.zeros(121))
            op.transpmult(Vector(x.copy()), y2)
            np.testing.assert_array_equal(y.get_local(), y2.get_local())


@pytest.mark.parametrize("preconditioned", [False, True])
def test_active_subspace_vs_golden(hf, cuda_device, golden_jtj, golden_dp, preconditioned):
    J, Om = golden_jtj["J"], golden_dp["Omega_as"]
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 64, 10, False, False
    Mp = syn.p1_mass_matrix(10)
    prior = hf.SparsePrior(Mp, device=cuda_device) if preconditioned else None
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J), prior, collective=hf.NullCollective(), parameters=params,
                                      device=cuda_device)
    proj.Omega_GN = Om
    d, dec, enc = proj.construct_input_subspace(prior_preconditioned=preconditioned)
    dref = golden_dp["d_asg"] if preconditioned else golden_dp["d_as"]
    Vref = golden_dp["V_asg"] if preconditioned else golden_dp["V_as"]
    k = leading(dref, 1e-5)
    np.testing.assert_allclose(d[:k], dref[:k], rtol=EIG_RTOL)
    np.testing.assert_allclose(d, dref, rtol=0, atol=1e-10 * dref[0])   # absolute accuracy ~ eps * lambda_1
    V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
    assert subspace_angle(V[:, :k], Vref[:, :k], Mp if preconditioned else None) < ANGLE_TOL
    if preconditioned:
        np.testing.assert_allclose(E, Mp @ V, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(V.T @ E, np.eye(64), atol=1e-10)
    else:
        np.testing.assert_array_equal(V, E)
    assert proj.prior_preconditioned == preconditioned and proj.d_GN is d


def test_active_subspace_noise_weighted_vs_oracle(hf, cuda_device, golden_jtj, golden_dp):
    J, G, Om = golden_jtj["J"], golden_jtj["noise_cov_inv"], golden_dp["Omega_as"]
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["verbose"], params["save_and_plot"] = 32, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J, G), None, parameters=params, device=cuda_device)
    proj.Omega_GN = Om[:, :42]
    d, dec, _ = proj.construct_input_subspace(prior_preconditioned=False)
    d0, V0, _ = P.as_input_from_jacobians(J, 32, Om[:, :42], noise_cov_inv=G)
    k = leading(d0, 1e-5)
    np.testing.assert_allclose(d[:k], d0[:k], rtol=EIG_RTOL)
    assert subspace_angle(hf.mv_to_dense(dec)[:, :k], V0[:, :k]) < ANGLE_TOL


def test_active_output_subspace_vs_oracle(hf, cuda_device, golden_jtj):
    J = golden_jtj["J"][:16]
    Om = syn.gaussian_omega(100, 30, seed=9)
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 20, 10, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J), None, parameters=params, device=cuda_device)
    proj.Omega_NG = Om
    d, dec, _ = proj.construct_output_subspace()
    d0, U0 = P.as_output_from_jacobians(J, 20, Om)
    k = leading(d0, 1e-5)
    np.testing.assert_allclose(d[:k], d0[:k], rtol=EIG_RTOL)
    assert subspace_angle(hf.mv_to_dense(dec)[:, :k], U0[:, :k]) < ANGLE_TOL


# ------------------------------------------------------------------ KLE from stored parameter draws
def test_kle_mass_vs_golden(hf, cuda_device, golden_pod, golden_dp):
    M = syn.p1_mass_matrix(int(golden_pod["nx"]))
    n = M.shape[0]
    m_data = syn.snapshots(n, 512, r0=200, decay=2.0, eps=1e-9, seed=int(golden_dp["m_data_seed"]))
    params = hf.KLEParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 128, 10, False, False
    proj = hf.KLEProjector(hf.SampleCovariancePrior(m_data, M, device=cuda_device), parameters=params)
    d, dec, enc = proj.construct_input_subspace("mass", Omega=golden_dp["Omega_kle"])
    dref = golden_dp["d_kle"]
    k = leading(dref)
    np.testing.assert_allclose(d[:k], dref[:k], rtol=EIG_RTOL)
    V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
    assert subspace_angle(V[:, :k], golden_dp["V_kle"][:, :k], M) < ANGLE_TOL
    r = 128
    assert np.linalg.norm(V.T @ E - np.eye(r)) / np.sqrt(r) < 1e-10            # test_KLEProjector.py:97-99
    assert np.linalg.norm(M @ V - E) / np.linalg.norm(E) < 1e-10                 # :102-108
    assert proj.M_orthogonal is True
    d2, dec2, enc2 = proj.construct_input_subspace("identity", Omega=golden_dp["Omega_kle"])
    d0, V0, _ = P.kle_from_samples(m_data, M, 128, golden_dp["Omega_kle"], "identity")
    k2 = leading(d0)
    np.testing.assert_allclose(d2[:k2], d0[:k2], rtol=EIG_RTOL)
    assert subspace_angle(hf.mv_to_dense(dec2)[:, :k2], V0[:, :k2]) < ANGLE_TOL
    np.testing.assert_array_equal(hf.mv_to_dense(dec2), hf.mv_to_dense(enc2))
    with pytest.raises(NotImplementedError):
        proj.construct_input_subspace("prior")


# ------------------------------------------------------------------ projection of stored data
def test_projection_of_stored_data_vs_oracle(hf, cuda_device, golden_jtj):
    rng = np.random.default_rng(0)
    M = syn.p1_mass_matrix(10)                       # 121 parameter dofs
    m_data = rng.standard_normal((300, 121))
    V = np.linalg.qr(rng.standard_normal((121, 20)))[0]
    enc = M @ V
    red = hf.project_data(m_data, enc, cuda_device).cpu().numpy()
    ref = P.project_data(m_data, enc)
    assert np.linalg.norm(red - ref) / np.linalg.norm(ref) < PROJ_RTOL
    J = golden_jtj["J"][:24]                         # (24, 100, 121)
    Phi = np.linalg.qr(rng.standard_normal((100, 12)))[0]
    a = hf.jacobian_action(J, V, cuda_device).cpu().numpy()
    assert np.linalg.norm(a - P.j_psi(J, V)) / np.linalg.norm(P.j_psi(J, V)) < PROJ_RTOL
    b = hf.jacobian_transpose_action(J, Phi, cuda_device).cpu().numpy()
    assert np.linalg.norm(b - P.jstar_phi(J, Phi)) / np.linalg.norm(P.jstar_phi(J, Phi)) < PROJ_RTOL
    c = hf.reduced_jacobians(J, Phi, V, cuda_device).cpu().numpy()
    refc = P.reduced_jacobians(J, Phi, V)
    assert np.linalg.norm(c - refc) / np.linalg.norm(refc) < PROJ_RTOL


# ------------------------------------------------------------------ mid-size parity (oracle finishes in seconds)
def test_pod_randomized_midsize_vs_blocked_oracle(hf, cuda_device):
    """n = 66049 (257^2 P1), N = 512, rank 64 + 10: eigenvalues against a blocked NumPy evaluation of the same
    double-pass algorithm (d and span(U) do not depend on the choice of M-orthonormal basis of span(Y))."""
    nx = 256
    M = syn.p1_mass_matrix(nx)
    n = M.shape[0]
    u = syn.snapshots(n, 512, r0=128, decay=1.0, eps=1e-6, seed=2)
    Om = syn.gaussian_omega(n, 74, seed=3)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u, 64, shifted=True, method="randomized", Omega=Om)
    X = u - u.mean(0)
    Y = X.T @ (X @ (M @ Om)) / X.shape[0]
    G = Y.T @ (M @ Y)
    w, V = np.linalg.eigh((G + G.T) / 2)
    Q = Y @ (V / np.sqrt(w))
    G2 = Q.T @ (M @ Q)
    w2, V2 = np.linalg.eigh((G2 + G2.T) / 2)
    Q = Q @ (V2 / np.sqrt(w2))
    W = X @ (M @ Q)
    T = W.T @ W / X.shape[0]
    dd, VV = np.linalg.eigh(T)
    d0 = dd[::-1][:64]
    U0 = Q @ VV[:, ::-1][:, :64]
    np.testing.assert_allclose(d, d0, rtol=EIG_RTOL)
    k = leading(d0)
    assert subspace_angle(phi[:, :k], U0[:, :k], M) < ANGLE_TOL
    assert np.linalg.norm(phi.T @ Mphi - np.eye(64)) < 1e-10


# ------------------------------------------------------------------ on-disk data contract
def test_data_contract_roundtrip(hf, cuda_device, tmp_path, golden_jtj):
    from hippyflow_b200 import dataIO
    rng = np.random.default_rng(3)
    N, dM, dQ, r = 24, 121, 30, 6
    m_data, q_data = rng.standard_normal((N, dM)), rng.standard_normal((N, dQ))
    np.savez_compressed(tmp_path / "mq_data.npz", m_data=m_data, q_data=q_data)
    m0, q0 = dataIO.load_mq_data(str(tmp_path), cuda_device, rank=1, world=2)
    np.testing.assert_array_equal(m0.cpu().numpy(), m_data[12:24])
    np.testing.assert_array_equal(q0.cpu().numpy(), q_data[12:24])
    # low-rank stored Jacobians: the SVD factor reproduces mean(J^T J) exactly
    U = np.linalg.qr(rng.standard_normal((N, dQ, r)))[0]
    V = np.linalg.qr(rng.standard_normal((N, dM, r)))[0]
    s = np.abs(rng.standard_normal((N, r))) + 0.1
    np.savez_compressed(tmp_path / "Jsvd_data.npz", U_data=U, sigma_data=s, V_data=V)
    J = np.einsum("iqr,ir,imr->iqm", U, s, V)
    Xt, blk = dataIO.load_jacobian_svd_factor(str(tmp_path), cuda_device)
    assert blk == r and tuple(Xt.shape) == (N * r, dM)
    from hippyflow_b200.linalg import SampleCovariance
    Om = syn.gaussian_omega(dM, 9, seed=1)
    from hippyflow_b200 import _lib as K
    Y = SampleCovariance(Xt, block=r).apply(K.to_padded(Om, cuda_device)).cpu().numpy()
    ref = np.einsum("iqm,iqk->mk", J, J @ Om) / N
    assert np.linalg.norm(Y - ref) / np.linalg.norm(ref) < PROJ_RTOL
    # reduced data set + projector file names
    enc_in = np.linalg.qr(rng.standard_normal((dM, 5)))[0]
    enc_out = np.linalg.qr(rng.standard_normal((dQ, 4)))[0]
    shift = q_data.mean(0)
    m_r, q_r = dataIO.reduce_dataset(str(tmp_path), enc_in, enc_out, cuda_device, q_shift=shift, out_name="mq_reduced.npz")
    np.testing.assert_allclose(m_r.cpu().numpy(), m_data @ enc_in, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(q_r.cpu().numpy(), (q_data - shift) @ enc_out, rtol=1e-12, atol=1e-14)
    z = np.load(tmp_path / "mq_reduced.npz")
    assert z["m_data"].shape == (N, 5) and z["q_data"].shape == (N, 4)
    dataIO.save_pod(str(tmp_path), enc_out, np.arange(4.0))
    dataIO.save_kle(str(tmp_path), enc_in, np.arange(5.0))
    dataIO.save_active_subspace(str(tmp_path), enc_in, np.arange(5.0), 128)
    for f in ("POD_projector.npy", "POD_d.npy", "KLE_decoder.npy", "KLE_d.npy", "AS_128_input_decoder.npy", "AS_128_d_GN.npy"):
        assert (tmp_path / f).exists(), f


# ------------------------------------------------------------------ batched vs stacked operator forms, edge cases
def test_batched_list_and_stacked_operator_give_equal_eigenvalues(hf, cuda_device, golden_jtj, golden_dp):
    """The invariant of test_derivativeSubspace.py:83-102: two ways of applying the sample-averaged operator
    (a SummedListOperator of per-sample J^T J wrapped in CollectiveOperator vs the stacked one-GEMM-pair form in
    MatrixMultCollectiveOperator) give the same eigenvalues with the same Omega, ||d1 - d2||_2 < 1e-12."""
    J = golden_jtj["J"][:16]
    Om = hf.DeviceMultiVector.from_dense(golden_dp["Omega_as"][:, :30], cuda_device)
    coll = hf.NullCollective()
    batched = hf.CollectiveOperator(hf.SummedListOperator([hf.JTJ(J[i], cuda_device) for i in range(16)], average=True), coll, "avg")
    stacked = hf.MatrixMultCollectiveOperator(hf.MeanJTJfromDataOperator(J, None, None, device=cuda_device), coll, "avg")
    d1, U1 = hf.doublePass(batched, Om, 20, s=1)
    d2, U2 = hf.doublePass(stacked, Om, 20, s=1, faithful=True)
    d3, U3 = hf.doublePass(stacked.local_op, Om, 20, s=1)
    assert np.linalg.norm(d1 - d2) < 1e-12 * max(1.0, d1[0])
    assert np.linalg.norm(d1 - d3) < 1e-12 * max(1.0, d1[0])


def test_rank_deficient_snapshots_and_extreme_ranks(hf, cuda_device):
    """Edge cases: fewer independent snapshots than requested modes (trailing eigenvalues are round-off, nothing is
    NaN), u_rank == n_data, no oversampling, a single snapshot."""
    M = syn.p1_mass_matrix(9)                        # 100 dofs
    n = M.shape[0]
    rng = np.random.default_rng(0)
    base = rng.standard_normal((5, n))
    u = rng.standard_normal((30, 5)) @ base            # 30 snapshots of rank 5
    Om = syn.gaussian_omega(n, 22, seed=2)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 12, shifted=False, method="randomized", Omega=Om)
    d0, U0, _, _ = P.pod_randomized_weighted(u, M, 12, Om, shifted=False)
    assert np.all(np.isfinite(d)) and np.all(np.isfinite(phi))
    np.testing.assert_allclose(d[:5], d0[:5], rtol=EIG_RTOL)
    assert np.all(np.abs(d[5:]) < 1e-12 * d[0])
    assert subspace_angle(phi[:, :5], U0[:, :5], M) < ANGLE_TOL
    # u_rank == n_data with zero oversampling
    u2 = syn.snapshots(n, 8, r0=8, seed=5)
    d2, phi2, _, _ = proj.construct_subspace(u2.copy(), 8, shifted=False, method="randomized", oversampling=0,
                                             Omega=Om[:, :8])
    d20, U20, _, _ = P.pod_randomized_weighted(u2, M, 8, Om[:, :8], shifted=False)
    k = leading(d20)
    np.testing.assert_allclose(d2[:k], d20[:k], rtol=EIG_RTOL)
    # deterministic route with u_rank == n_data
    d3, phi3, Mphi3, _ = proj.construct_subspace(u2.copy(), 8, shifted=False, method="hep")
    d30, _, _, _ = P.pod_from_data(u2.copy(), M, 8, shifted=False, method="hep")
    np.testing.assert_allclose(d3, d30, rtol=1e-9)
    # one snapshot, rank 1
    d4, phi4, Mphi4, _ = proj.construct_subspace(u2[:1].copy(), 1, shifted=False, method="randomized", oversampling=2,
                                                 Omega=Om[:, :3])
    x = u2[0]
    np.testing.assert_allclose(d4[0], x @ (M @ x), rtol=1e-10)


def test_projector_files_written_with_reference_names(hf, cuda_device, tmp_path, golden_pod, golden_dp):
    out = str(tmp_path) + "/"
    params = hf.PODParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["output_directory"] = 15, 10, False, out
    proj = hf.PODProjector(hf.StoredSnapshots(golden_pod["u_data"]), parameters=params, device=cuda_device)
    proj.construct_subspace(Omega=golden_dp["Omega"])
    U = np.load(out + "POD_projector.npy")
    np.testing.assert_allclose(np.load(out + "POD_d.npy"), golden_dp["d_pod"], rtol=EIG_RTOL)   # PODProjector.py:383-384
    assert U.shape == (289, 15)


# ------------------------------------------------------------------ projection-error sweep and per-sample Jacobian SVD (8(f) ranks 2, 3)
def test_projection_error_sweep_vs_numpy(hf, cuda_device):
    rng = np.random.default_rng(4)
    M = syn.p1_mass_matrix(10)
    n = M.shape[0]
    test = syn.snapshots(n, 37, r0=30, seed=8)
    # an M-orthonormal basis and its encoder
    A = rng.standard_normal((n, 20))
    L = np.linalg.cholesky(A.T @ (M @ A))
    V = np.linalg.solve(L, A.T).T
    E = M @ V
    ranks = [5, 1, 20, 12]
    avg, std = hf.projection_errors(test, V, E, ranks, device=cuda_device)
    for i, r in enumerate(sorted(ranks)):
        rec = (test @ E[:, :r]) @ V[:, :r].T
        rel = np.linalg.norm(test - rec, axis=1) / np.linalg.norm(test, axis=1)      # KLEProjector.py:262-270
        np.testing.assert_allclose(avg[i], rel.mean(), rtol=1e-10)
        np.testing.assert_allclose(std[i], rel.std(), rtol=1e-8, atol=1e-14)
    # PriorPreconditionedProjector.mult: y = U U^T C^-1 x   (priorPreconditionedProjector.py:48-55)
    from hippyflow_b200.linalg import CsrMatrix
    Pj = hf.PriorPreconditionedProjector(hf.DeviceMultiVector.from_dense(V, cuda_device), CsrMatrix(M, cuda_device))
    X = hf.DeviceMultiVector.from_dense(test.T.copy(), cuda_device)
    Y = hf.DeviceMultiVector(n, 37, device=cuda_device)
    Pj.matMvMult(X, Y)
    ref = V @ (V.T @ (M @ test.T))
    assert np.linalg.norm(Y.to_dense() - ref) / np.linalg.norm(ref) < PROJ_RTOL


# ------------------------------------------------------------------ mean shift: implicit / pipelined / explicit routes
def _blocked_weighted_pod(u, M, Om, rank):
    """Blocked NumPy evaluation of the M-weighted double pass on explicitly shifted data (the reference's
    u_data - np.mean(u_data, axis=0), PODProjector.py:732-734)."""
    X = u - u.mean(0)
    Y = X.T @ (X @ (M @ Om)) / X.shape[0]
    Q = Y
    for _ in range(2):
        G = Q.T @ (M @ Q)
        w, V = np.linalg.eigh((G + G.T) / 2)
        Q = Q @ (V / np.sqrt(w))
    W = X @ (M @ Q)
    dd, VV = np.linalg.eigh(W.T @ W / X.shape[0])
    return dd[::-1][:rank], Q @ VV[:, ::-1][:, :rank]


@pytest.mark.parametrize("mean_scale", [0.0, 30.0])
@pytest.mark.parametrize("route", ["pipelined", "resident_implicit", "resident_explicit", "host_unpipelined"])
def test_pod_randomized_mean_shift_routes(hf, cuda_device, route, mean_scale):
    """All routes of the mean shift give the reference's eigenpairs, also when the mean is 30x larger than the
    fluctuations: (i) host input, upload pipelined in 4 chunks with a provisional mean and the lift accumulated per
    chunk; (ii) device-resident input, shift applied inside the products (input array must stay untouched);
    (iii) device-resident input shifted explicitly on a copy; (iv) host input without the pipelined upload."""
    from hippyflow_b200 import _lib as K
    nx = 64
    M = syn.p1_mass_matrix(nx)
    n = M.shape[0]
    N, rank = 1024, 32
    u = syn.snapshots(n, N, r0=96, decay=1.0, eps=1e-6, seed=5)
    rms = np.sqrt(np.mean(u ** 2))
    u = u + mean_scale * rms * (1.0 + 0.3 * np.sin(np.arange(n) * 0.01))[None, :]
    Om = syn.gaussian_omega(n, rank + 10, seed=6)
    d0, U0 = _blocked_weighted_pod(u, M, Om, rank)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    if route == "pipelined":
        d, phi, Mphi, shift = proj.construct_subspace(u.copy(), rank, shifted=True, method="randomized", Omega=Om)
    elif route == "host_unpipelined":
        d, phi, Mphi, shift = proj.construct_subspace(u.copy(), rank, shifted=True, method="randomized", Omega=Om,
                                                      pipelined_upload=False)
    else:
        ud = K.to_padded(u, cuda_device)
        keep = ud.clone()
        d, phi, Mphi, shift = proj.construct_subspace(ud, rank, shifted=True, method="randomized", Omega=Om,
                                                      implicit_shift=(route == "resident_implicit"))
        assert torch.equal(ud, keep)                       # the caller's device array is never modified
        assert proj.shift_route.startswith('implicit' if route == "resident_implicit" else 'explicit')
    np.testing.assert_allclose(d, d0, rtol=EIG_RTOL)
    k = leading(d0)
    assert subspace_angle(phi[:, :k], U0[:, :k], M) < ANGLE_TOL
    np.testing.assert_allclose(shift, u.mean(0), rtol=1e-13, atol=1e-14 * max(1.0, mean_scale))
    assert np.linalg.norm(phi.T @ Mphi - np.eye(rank)) < 1e-10


def test_pod_randomized_dominant_mean_falls_back_to_explicit_shift(hf, cuda_device):
    """|mean| = 1e5 x fluctuations: (|mean|/rms)^2 = 1e10 is above IMPLICIT_SHIFT_MAX_RATIO, the implicit route must hand
    over to the explicit shift and still resolve the fluctuation eigenpairs."""
    from hippyflow_b200 import _lib as K
    M = syn.p1_mass_matrix(64)
    n = M.shape[0]
    u = syn.snapshots(n, 512, r0=64, decay=1.0, eps=1e-6, seed=8)
    u = u + 1.0e5 * np.sqrt(np.mean(u ** 2))
    Om = syn.gaussian_omega(n, 26, seed=9)
    d0, U0 = _blocked_weighted_pod(u, M, Om, 16)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    ud = K.to_padded(u, cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(ud, 16, shifted=True, method="randomized", Omega=Om)
    # the data themselves carry only ~11 digits of the fluctuations here (eps * 1e5), so the comparison is looser
    assert proj.shift_route == 'explicit-fallback'
    np.testing.assert_allclose(d, d0, rtol=1e-7)
    assert subspace_angle(phi[:, :8], U0[:, :8], M) < 1e-6

def test_weighted_l2_norm_vector_direct(hf, cuda_device):
    """sqrt(diag(x^T W x)) per column (PODProjector.py:658-661) against SciPy, for one column and for blocks on either side
    of the SpMM kernel switch."""
    from hippyflow_b200 import _lib as K
    from hippyflow_b200.linalg import CsrMatrix
    from hippyflow_b200.modeling import weighted_l2_norm_vector
    W = syn.p1_mass_matrix(70, 63).tocsr()
    rng = np.random.default_rng(5)
    Wd = CsrMatrix(W, cuda_device)
    for r in (1, 7, 74):
        x = rng.standard_normal((W.shape[0], r))
        got = weighted_l2_norm_vector(K.to_padded(x, cuda_device), Wd).cpu().numpy()
        np.testing.assert_allclose(got, np.sqrt(np.einsum("ij,ij->j", x, W @ x)), rtol=1e-13)
