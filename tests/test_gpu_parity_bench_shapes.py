"""GPU parity at the BENCHMARK shapes (BASELINE.json configs[1..4]): the code paths bench.py times -- NT = 17 tiles at
m = 266, split-K, the fragment-record SpMM at m = 266 / 210 / 138, the deferred clean-up pass, the implicit and pipelined
mean shift -- compared eigenpair by eigenpair with the blocked CPU oracle (oracle/projectors_np.py, pinned against the
column-by-column port and the reference-driven golden vectors in tests/test_oracle_golden.py) on the same seeded inputs
and the same Omega.  Sample counts are reduced so that the oracle finishes in seconds; dofs, widths and ranks are the
benchmark's.

Tolerances (BASELINE.json north_star, fp64): eigenvalues relative 1e-10; largest principal angle between the leading
subspaces < 1e-8 (M inner product where weighted); projected data relative Frobenius 1e-12.  "Leading" = modes with
lambda_i / lambda_1 > 1e-5; modes below are compared with the absolute tolerance 1e-10 * lambda_1 (a backward-stable
method resolves lambda_i to ~eps * lambda_1 absolute)."""
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
    return int(np.sum(d / d[0] > floor))


def check_eigs(d, d0):
    k = leading(d0)
    np.testing.assert_allclose(d[:k], d0[:k], rtol=EIG_RTOL)
    np.testing.assert_allclose(d, d0, rtol=0, atol=1e-10 * d0[0])
    return k


# ------------------------------------------------------------------ cfg2: confusion output POD, n = 263,169, rank 256 + 10
@pytest.fixture(scope="module")
def cfg2_case():
    n, N, rank, p = 263169, 1024, 256, 10
    M = syn.p1_mass_matrix_for(n)
    u = syn.snapshots(n, N, r0=512, decay=1.0, eps=1e-6, seed=7)
    u += 0.5 * np.cos(np.linspace(0.0, 3.0, n))[None, :]                    # a mean of the size of the fluctuations
    Om = syn.gaussian_omega(n, rank + p, seed=1)
    d0, U0, E0, s0 = P.pod_randomized_weighted_blocked(u, M, rank, Om, shifted=True)
    return dict(n=n, N=N, rank=rank, M=M, u=u, Om=Om, d0=d0, U0=U0, E0=E0, s0=s0)


@pytest.mark.parametrize("entry", ["host_pipelined", "device_implicit"])
def test_cfg2_shape_weighted_pod_vs_blocked_oracle(hf, cuda_device, cfg2_case, entry):
    from hippyflow_b200 import _lib as K
    c = cfg2_case
    proj = hf.PODProjectorFromData(None, M_output=c["M"], device=cuda_device)
    data = c["u"] if entry == "host_pipelined" else K.to_padded(c["u"], cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(data, c["rank"], shifted=True, method="randomized", Omega=c["Om"])
    assert proj.shift_route.startswith("pipelined" if entry == "host_pipelined" else "implicit")
    assert proj.info["passes"] >= 1
    k = check_eigs(d, c["d0"])
    assert k >= 200                                                           # the comparison covers (nearly) the whole basis
    assert subspace_angle(phi[:, :k], c["U0"][:, :k], c["M"]) < ANGLE_TOL
    np.testing.assert_allclose(shift, c["s0"], rtol=1e-13, atol=1e-15)
    assert np.abs(phi.T @ Mphi - np.eye(c["rank"])).max() < 1e-10            # test_PODProjector.py:161-168
    assert np.linalg.norm(c["M"] @ phi - Mphi) / np.linalg.norm(Mphi) < 1e-12  # :171-174
    # projected training data (M phi)^T (u_i - shift): north star (c), relative Frobenius 1e-12 against the oracle's
    # product with the SAME encoder
    red = hf.project_data(c["u"][:256] - shift, Mphi, cuda_device).cpu().numpy()
    ref = P.project_data(c["u"][:256] - shift, Mphi)
    assert np.linalg.norm(red - ref) / np.linalg.norm(ref) < PROJ_RTOL


def test_cfg2_shape_hep_vs_oracle(hf, cuda_device, cfg2_case):
    """The reference's own method (method of snapshots, PODProjector.py:812-833) at cfg2's dof count."""
    c = cfg2_case
    u = c["u"][:512]
    proj = hf.PODProjectorFromData(None, M_output=c["M"], device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 64, shifted=True, method="hep")
    d0, phi0, Mphi0, s0 = P.pod_from_data(u.copy(), c["M"], 64, shifted=True, method="hep")
    k = check_eigs(d, d0)
    assert subspace_angle(phi[:, :k], phi0[:, :k], c["M"]) < ANGLE_TOL
    np.testing.assert_allclose(shift, s0, rtol=1e-13, atol=1e-15)
    nrm = hf.weighted_l2_norm_vector(hf._lib.to_padded(phi0, cuda_device), proj.M_device).cpu().numpy()
    np.testing.assert_allclose(nrm, P.weighted_l2_norm_vector(phi0, c["M"]), rtol=1e-12)   # PODProjector.py:658-661


def test_benchmark_width_ill_conditioned_sketch_takes_the_shifted_route(hf, cuda_device):
    """m = 266 columns of a sketch whose condition number is ~1e15 (fast spectral decay): the shifted first Cholesky
    pass (info['shifted'] > 0) and the extra clean-up passes at benchmark width.  Checked against the column-by-column
    port (hIPPYlib's MGS with re-orthogonalisation keeps every direction of the sketch, like the shifted Cholesky-QR; the
    blocked oracle drops directions below sqrt(eps) and is not equivalent in this regime)."""
    n, N, rank = 16641, 512, 256
    M = syn.p1_mass_matrix_for(n)
    u = syn.snapshots(n, N, r0=400, decay=2.5, eps=1e-10, seed=11)
    Om = syn.gaussian_omega(n, rank + 10, seed=12)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, _ = proj.construct_subspace(u, rank, shifted=False, method="randomized", Omega=Om)
    assert proj.info["shifted"] > 0 and proj.info["passes"] >= 2, proj.info
    d0, U0, _, _ = P.pod_randomized_weighted(u, M, rank, Om, shifted=False)
    k = check_eigs(d, d0)
    assert subspace_angle(phi[:, :k], U0[:, :k], M) < ANGLE_TOL
    assert np.all(np.isfinite(phi))
    assert np.abs(phi.T @ Mphi - np.eye(rank)).max() < 1e-10


# ------------------------------------------------------------------ cfg3: active subspace, dM = 65,536, dQ = 100, rank 200 + 10
@pytest.fixture(scope="module")
def cfg3_case():
    N, dQ, dM, rank, p = 32, 100, 65536, 200, 10
    J = syn.jacobians(N, dQ, dM, r0=64, decay=1.0, seed=21)
    Om = syn.gaussian_omega(dM, rank + p, seed=22)
    return dict(N=N, dQ=dQ, dM=dM, rank=rank, J=J, Om=Om)


@pytest.mark.parametrize("preconditioned", [False, True])
def test_cfg3_shape_active_subspace_vs_blocked_oracle(hf, cuda_device, cfg3_case, preconditioned):
    c = cfg3_case
    R = syn.p1_mass_matrix_for(c["dM"]) if preconditioned else None          # 256^2 P1 mass matrix as the CSR prior
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = c["rank"], 10, False, False
    prior = hf.SparsePrior(R, device=cuda_device) if preconditioned else None
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(c["J"]), prior, collective=hf.NullCollective(), parameters=params,
                                      device=cuda_device)
    proj.Omega_GN = c["Om"]
    d, dec, enc = proj.construct_input_subspace(prior_preconditioned=preconditioned)
    d0, V0, E0 = P.as_input_from_jacobians_blocked(c["J"], c["rank"], c["Om"], B_csr=R)
    k = check_eigs(d, d0)
    V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
    assert subspace_angle(V[:, :k], V0[:, :k], R) < ANGLE_TOL
    if preconditioned:
        assert np.linalg.norm(R @ V - E) / np.linalg.norm(E) < 1e-12
        assert np.abs(V.T @ E - np.eye(c["rank"])).max() < 1e-10
    else:
        assert np.abs(V.T @ V - np.eye(c["rank"])).max() < 1e-10


# ------------------------------------------------------------------ cfg4: KLE + projection, n_m = 251,001, rank 128 + 10
def test_cfg4_shape_kle_and_projection_vs_blocked_oracle(hf, cuda_device):
    n, N, rank = 251001, 1024, 128
    M = syn.p1_mass_matrix_for(n)
    m_data = syn.snapshots(n, N, r0=256, decay=1.0, eps=1e-6, seed=31)
    Om = syn.gaussian_omega(n, rank + 10, seed=32)
    params = hf.KLEParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = rank, 10, False, False
    proj = hf.KLEProjector(hf.SampleCovariancePrior(m_data, M, device=cuda_device), parameters=params)
    d, dec, enc = proj.construct_input_subspace("mass", Omega=Om)
    d0, V0, E0, _ = P.pod_randomized_weighted_blocked(m_data, M, rank, Om, shifted=False)
    k = check_eigs(d, d0)
    V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
    assert subspace_angle(V[:, :k], V0[:, :k], M) < ANGLE_TOL
    assert np.linalg.norm(V.T @ E - np.eye(rank)) / np.sqrt(rank) < 1e-10      # test_KLEProjector.py:97-99
    assert np.linalg.norm(M @ V - E) / np.linalg.norm(E) < 1e-10                # :102-108
    # reduced inputs (M V)^T m_i for all stored draws
    red = hf.project_data(m_data, enc, cuda_device).cpu().numpy()
    ref = P.project_data(m_data, E)
    assert np.linalg.norm(red - ref) / np.linalg.norm(ref) < PROJ_RTOL
    # reduced Jacobians Phi^T J_i V and the two one-sided products at dQ = 200 (100 pointwise targets x 2 components)
    Nj, dQ, rQ = 4, 200, 128
    J = syn.jacobians(Nj, dQ, n, r0=48, seed=33)
    rng = np.random.default_rng(34)
    Phi = np.linalg.qr(rng.standard_normal((dQ, rQ)))[0]
    rj = hf.reduced_jacobians(J, Phi, V, cuda_device).cpu().numpy()
    rj0 = P.reduced_jacobians(J, Phi, V)
    assert np.linalg.norm(rj - rj0) / np.linalg.norm(rj0) < PROJ_RTOL
    jp = hf.jacobian_action(J, V, cuda_device).cpu().numpy()
    jp0 = P.j_psi(J, V)
    assert np.linalg.norm(jp - jp0) / np.linalg.norm(jp0) < PROJ_RTOL
    js = hf.jacobian_transpose_action(J, Phi, cuda_device).cpu().numpy()
    js0 = P.jstar_phi(J, Phi)
    assert np.linalg.norm(js - js0) / np.linalg.norm(js0) < PROJ_RTOL


# ------------------------------------------------------------------ cfg5: one shard of the scaling sweep, n = 1,002,001
def test_cfg5_shard_shape_weighted_pod_vs_blocked_oracle(hf, cuda_device):
    n, N, rank = 1002001, 512, 256
    M = syn.p1_mass_matrix_for(n)
    u = syn.snapshots(n, N, r0=512, decay=1.0, eps=1e-6, seed=41)
    Om = syn.gaussian_omega(n, rank + 10, seed=42)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(u, rank, shifted=True, method="randomized", Omega=Om)
    d0, U0, _, s0 = P.pod_randomized_weighted_blocked(u, M, rank, Om, shifted=True)
    k = check_eigs(d, d0)
    assert subspace_angle(phi[:, :k], U0[:, :k], M) < ANGLE_TOL
    np.testing.assert_allclose(shift, s0, rtol=1e-13, atol=1e-15)
    assert np.abs(phi.T @ Mphi - np.eye(rank)).max() < 1e-10
