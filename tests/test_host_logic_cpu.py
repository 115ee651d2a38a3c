"""CPU: the HOST logic of the projector classes (routes, order of products, mean-shift corrections, deferred clean-up
pass, file contract) run through ``cpu_device_shim`` -- a test double that evaluates the C-ABI wrappers with torch fp64
on the CPU -- against the oracle and the golden vectors.  The bodies are the parity tests of test_gpu_parity.py, so on a
B200 the very same assertions run against the CUDA kernels.  Nothing here says anything about the kernels themselves."""
import inspect

import pytest
import torch

import test_gpu_parity as G
from cpu_device_shim import emulated_device


def _run(fn, request, **params):
    import hippyflow_b200 as hf
    with emulated_device() as dev:
        kwargs = {}
        for name in inspect.signature(fn).parameters:
            if name == "hf":
                kwargs[name] = hf
            elif name == "cuda_device":
                kwargs[name] = dev
            elif name in params:
                kwargs[name] = params[name]
            else:
                kwargs[name] = request.getfixturevalue(name)
        fn(**kwargs)


@pytest.mark.parametrize("shifted", [True, False])
def test_pod_randomized_weighted(request, shifted):
    _run(G.test_pod_randomized_weighted_vs_oracle, request, shifted=shifted)


def test_pod_faithful_equals_shortcut(request):
    _run(G.test_pod_randomized_faithful_equals_shortcut, request)


@pytest.mark.parametrize("method", ["hep", "ghep", "inverse_ghep"])
def test_pod_deterministic_methods(request, method):
    _run(G.test_pod_hep_vs_reference_golden, request, shifted=True, method=method)


def test_pod_errors_and_doublepass(request):
    _run(G.test_pod_unavailable_method_and_rank_check, request)
    _run(G.test_pod_doublepass_vs_golden, request)


def test_meanjtj_operator(request):
    _run(G.test_meanjtj_operator_vs_reference_golden, request)


@pytest.mark.parametrize("preconditioned", [False, True])
def test_active_subspace(request, preconditioned):
    _run(G.test_active_subspace_vs_golden, request, preconditioned=preconditioned)


def test_active_subspace_noise(request):
    _run(G.test_active_subspace_noise_weighted_vs_oracle, request)


def test_kle_mass(request):
    _run(G.test_kle_mass_vs_golden, request)


def test_data_contract(request, tmp_path):
    _run(G.test_data_contract_roundtrip, request, tmp_path=tmp_path)


def test_edge_cases_and_files(request, tmp_path):
    _run(G.test_rank_deficient_snapshots_and_extreme_ranks, request)
    _run(G.test_projector_files_written_with_reference_names, request, tmp_path=tmp_path)
    _run(G.test_batched_list_and_stacked_operator_give_equal_eigenvalues, request)


def test_error_sweeps(request):
    _run(G.test_projection_error_sweep_vs_numpy, request)


@pytest.mark.parametrize("mean_scale", [0.0, 30.0])
@pytest.mark.parametrize("route", ["pipelined", "resident_implicit", "resident_explicit", "host_unpipelined"])
def test_mean_shift_routes(request, route, mean_scale):
    _run(G.test_pod_randomized_mean_shift_routes, request, route=route, mean_scale=mean_scale)


def test_dominant_mean_falls_back(request):
    _run(G.test_pod_randomized_dominant_mean_falls_back_to_explicit_shift, request)


def test_shim_decodes_the_cluster_plans():
    """The two packed forms of the cluster plan (panel records / fragment records) describe the same matrix."""
    import numpy as np
    from hippyflow_b200 import _lib as K, synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    with emulated_device() as dev:
        M = syn.p1_mass_matrix(70, 63)
        Md = CsrMatrix(M, dev)
        assert Md.plan is not None
        B = K.to_padded(np.random.default_rng(0).standard_normal((M.shape[0], 138)), dev)
        ref = M @ B.numpy()
        for impl in ("dmma", "frag"):
            Md.impl = impl
            np.testing.assert_allclose(Md.matmat(B).numpy(), ref, rtol=1e-13, atol=1e-16)


def test_cluster_plan_adapts_to_denser_rows_and_kernel_choice():
    """Host side of the SpMM dispatch: a 7-point mesh matrix gets (16, 32) clusters, a 21-point one the (16, 48) budget
    (so that clusters keep ~9 rows instead of ~2); 'auto' picks the run-staged kernel for 65..384 columns, the fragment kernel
    for other blocks of >= 32 columns, the generic CSR kernel below, and falls back from the run-staged kernel when the block's
    pitch is too wide for its ring; duplicates in the input matrix are summed first."""
    import numpy as np
    import scipy.sparse as sp
    import torch
    from hippyflow_b200 import _lib as K, synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    nx = 70
    idx = np.arange(nx * nx).reshape(nx, nx)
    r, c = [], []
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            if abs(dx) + abs(dy) <= 3:
                a = idx[max(0, -dx):nx - max(0, dx), max(0, -dy):nx - max(0, dy)]
                b = idx[max(0, dx):nx - max(0, -dx), max(0, dy):nx - max(0, -dy)]
                r.append(a.ravel()); c.append(b.ravel())
    r, c = np.concatenate(r), np.concatenate(c)
    dense = sp.csr_matrix((np.random.default_rng(1).standard_normal(r.size), (r, c)), shape=(nx * nx, nx * nx))
    with emulated_device() as dev:
        Md = CsrMatrix(syn.p1_mass_matrix(70, 63), dev)
        assert Md.impl == "auto" and Md.plan["max_rows"] <= 16 and Md.plan["max_cols_cap"] <= 32
        Dd = CsrMatrix(dense, dev)
        assert Dd.plan is not None and 32 < Dd.plan["max_cols_cap"] <= 48
        assert Dd.shape[0] / Dd.plan["nclusters"] > 6                      # (16, 32) would leave ~2.4 rows per cluster
        rng = np.random.default_rng(0)
        for m, kernel in ((266, "csr_spmm_runs_kernel"), (500, "csr_spmm_dmma_%s_kernel" % CsrMatrix.WIDE_DEFAULT), (138, "csr_spmm_runs_kernel"),
                          (74, "csr_spmm_runs_kernel"), (40, "csr_spmm_dmma_frag_kernel"), (20, "csr_spmm_panel_kernel")):
            B = K.to_padded(rng.standard_normal((dense.shape[0], m)), dev)
            out, used = Dd._matmat(B, None)
            assert used == kernel
            np.testing.assert_allclose(out.numpy(), dense @ B.numpy(), rtol=1e-12, atol=1e-13)
        wide = torch.zeros((dense.shape[0], 4096), dtype=torch.float64)      # a 138-column view of a 4096-column block
        wide[:, :138] = torch.from_numpy(rng.standard_normal((dense.shape[0], 138)))
        out, used = Dd._matmat(wide[:, :138], None)
        assert used == "csr_spmm_dmma_frag_kernel"                           # no two ring slots at this pitch
        np.testing.assert_allclose(out.numpy(), dense @ wide[:, :138].numpy(), rtol=1e-12, atol=1e-13)
        # duplicate entries (non-canonical CSR) are summed before the plan is built
        raw = sp.csr_matrix(dense.shape)                                   # row 0 carries its first 7 entries twice
        raw.data, raw.indices, raw.indptr = np.r_[dense.data[:7], dense.data], np.r_[dense.indices[:7], dense.indices], \
            np.r_[0, dense.indptr[1:] + 7]
        assert not raw.has_canonical_format
        Rd = CsrMatrix(raw, dev)
        B = K.to_padded(rng.standard_normal((dense.shape[0], 266)), dev)
        np.testing.assert_allclose(Rd.matmat(B).numpy(), raw @ B.numpy(), rtol=1e-12, atol=1e-13)


def test_small_operator_classes(request):
    import test_gpu_operators_extra as E
    _run(E.test_jjt_and_dense_operator, request)
    _run(E.test_low_rank_rectangular_operator, request)


def test_small_operator_classes_vs_reference_golden():
    """LowRankRectangularOperator, PriorPreconditionedProjector and npToDolfinOperator of this package (host logic under the
    CPU test double) against outputs of the UNMODIFIED reference classes (tests/golden/small_operators_ref.npz, generated by
    oracle/make_golden_small_ops.py)."""
    import os
    import numpy as np
    import hippyflow_b200 as hf
    from hippyflow_b200 import synthetic as syn
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_operators_ref.npz"))
    # the golden vectors themselves obey the defining formulas
    A = g["lr_U"] @ np.diag(g["lr_s"]) @ g["lr_V"].T
    np.testing.assert_allclose(g["lr_mult"], g["lr_x"] @ A.T, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(g["lr_transpmult"], g["lr_w"] @ A, rtol=1e-13, atol=1e-14)
    with emulated_device() as dev:
        op = hf.LowRankRectangularOperator(hf.DeviceMultiVector.from_dense(g["lr_U"], dev), g["lr_s"],
                                           hf.DeviceMultiVector.from_dense(g["lr_V"], dev))
        x, y = hf.DeviceVector(121, dev), hf.DeviceVector(100, dev)
        for i in range(4):
            x.set_local(g["lr_x"][i])
            y.set_local(np.full(100, 3.0))
            op.mult(x, y)
            np.testing.assert_allclose(y.get_local(), g["lr_mult"][i], rtol=1e-12, atol=1e-13)
            y.set_local(g["lr_w"][i])
            op.transpmult(y, x)
            np.testing.assert_allclose(x.get_local(), g["lr_transpmult"][i], rtol=1e-12, atol=1e-13)
        X = hf.DeviceMultiVector.from_dense(g["lr_x"].T.copy(), dev)
        Y = hf.DeviceMultiVector(100, 4, device=dev)
        op.matMvMult(X, Y)
        np.testing.assert_allclose(Y.to_dense(), g["lr_mult"].T, rtol=1e-12, atol=1e-13)

        M = syn.p1_mass_matrix(int(g["pp_nx"]))
        proj = hf.PriorPreconditionedProjector(hf.DeviceMultiVector.from_dense(g["pp_U"], dev), hf.CsrMatrix(M, dev))
        n = M.shape[0]
        px, py = hf.DeviceVector(n, dev), hf.DeviceVector(n, dev)
        for i in range(4):
            px.set_local(g["pp_x"][i])
            py.set_local(np.ones(n))
            proj.mult(px, py)
            np.testing.assert_allclose(py.get_local(), g["pp_mult"][i], rtol=1e-12, atol=1e-13)

        dop = hf.npToDolfinOperator(g["np_A"], device=dev)
        u, v = hf.DeviceVector(1, dev), hf.DeviceVector(1, dev)
        dop.init_vector(u, 0)
        dop.init_vector(v, 1)
        v.set_local(g["np_x"])
        dop.mult(v, u)
        np.testing.assert_allclose(u.get_local(), g["np_mult"], rtol=1e-13, atol=1e-14)
        u.set_local(g["np_w"])
        dop.transpmult(u, v)
        np.testing.assert_allclose(v.get_local(), g["np_transpmult"], rtol=1e-13, atol=1e-14)


def test_pod_randomized_random_shapes_vs_oracle():
    """Property test of the host logic (hypothesis): for random mesh sizes, sample counts, ranks, oversampling, mean
    offsets, host / resident input and all mean-shift routes, the weighted randomized POD driven through the CPU test
    double reproduces the oracle's eigenvalues (rel 1e-10 on the leading ones), M-orthonormality, encoder = M decoder and
    the sample mean."""
    import numpy as np
    from hypothesis import given, settings, strategies as st, HealthCheck
    import hippyflow_b200 as hf
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P

    @settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(nx=st.integers(5, 12), ny=st.integers(5, 12), N=st.integers(24, 90), rank=st.integers(2, 12),
           over=st.integers(2, 10), mean=st.sampled_from([0.0, 0.5, 20.0]), shifted=st.booleans(),
           route=st.sampled_from(["host_pipelined", "host_unpipelined", "resident_implicit", "resident_explicit"]),
           seed=st.integers(0, 10_000))
    def run(nx, ny, N, rank, over, mean, shifted, route, seed):
        M = syn.p1_mass_matrix(nx, ny)
        n = M.shape[0]
        rank = min(rank, N - 1, n - 1)
        m = min(rank + over, n)
        u = syn.snapshots(n, N, r0=min(20, N), seed=seed) + mean * np.cos(np.linspace(0, 2, n))[None, :]
        Om = syn.gaussian_omega(n, m, seed=seed + 1)
        d0, U0, E0, s0 = P.pod_randomized_weighted(u, M, rank, Om, shifted=shifted)
        with emulated_device() as dev:
            proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
            kw = dict(shifted=shifted, method="randomized", Omega=Om, oversampling=m - rank)
            if route.startswith("host"):
                d, phi, Mphi, shift = proj.construct_subspace(u.copy(), rank, pipelined_upload=(route == "host_pipelined"), **kw)
            else:
                from hippyflow_b200 import _lib as K
                Xd = K.to_padded(u, dev)
                d, phi, Mphi, shift = proj.construct_subspace(Xd, rank, implicit_shift=(route == "resident_implicit"), **kw)
        lead = d0 / d0[0] > 1e-5
        np.testing.assert_allclose(d[lead], d0[lead], rtol=1e-10)
        np.testing.assert_allclose(d[~lead], d0[~lead], rtol=0, atol=1e-12 * d0[0])
        np.testing.assert_allclose(shift, s0, rtol=0, atol=1e-13 * (1.0 + abs(mean)))
        G = phi.T @ Mphi
        assert np.abs(G - np.eye(rank)).max() < 1e-9
        np.testing.assert_allclose(Mphi, M @ phi, rtol=1e-11, atol=1e-14)
        k = int(lead.sum())
        assert P.principal_angle(phi[:, :k], U0[:, :k], M) < 1e-7

    run()


def test_active_subspace_and_kle_random_shapes_vs_oracle():
    """Property test of the host logic (hypothesis) for the other two projectors: input active subspace from stored
    Jacobians (plain / noise-weighted / prior-preconditioned with a CSR prior and the device block CG) and KLE from stored
    draws ('mass' / 'identity'), random shapes and ranks, CPU test double vs oracle."""
    import numpy as np
    from hypothesis import given, settings, strategies as st, HealthCheck
    import hippyflow_b200 as hf
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P

    def check(d, d0, V, V0, Mw):
        lead = d0 / d0[0] > 1e-5
        np.testing.assert_allclose(d[lead], d0[lead], rtol=1e-9)
        np.testing.assert_allclose(d, d0, rtol=0, atol=1e-10 * d0[0])
        k = int(lead.sum())
        assert P.principal_angle(V[:, :k], V0[:, :k], Mw) < 1e-7

    @settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(nx=st.integers(4, 9), N=st.integers(4, 20), dQ=st.integers(3, 24), rank=st.integers(2, 10),
           over=st.integers(2, 8), mode=st.sampled_from(["plain", "noise", "prior"]), seed=st.integers(0, 10_000))
    def run_as(nx, N, dQ, rank, over, mode, seed):
        Mp = syn.p1_mass_matrix(nx)
        dM = Mp.shape[0]
        rank = min(rank, dM - 1)
        m = min(rank + over, dM)
        J = syn.jacobians(N, dQ, dM, r0=min(16, dQ), seed=seed)
        Om = syn.gaussian_omega(dM, m, seed=seed + 1)
        G = None
        if mode == "noise":
            A = np.random.default_rng(seed).standard_normal((dQ, dQ))
            G = A @ A.T / dQ + np.eye(dQ)
        d0, V0, E0 = P.as_input_from_jacobians(J, rank, Om, noise_cov_inv=G, B_csr=Mp if mode == "prior" else None)
        params = hf.ActiveSubspaceParameterList()
        params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = rank, m - rank, False, False
        with emulated_device() as dev:
            prior = hf.SparsePrior(Mp, device=dev) if mode == "prior" else None
            proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J, G), prior, parameters=params, device=dev)
            proj.Omega_GN = Om
            d, dec, enc = proj.construct_input_subspace(prior_preconditioned=(mode == "prior"))
            V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
        check(d, d0, V, V0, Mp if mode == "prior" else None)
        if mode == "prior":
            np.testing.assert_allclose(E, Mp @ V, rtol=1e-11, atol=1e-14)
        else:
            np.testing.assert_array_equal(E, V)

    @settings(max_examples=10, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(nx=st.integers(5, 10), N=st.integers(30, 80), rank=st.integers(2, 12), over=st.integers(2, 8),
           orth=st.sampled_from(["mass", "identity"]), seed=st.integers(0, 10_000))
    def run_kle(nx, N, rank, over, orth, seed):
        M = syn.p1_mass_matrix(nx)
        n = M.shape[0]
        m = min(rank + over, n)
        m_data = syn.snapshots(n, N, r0=min(24, N), decay=1.5, seed=seed)
        Om = syn.gaussian_omega(n, m, seed=seed + 1)
        d0, V0, E0 = P.kle_from_samples(m_data, M, rank, Om, orth)
        params = hf.KLEParameterList()
        params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = rank, m - rank, False, False
        with emulated_device() as dev:
            proj = hf.KLEProjector(hf.SampleCovariancePrior(m_data, M, device=dev), parameters=params)
            d, dec, enc = proj.construct_input_subspace(orth, Omega=Om)
            V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
        check(d, d0, V, V0, M if orth == "mass" else None)
        np.testing.assert_allclose(E, (M @ V) if orth == "mass" else V, rtol=1e-11, atol=1e-14)

    run_as()
    run_kle()


def test_list_and_collective_operators_vs_reference_golden(request):
    import test_gpu_operators_extra as E
    _run(E.test_list_and_collective_operators_vs_reference_golden, request)
