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


def test_active_subspace_noise_and_output(request):
    _run(G.test_active_subspace_noise_weighted_vs_oracle, request)
    _run(G.test_active_output_subspace_vs_oracle, request)


def test_kle_mass(request):
    _run(G.test_kle_mass_vs_golden, request)


def test_projection_and_data_contract(request, tmp_path):
    _run(G.test_projection_of_stored_data_vs_oracle, request)
    _run(G.test_data_contract_roundtrip, request, tmp_path=tmp_path)


def test_edge_cases_and_files(request, tmp_path):
    _run(G.test_rank_deficient_snapshots_and_extreme_ranks, request)
    _run(G.test_projector_files_written_with_reference_names, request, tmp_path=tmp_path)
    _run(G.test_batched_list_and_stacked_operator_give_equal_eigenvalues, request)


def test_error_sweeps_and_jacobian_svd(request):
    _run(G.test_projection_error_sweep_vs_numpy, request)
    _run(G.test_jacobian_truncated_svd_vs_numpy, request)


@pytest.mark.parametrize("mean_scale", [0.0, 30.0])
@pytest.mark.parametrize("route", ["pipelined", "resident_implicit", "resident_explicit", "host_unpipelined"])
def test_mean_shift_routes(request, route, mean_scale):
    _run(G.test_pod_randomized_mean_shift_routes, request, route=route, mean_scale=mean_scale)


def test_dominant_mean_falls_back(request):
    _run(G.test_pod_randomized_dominant_mean_falls_back_to_explicit_shift, request)


def test_shim_decodes_the_cluster_plans():
    """The two cluster plans (cp.async panels / TMA blobs) built by CsrMatrix describe the same matrix."""
    import numpy as np
    from hippyflow_b200 import _lib as K, synthetic as syn
    from hippyflow_b200.linalg import CsrMatrix
    with emulated_device() as dev:
        M = syn.p1_mass_matrix(70, 63)
        Md = CsrMatrix(M, dev)
        assert Md.plan is not None
        B = K.to_padded(np.random.default_rng(0).standard_normal((M.shape[0], 138)), dev)
        ref = M @ B.numpy()
        for impl in ("tma", "staged", "regblock", "dmma", "frag"):
            Md.impl = impl
            np.testing.assert_allclose(Md.matmat(B).numpy(), ref, rtol=1e-13, atol=1e-16)
