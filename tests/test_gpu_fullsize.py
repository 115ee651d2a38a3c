"""GPU: size-independent properties at BASELINE.json's full single-GPU size (configs[1]: 4096 snapshots x 263,169
dofs, rank 256 + 10, mass-matrix weighted), where the CPU oracle cannot run in seconds."""
import numpy as np
import pytest
import torch

from hippyflow_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

N_DOF, N_SNAP, RANK, OVER = 263169, 4096, 256, 10


@pytest.fixture(scope="module")
def solved(cuda_device):
    import hippyflow_b200 as hf
    from hippyflow_b200 import _lib as K
    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 40e9:
        pytest.skip("needs ~40 GB of free device memory")
    M = syn.p1_mass_matrix_for(N_DOF)
    Xt = syn.snapshots_device(N_DOF, N_SNAP, cuda_device, r0=512, seed=7)
    proj = hf.PODProjectorFromData(None, M_output=M, device=cuda_device)
    d, phi, Mphi, shift = proj.construct_subspace(Xt, RANK, shifted=True, method="randomized", oversampling=OVER,
                                                  return_device=True, overwrite_data=True)
    return dict(hf=hf, K=K, M=M, Xt=Xt, proj=proj, d=d, phi=phi, Mphi=Mphi, shift=shift)


def test_fullsize_basis_is_M_orthonormal_and_encoder_consistent(solved):
    K, phi, Mphi, M = solved["K"], solved["phi"], solved["Mphi"], solved["M"]
    G = K.dgemm(K.HFB_TN, phi, Mphi).cpu().numpy()
    assert np.abs(G - np.eye(RANK)).max() < 1e-10                          # test_PODProjector.py:161-168 at full size
    cols = [0, 17, 255]
    ph = phi[:, cols].cpu().numpy()
    np.testing.assert_allclose(Mphi[:, cols].cpu().numpy(), M @ ph, rtol=1e-12, atol=1e-18)   # encoder = M phi (:171-174)
    d = solved["d"]
    assert np.all(np.diff(d) <= 0) and d[-1] > 0


def test_fullsize_eigen_relation(solved):
    """C M phi_i = d_i phi_i for the leading modes, evaluated with independent products (test_PODProjector.py:188-208)."""
    K, Xt, phi, Mphi, d = solved["K"], solved["Xt"], solved["phi"], solved["Mphi"], solved["d"]
    lead = 16
    W = K.dgemm(K.HFB_NN, Xt, Mphi[:, :lead].contiguous() if False else K.to_padded(Mphi[:, :lead], Xt.device))
    CMphi = K.dgemm(K.HFB_TN, Xt, W, alpha=1.0 / N_SNAP)                    # X X^T M phi / N (data already shifted)
    R = CMphi - phi[:, :lead] * torch.as_tensor(d[:lead], device=Xt.device)
    num = torch.linalg.norm(R, dim=0)
    den = torch.linalg.norm(phi[:, :lead] * torch.as_tensor(d[:lead], device=Xt.device), dim=0)
    assert float((num / den).max()) < 1e-2


def test_fullsize_projection_is_idempotent_and_matches_torch(solved):
    K, hf, Xt, phi, Mphi = solved["K"], solved["hf"], solved["Xt"], solved["phi"], solved["Mphi"]
    red = hf.project_data(Xt, Mphi, Xt.device)                              # (N, r) = (M phi)^T u_i
    rows = torch.tensor([0, 1000, 4095], device=Xt.device)
    ref = Xt[rows] @ Mphi                                                   # torch fp64 reference on a row subset
    assert float((red[rows] - ref).norm() / ref.norm()) < 1e-12
    # projecting the reconstruction again changes nothing: (phi red^T) projected = red
    rec = K.dgemm(K.HFB_NT, K.to_padded(red[:64], Xt.device), phi)          # (64, n)
    red2 = hf.project_data(rec, Mphi, Xt.device)
    assert float((red2 - red[:64]).norm() / red[:64].norm()) < 1e-10


def test_fullsize_bitwise_reproducible_and_linear(solved):
    K, Xt, Mphi = solved["K"], solved["Xt"], solved["Mphi"]
    B = K.padded_empty(N_DOF, RANK + OVER, Xt.device)
    K.fill_random_(B, 5)
    W1 = K.dgemm(K.HFB_NN, Xt, B).clone()
    W2 = K.dgemm(K.HFB_NN, Xt, B).clone()
    assert torch.equal(W1, W2)                                              # deterministic split-K
    B2 = K.padded_empty(N_DOF, RANK + OVER, Xt.device)
    B2.copy_(2.0 * B)
    W3 = K.dgemm(K.HFB_NN, Xt, B2)
    assert torch.equal(W3, 2.0 * W1)                                        # scaling by 2 is exact in binary fp
