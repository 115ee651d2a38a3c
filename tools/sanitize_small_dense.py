"""Small run of the round-2 kernels for compute-sanitizer (memcheck / racecheck): device Cholesky-QR factor, batched Jacobi
SVD, strided-batch GEMM (independent + reduce), batched randomized SVD."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for m in (5, 37, 70):
    Y = rng.standard_normal((m + 9, m))
    G = Y.T @ Y
    S, st = K.chol_inverse(K.to_padded(G, dev))
    S = S.cpu().numpy()
    assert st.cpu().numpy()[0] == 0 and np.abs(S.T @ G @ S - np.eye(m)).max() < 1e-9
A = rng.standard_normal((3, 40, 21))
Ad = K.batched_empty(3, 40, 21, dev); Ad.copy_(torch.as_tensor(A))
sig, info = K.jacobi_svd_batched_(Ad)
assert np.allclose(sig.cpu().numpy()[0], np.linalg.svd(A[0], compute_uv=False), rtol=1e-10)
J = syn.jacobians(5, 20, 150, r0=12, seed=2)
U, s, V = hf.jacobian_truncated_svd(J, 6, dev, oversampling=4)
op = hf.MeanJJTfromDataOperator(J, device=dev)
C = op.dense().cpu().numpy()
assert np.abs(C - np.einsum("iqm,irm->qr", J, J) / 5).max() < 1e-12
out = hf.jacobian_transpose_action(J, rng.standard_normal((20, 7)), dev)
M = syn.p1_mass_matrix(20)
u = syn.snapshots(M.shape[0], 60, r0=40, seed=1)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 12, shifted=True, method="randomized")
assert proj.info.get("route") == "device" and np.abs(phi.T @ Mphi - np.eye(12)).max() < 1e-9
torch.cuda.synchronize(); print("SANITIZE_RUN_OK")
