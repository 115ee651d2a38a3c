"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family once, ragged sizes."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
dev = torch.device("cuda:0")
for layout in (0, 1, 2):
    for (M, N, Kd) in ((130, 138, 77), (257, 25, 50), (64, 266, 333)):
        A = torch.randn(M, Kd, dtype=torch.float64, device=dev); B = torch.randn(Kd, N, dtype=torch.float64, device=dev)
        Ad = K.to_padded(A.t().contiguous() if layout == 1 else A, dev); Bd = K.to_padded(B.t().contiguous() if layout == 2 else B, dev)
        for sp in (1, 3):
            C = K.dgemm(layout, Ad, Bd, splits=sp)
            assert float((C - A @ B).norm() / (A @ B).norm()) < 1e-13
M = syn.p1_mass_matrix(70)          # 5041 dofs -> staged SpMM plan
u = syn.snapshots(M.shape[0], 40, r0=20, seed=1)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
for method in ("randomized", "hep"):
    d, phi, Mphi, shift = proj.construct_subspace(u.copy(), 12, shifted=True, method=method)
    assert np.abs(phi.T @ Mphi - np.eye(12)).max() < 1e-8
J = syn.jacobians(6, 20, 150, r0=8, seed=2)
pa = hf.ActiveSubspaceParameterList(); pa["rank"], pa["oversampling"], pa["verbose"], pa["save_and_plot"] = 8, 4, False, False
asp = hf.ActiveSubspaceProjector(hf.StoredJacobians(J, np.eye(20) * 2.0), hf.SparsePrior(syn.p1_mass_matrix(14, 9), device=dev), parameters=pa, device=dev)
asp.construct_input_subspace(prior_preconditioned=True); asp.construct_output_subspace()
hf.reduced_jacobians(J, np.random.randn(20, 5), np.random.randn(150, 4), dev)
torch.cuda.synchronize(); print("SANITIZE_RUN_OK")
