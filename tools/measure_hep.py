"""Timing of the deterministic method='hep' (method of snapshots, PODProjector.py:812-833) at cfg2 on one GPU."""
import sys, time
sys.path.insert(0, ".")
import torch
import hippyflow_b200 as hf
from hippyflow_b200 import synthetic as syn
dev = torch.device("cuda:0")
n, N, k = 263169, 4096, 256
M = syn.p1_mass_matrix_for(n)
Xt = syn.snapshots_device(n, N, dev, r0=512, seed=7)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, phi, Mphi, sh = proj.construct_subspace(Xt, k, shifted=True, method="hep", return_device=True, overwrite_data=True)
    torch.cuda.synchronize(); print("hep step", (time.perf_counter() - t0) * 1e3, "ms", d[:3])
torch.cuda.synchronize(); t0 = time.perf_counter()
d2, phi2, _, _ = proj.construct_subspace(Xt, k, shifted=True, method="randomized", return_device=True, overwrite_data=True)
torch.cuda.synchronize(); print("randomized step", (time.perf_counter() - t0) * 1e3, "ms", d2[:3])
import numpy as np
print("rel diff of leading 8 eigenvalues hep vs randomized:", np.abs(d[:8] - d2[:8]) / d[:8])
from hippyflow_b200 import _lib as K
t0 = time.perf_counter(); Zt = proj.M_device.matmat_rows(Xt); torch.cuda.synchronize(); print("spmm_rows", (time.perf_counter()-t0)*1e3)
t0 = time.perf_counter(); G = K.dgemm(K.HFB_NT, Xt, Zt); torch.cuda.synchronize(); print("gram NT", (time.perf_counter()-t0)*1e3, 2.0*N*N*n/((time.perf_counter()-t0))*1e-12, "TF/s")
t0 = time.perf_counter(); s, U = torch.linalg.eigh(0.5*(G+G.t()).contiguous()); torch.cuda.synchronize(); print("eigh 4096 (torch/cusolver)", (time.perf_counter()-t0)*1e3)
