"""Breakdown of the end-to-end (host -> NumPy) eigensolve at cfg2."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
import bench
wl = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda:0")
n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
M = syn.p1_mass_matrix_for(n)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7)
host = torch.empty((n_loc, n), dtype=torch.float64, pin_memory=True); host.copy_(Xt); torch.cuda.synchronize()
def T(f, name, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
    print(f"{name}: {t*1e3:.1f} ms", flush=True); return r
dst = K.padded_empty(n_loc, n, dev)
T(lambda: dst.copy_(host, non_blocking=True), "H2D pinned->padded device (8.6 GB)")
dst2 = torch.empty((n_loc, n), dtype=torch.float64, device=dev)
T(lambda: dst2.copy_(host, non_blocking=True), "H2D pinned->contiguous device")
del dst2
T(lambda: K.to_padded(host, dev), "to_padded(host)")
phi = K.padded_empty(n, k, dev).normal_()
T(lambda: sys.modules["hippyflow_b200.modeling.PODProjector"].to_host(phi), "to_host (n x 256)")
T(lambda: phi.cpu(), "phi.cpu() pageable")
T(lambda: proj.construct_subspace(host, k, shifted=True, method="randomized", oversampling=p), "e2e step")
print(proj.timings)
T(lambda: proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, return_device=True, overwrite_data=True), "resident step")
