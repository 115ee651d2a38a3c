"""Host-side profile (cProfile) and phase timing of one resident eigensolve step at a bench workload."""
import cProfile, pstats, sys, time, io
sys.path.insert(0, ".")
import torch
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
M = syn.p1_mass_matrix_for(n)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7)
def step():
    return proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, return_device=True, overwrite_data=True)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("step wall ms", (time.perf_counter() - t0) * 1e3)
if "--ncu" in sys.argv:
    sys.exit(0)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
print(proj.info, proj.timings)
