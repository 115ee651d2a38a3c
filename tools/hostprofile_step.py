"""Host-side profile of the resident eigensolve step (cProfile over 5 steps) plus wall-clock stamps around the first calls of
a step: finds host work that leaves the GPU idle (the step synchronises at its end, so the host cannot run ahead).

    python tools/hostprofile_step.py [workload]
"""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import torch

import bench
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K
from hippyflow_b200 import synthetic as syn

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
M = syn.p1_mass_matrix_for(n)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7)


def step():
    return proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, return_device=True,
                                   overwrite_data=True)


for _ in range(4):
    step()
torch.cuda.synchronize()

# wall-clock duration of every C-ABI call of one step (a blocking launch shows up here)
L = K.lib()
stamps = []
orig = {}
for name in K.EXPORTED:
    fn = getattr(L, name)
    orig[name] = fn

    def wrap(fn=fn, name=name):
        def call(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            stamps.append((name, t0, time.perf_counter()))
            return r
        return call
    setattr(L, name, wrap())
t_begin = time.perf_counter()
step()
torch.cuda.synchronize()
t_end = time.perf_counter()
for name, fn in orig.items():
    setattr(L, name, fn)
print("one step: %.2f ms wall; C-ABI calls in order (start ms, duration us):" % ((t_end - t_begin) * 1e3))
for name, t0, t1 in stamps:
    print("  %8.3f  %8.1f  %s" % ((t0 - t_begin) * 1e3, (t1 - t0) * 1e6, name))

pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
