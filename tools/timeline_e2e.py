"""GPU timeline of ONE end-to-end solve (pinned host snapshots -> NumPy results, cfg2): when the last snapshot chunk lands,
what runs after it (the tail that PCIe cannot hide) and when the result download starts.  torch.profiler / CUPTI; the numbers
are under a profiler -- shares and order matter, not absolutes.

    python tools/timeline_e2e.py
"""
import sys
import time

sys.path.insert(0, ".")
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import hippyflow_b200 as hf
from hippyflow_b200 import synthetic as syn

wl = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda:0")
n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
M = syn.p1_mass_matrix_for(n)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7)
host = torch.empty((n_loc, n), dtype=torch.float64, pin_memory=True)
host.copy_(Xt)
del Xt
torch.cuda.synchronize()


def step():
    return proj.construct_subspace(host, k, shifted=True, method="randomized", oversampling=p)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
print("e2e step wall (no profiler): %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
gpu = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
             key=lambda x: x[0])
T0 = gpu[0][0]
h2d = [g for g in gpu if "HtoD" in g[2] and g[1] - g[0] > 500]
d2h = [g for g in gpu if "DtoH" in g[2] and g[1] - g[0] > 500]
last_h2d = max(g[1] for g in h2d)
print("big H2D copies: %d, first starts %.2f ms, last ends %.2f ms" % (len(h2d), (h2d[0][0] - T0) / 1e3, (last_h2d - T0) / 1e3))
for g in h2d:
    print("   H2D %8.2f -> %8.2f ms (%.2f ms)" % ((g[0] - T0) / 1e3, (g[1] - T0) / 1e3, (g[1] - g[0]) / 1e3))
print("big D2H copies:")
for g in d2h:
    print("   D2H %8.2f -> %8.2f ms (%.2f ms)" % ((g[0] - T0) / 1e3, (g[1] - T0) / 1e3, (g[1] - g[0]) / 1e3))
print("end of GPU activity %.2f ms" % ((max(g[1] for g in gpu) - T0) / 1e3))
print("kernels that START after the last H2D byte landed (start ms after it, duration us):")
busy = 0.0
for s, e, name in gpu:
    if s >= last_h2d and "Memcpy" not in name:
        busy += (e - s)
        if e - s > 50:
            print("   +%7.2f  %9.1f  %s" % ((s - last_h2d) / 1e3, e - s, name[:80]))
print("kernel time after the last H2D: %.2f ms" % (busy / 1e3))
print("kernels running DURING the upload (sum): %.2f ms" % (sum(e - s for s, e, nm in gpu if s < last_h2d and "Memcpy" not in nm) / 1e3))
