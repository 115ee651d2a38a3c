"""GPU timeline of ONE resident eigensolve step (torch.profiler / CUPTI): every kernel and memcpy with start time and
duration, the idle gaps between them and the host runtime calls (synchronisations, copies) that sit inside each gap.
Writes gpurun_out/<tag>_timeline.json and prints a summary.  Development aid: numbers taken under the profiler are
not bench values; the SHARES and the gap list are what matters.

    python tools/timeline_step.py [workload] [tag]
    python -m torch.distributed.run --nproc-per-node N ... tools/timeline_step.py [workload] [tag]   (rank 0 profiles)
"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import hippyflow_b200 as hf
from hippyflow_b200 import synthetic as syn

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
coll = hf.NullCollective()
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    coll = hf.TorchCollective()
n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
M = syn.p1_mass_matrix_for(n)
proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7, row_offset=rank * n_loc)


def step():
    return proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, return_device=True,
                                   overwrite_data=True, collective=coll)


for _ in range(4):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
print("step wall (no profiler): %.2f ms" % wall)

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
ev = prof.events()
gpu = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev if e.device_type == torch.autograd.DeviceType.CUDA],
             key=lambda x: x[0])
cpu = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev if e.device_type == torch.autograd.DeviceType.CPU
              and (e.name.startswith("cuda") or e.name.startswith("cu"))], key=lambda x: x[0])
if not gpu:
    print("no CUDA activity records (CUPTI unavailable?)")
    sys.exit(0)
T0, T1 = gpu[0][0], max(g[1] for g in gpu)
busy = 0.0
gaps = []
cur_end = gpu[0][0]
for s, e, name in gpu:
    if s > cur_end:
        inside = [c[2] for c in cpu if c[0] < s and c[1] > cur_end]
        gaps.append({"at_ms": (cur_end - T0) / 1e3, "gap_us": (s - cur_end), "before": name[:60], "host_calls": sorted(set(inside))[:6]})
    busy += max(0.0, e - max(s, cur_end))
    cur_end = max(cur_end, e)
span = (T1 - T0) / 1e3
print("GPU span %.2f ms, busy %.2f ms, idle %.2f ms over %d kernels/copies" % (span, busy / 1e3, span - busy / 1e3, len(gpu)))
gaps.sort(key=lambda g: -g["gap_us"])
for g in gaps[:25]:
    print("  gap %8.1f us at %7.2f ms before %-60s host: %s" % (g["gap_us"], g["at_ms"], g["before"], ", ".join(g["host_calls"])))
agg = {}
for s, e, name in gpu:
    a = agg.setdefault(name[:70], [0, 0.0])
    a[0] += 1
    a[1] += (e - s) / 1e3
print("kernels by total time:")
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print("  %8.3f ms  x%-3d %s" % (t, c, name))
json.dump({"wall_ms_no_profiler": wall, "gpu_span_ms": span, "gpu_busy_ms": busy / 1e3, "gaps": gaps[:60],
           "kernels": [{"start_ms": (s - T0) / 1e3, "dur_us": e - s, "name": nm[:90]} for s, e, nm in gpu]},
          open("gpurun_out/%s_timeline.json" % tag, "w"), indent=0)
if world > 1:
    dist.destroy_process_group()
