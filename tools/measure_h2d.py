"""Concurrent pinned host -> device bandwidth at 1 / 2 / 4 / ... ranks of ONE box, no compute: the host-side floor of the
end-to-end leg of bench.py (the e2e step does not weak-scale because all ranks upload 8.6 GB at once).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/measure_h2d.py [bytes_per_rank] [tag]

For k in 1, 2, 4, ..., world the ranks < k copy `bytes_per_rank` from pinned host memory to their GPU at the same time
(the others wait at the barrier); rank 0 prints / writes {k: {per-rank GB/s (min, mean), aggregate GB/s}} plus the NUMA
view NVML gives for every GPU.  Also times the same copy from PAGEABLE memory through the pinned staging ring."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
nbytes = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4096 * 263184 * 8
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
bind = os.environ.get("HFB_BIND_NUMA", "1") != "0"
cpus = None
if bind:
    from hippyflow_b200.utilities import bind_to_gpu_numa_node
    cpus = bind_to_gpu_numa_node(local)
n = nbytes // 8
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
host.fill_(1.0)
devbuf = torch.empty(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()


def barrier():
    if world > 1:
        dist.barrier()


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o) for o in out]
    return [float(t)]


res = {"bytes_per_rank": nbytes, "world": world, "bound_cpus_rank0": (len(cpus) if cpus else None)}
k = 1
while k <= world:
    best = 1e30
    for rep in range(3):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank < k:
            devbuf.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt) if rank < k else best
    rates = gather(nbytes / best * 1e-9 if rank < k else 0.0)
    act = rates[:k]
    res["h2d_pinned_%d_ranks" % k] = {"per_rank_GBps_min": min(act), "per_rank_GBps_mean": sum(act) / k, "aggregate_GBps": sum(act)}
    # device -> pinned host (the result copies of the e2e leg are 1/8 of the upload)
    best = 1e30
    for rep in range(2):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank < k:
            host[: n // 8].copy_(devbuf[: n // 8], non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt) if rank < k else best
    rates = gather(nbytes / 8 / best * 1e-9 if rank < k else 0.0)
    res["d2h_pinned_%d_ranks" % k] = {"per_rank_GBps_min": min(rates[:k]), "aggregate_GBps": sum(rates[:k])}
    k *= 2

# pageable source through the pinned staging ring (what construct_subspace does for a plain NumPy array), all ranks
pg = torch.empty(n // 4, dtype=torch.float64)
pg.fill_(2.0)
ring = [(torch.empty(12 << 20, dtype=torch.float64, pin_memory=True), torch.cuda.Event()) for _ in range(3)]
for _, ev in ring:
    ev.record()
copy_stream = torch.cuda.Stream()
best = 1e30
for rep in range(2):
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    j = 0
    step = ring[0][0].numel()
    for s0 in range(0, pg.numel(), step):
        s1 = min(pg.numel(), s0 + step)
        buf, ev = ring[j % 3]
        ev.synchronize()
        buf[: s1 - s0].copy_(pg[s0:s1])
        with torch.cuda.stream(copy_stream):
            devbuf[s0:s1].copy_(buf[: s1 - s0], non_blocking=True)
            ev.record(copy_stream)
        j += 1
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
rates = gather(pg.numel() * 8 / best * 1e-9)
res["h2d_pageable_via_staging_ring_all_ranks"] = {"per_rank_GBps_min": min(rates), "aggregate_GBps": sum(rates),
                                                  "torch_threads": torch.get_num_threads()}
best = 1e30
for rep in range(2):
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    devbuf[: pg.numel()].copy_(pg, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
rates = gather(pg.numel() * 8 / best * 1e-9)
res["h2d_pageable_cudaMemcpy_all_ranks"] = {"per_rank_GBps_min": min(rates), "aggregate_GBps": sum(rates)}

if rank == 0:
    try:
        import pynvml
        pynvml.nvmlInit()
        topo = []
        for g in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            words = ((os.cpu_count() or 1) + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cp = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
            topo.append({"gpu": g, "cpus": "%d-%d (%d)" % (min(cp), max(cp), len(cp)) if cp else None,
                         "pcie_gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "pcie_width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)})
        res["topology"] = topo
    except Exception as e:  # pragma: no cover
        res["topology"] = str(e)
    res["host_cpus"] = os.cpu_count()
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/%s_h2d_%dranks.json" % (tag, world), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
