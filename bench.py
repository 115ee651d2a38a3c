#!/usr/bin/env python
"""Benchmark of the reduced-basis hot path: M-weighted randomized POD eigensolve (BASELINE.json configs[1],
"applications/confusion output POD: 4096 snapshots x 263k-dof P1 field, rank 256, mass-matrix weighted").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

A step = one complete eigensolve through PODProjectorFromData.construct_subspace(method='randomized'):
mean shift, range finding (SpMM + two DMMA GEMMs + allreduce), M-orthonormalisation, Rayleigh-Ritz, lift and
encoder.  Samples are sharded by GPU (4096 per GPU, weak scaling); inputs (8.6 GB per GPU) exceed L2.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n dofs, samples per GPU, rank, oversampling, modes r0)
    "cfg2": dict(n=263169, n_loc=4096, rank=256, oversampling=10, r0=512,
                 desc="confusion output POD: 4096 snapshots/GPU x 263,169-dof P1 field, rank 256 (+10), M-weighted"),
    "cfg5": dict(n=1002001, n_loc=4096, rank=256, oversampling=10, r0=512,
                 desc="POD scaling sweep shard: 4096 snapshots/GPU x 1,002,001 dofs (1001^2 P1), rank 256 (+10)"),
    "small": dict(n=66049, n_loc=512, rank=64, oversampling=10, r0=128,
                  desc="reduced: 512 snapshots/GPU x 66,049 dofs, rank 64 (+10)"),
}
METRIC = "pod_randomized_eigensolve_fp64_tflops"


def flops_short(n, N, m):
    """Executed-work accounting with the T = W^T W / N shortcut: 6 n N m + 2 N m^2."""
    return 6.0 * n * N * m + 2.0 * N * m * m


def flops_faithful(n, N, m):
    return 8.0 * n * N * m


class ClockSampler:
    """SM clock / power / throttle reasons sampled every 100 ms DURING the timed region through NVML in a
    background thread (same counters as the nvidia-smi clocks line of B200_PROFILING.md; an `nvidia-smi -lms`
    child process was measured to stall CUDA launches by ~25 ms per 100 ms step, NVML in-process does not)."""

    def __init__(self, gpu_index):
        import threading
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        self._active = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index) if bus is None else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active:
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append((sm, pw, rs))
                except Exception:
                    pass
            self._stop.wait(0.1)

    def mark(self):
        self._active = True

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.ok:
            return out
        self._active = False
        self._stop.set()
        self.t.join(timeout=2)
        nv = self.nv
        masks = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        if self.samples:
            sm = sorted(x[0] for x in self.samples)
            top = sm[len(sm) // 2:]
            reasons = sorted(k for k, mk in masks.items() if any(x[2] & mk for x in self.samples))
            out.update(sm_mhz=float(np.median(top)), sm_max_mhz=self.max_sm, reasons=reasons, samples=len(sm),
                       power_w_max=max(x[1] for x in self.samples))
        return out


# ------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_reference_step(sample, seed=0):
    """One eigensolve of the reference's algorithm on the host: hIPPYlib-style doublePassG (column-by-column operator
    applies through CollectiveOperator(NullCollective), MGS B-orthonormalisation, sparse direct B-solves) on a bounded
    sample of the workload.  Returns seconds."""
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P
    n, N, rank, p = sample["n"], sample["N"], sample["rank"], sample["oversampling"]
    M = syn.p1_mass_matrix_for(n)
    u = syn.snapshots(n, N, r0=min(sample["r0"], N), seed=seed)
    Om = syn.gaussian_omega(n, rank + p, seed=1)
    t0 = time.perf_counter()
    d, U, E, shift = P.pod_randomized_weighted(u, M, rank, Om, shifted=True)
    return time.perf_counter() - t0, d


CPU_SAMPLE = dict(n=66049, N=512, rank=256, oversampling=10, r0=512)
# the blocked (BLAS-3) CPU evaluation is fast enough for the full dof count of the workload
CPU_SAMPLE_BLOCKED = dict(n=263169, N=1024, rank=256, oversampling=10, r0=512)


def cpu_blocked_step(sample, seed=0):
    """Best-effort CPU number: the same double pass evaluated block-wise with threaded BLAS-3
    (oracle.projectors_np.pod_randomized_weighted_blocked).  Returns (seconds, d)."""
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P
    n, N, rank, p = sample["n"], sample["N"], sample["rank"], sample["oversampling"]
    M = syn.p1_mass_matrix_for(n)
    u = syn.snapshots(n, N, r0=min(sample["r0"], N), seed=seed)
    Om = syn.gaussian_omega(n, rank + p, seed=1)
    t0 = time.perf_counter()
    d, U, E, shift = P.pod_randomized_weighted_blocked(u, M, rank, Om, shifted=True)
    return time.perf_counter() - t0, d


def cpu_blocked_entry():
    s = CPU_SAMPLE_BLOCKED
    t, _ = cpu_blocked_step(s)
    return {"value": flops_short(s["n"], s["N"], s["rank"] + s["oversampling"]) / t * 1e-12, "unit": "TFLOP/s",
            "seconds": t, "flop_accounting": "6 n N m + 2 N m^2 (executed)",
            "sample": "blocked NumPy/BLAS-3 evaluation of the same double pass on n=%d dofs, N=%d snapshots, rank %d (+%d)"
                      % (s["n"], s["N"], s["rank"], s["oversampling"])}


def sample_desc(s):
    return ("oracle port (NumPy restatement of hIPPYlib doublePassG driven column-by-column like the reference) on "
            "n=%d dofs (257^2 P1 mesh), N=%d snapshots, rank %d (+%d); faithful flop count 8 n N m"
            % (s["n"], s["N"], s["rank"], s["oversampling"]))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    s = CPU_SAMPLE
    m = s["rank"] + s["oversampling"]
    nwarm = min(args.warmup, 1)
    nsteps = max(1, min(args.steps, 3))          # ~30 s of host work per step: keep the whole run to a few minutes
    for _ in range(nwarm):
        cpu_reference_step(s)
    times = []
    for _ in range(nsteps):
        t, _ = cpu_reference_step(s)
        times.append(t)
    T = float(np.mean(times))
    val = flops_faithful(s["n"], s["N"], m) / T * 1e-12
    wl = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": nsteps, "warmup": nwarm, "ms_per_step": T * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "n": wl["n"], "samples_per_gpu": wl["n_loc"],
                       "rank": wl["rank"], "oversampling": wl["oversampling"]},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample_desc(s),
                             "blocked": cpu_blocked_entry()},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bound = None
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    if os.environ.get("HFB_BIND_NUMA", "1") != "0":
        from hippyflow_b200.utilities import bind_to_gpu_numa_node
        bound = bind_to_gpu_numa_node(local_rank)      # pinned staging buffers are first-touched on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import hippyflow_b200 as hf
    from hippyflow_b200 import _lib as K
    from hippyflow_b200 import synthetic as syn

    wl = WORKLOADS[args.workload]
    n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
    m = k + p
    N = n_loc * world
    coll = hf.TorchCollective() if world > 1 else hf.NullCollective()

    M = syn.p1_mass_matrix_for(n)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=7, row_offset=rank * n_loc)
    torch.cuda.synchronize()

    def step_resident():
        return proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, collective=coll,
                                       return_device=True, overwrite_data=True)

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        out = step_resident()
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs
    barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.mark()
    launches0 = K.launch_count()
    segs0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
    K.start_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_resident()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    gemm_times = K.stop_timing()
    launches = K.launch_count() - launches0
    new_segments = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - segs0     # cudaMalloc calls while timing
    clocks = sampler.stop() if sampler is not None else None
    elapsed = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed.item()) / args.steps
    value = flops_short(n, N, m) / (ms_per_step * 1e-3) * 1e-12
    d_last = np.asarray(out[0])

    # ---- roofline of the dominant kernel: the NN tall-skinny DMMA GEMM  W = Xt (M Omega)
    tagNN = (K.HFB_NN, n_loc, m, n)
    tagTN = (K.HFB_TN, n, m, n_loc)
    roof = None
    if tagNN in gemm_times:
        calls, tot = gemm_times[tagNN]
        avg_ms = tot / calls
        peak = K.measure_dmma_peak(dev)
        ach = 2.0 * n * n_loc * m / (avg_ms * 1e-3) * 1e-12
        traffic = None
        tf = os.path.join(ROOT, "profiles", "dgemm_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get(args.workload, {}).get("NN_dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": "dgemm_dmma_kernel<NN,17> (+ split-K reduce)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": "measured live: register-resident DMMA.8x8x4 loop (hfb_measure_dmma_peak); "
                               "MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM 8192^3 measured 35.5 TFLOP/s",
                "launches_timed": calls, "avg_launch_ms": avg_ms,
                "algorithmic_flops_per_launch": 2.0 * n * n_loc * m}
        if tagTN in gemm_times:
            c2, t2 = gemm_times[tagTN]
            roof["second_kernel"] = {"kernel": "dgemm_dmma_kernel<TN,17>", "achieved": 2.0 * n * n_loc * m / (t2 / c2 * 1e-3) * 1e-12,
                                     "avg_launch_ms": t2 / c2, "launches_timed": c2}
        roof["gemm_share_of_step"] = sum(t for tag, (_, t) in gemm_times.items() if tag[0] != "spmm") / (ms_per_step * args.steps)

    # ---- roofline of the HBM-bound kernel on the path: the CSR SpMM  Z = M Q  (north star (b))
    roof_hbm = None
    spmm = [(tag, ct) for tag, ct in gemm_times.items() if tag[0] == "spmm" and tag[3] == m]
    if spmm:
        tag, (calls, tot) = max(spmm, key=lambda x: x[1][0])
        avg_ms = tot / calls
        by = proj.M_device.spmm_bytes(m)
        peak_hbm, src = 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s; MEASURED_PEAKS.json absent)"
        pf = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pf):
            try:
                peak_hbm, src = float(json.load(open(pf))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            except Exception:
                pass
        ach = by / (avg_ms * 1e-3) * 1e-9
        traffic_hbm = None
        tf = os.path.join(ROOT, "profiles", "dgemm_traffic.json")
        if os.path.exists(tf) and tag[1] in ("csr_spmm_dmma_frag_kernel", "csr_spmm_dmma_pipe_kernel"):   # same records, same bytes
            try:
                traffic_hbm = json.load(open(tf)).get(args.workload, {}).get("spmm_frag_dram_bytes_per_launch")
            except Exception:
                traffic_hbm = None
        roof_hbm = {"bound": "hbm", "kernel": tag[1], "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
                    "traffic": traffic_hbm, "peak_source": src, "launches_timed": calls, "avg_launch_ms": avg_ms,
                    "algorithmic_bytes_per_launch": by,
                    "byte_accounting": "nnz*12 + (n+1)*4 + 2*n*m*8 (CSR once, dense block read once, result written once)",
                    "share_of_step": tot / (ms_per_step * args.steps)}

    # ---- end to end: host (pinned) snapshots -> NumPy results, through the reference-facing API
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((n_loc, n), dtype=torch.float64, pin_memory=True)
        except Exception:
            host = torch.empty((n_loc, n), dtype=torch.float64)
        host.copy_(Xt)
        torch.cuda.synchronize()
        n_e2e = max(1, min(args.steps, args.e2e_steps))

        def step_e2e():
            return proj.construct_subspace(host, k, shifted=True, method="randomized", oversampling=p, collective=coll)

        for _ in range(2):                 # warm-up: fills torch's pinned-host cache used for the result arrays
            res = step_e2e()
            del res
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = step_e2e()
            chk = float(res[0][0])             # the eigenvalues / bases are NumPy arrays on the host at this point
            del res
        torch.cuda.synchronize()
        barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        t_e2e = float(el.item()) / n_e2e
        e2e = {"value": flops_short(n, N, m) / t_e2e * 1e-12, "unit": "TFLOP/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(world * n_loc * n * 8),
               "d2h_bytes_per_step": int(world * (2 * n * k + n + k) * 8), "steps": n_e2e,
               "api": "PODProjectorFromData.construct_subspace(host array, method='randomized') -> NumPy (d, phi, Mphi, u_shift)",
               "host_cpus_bound": (len(bound) if bound else None)}
        del host

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if bound and full_affinity:
            os.sched_setaffinity(0, full_affinity)      # the CPU baseline may use every host core
        s = CPU_SAMPLE
        t, _ = cpu_reference_step(s)
        cpu = {"value": flops_faithful(s["n"], s["N"], s["rank"] + s["oversampling"]) / t * 1e-12, "unit": "TFLOP/s",
               "cores": os.cpu_count(), "kind": "port", "sample": sample_desc(s), "seconds": t,
               "blocked": cpu_blocked_entry()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["desc"], "name": args.workload, "n": n, "samples_per_gpu": n_loc,
                           "samples_total": N, "rank": k, "oversampling": p, "parallelism": "sample-sharded x%d" % world,
                           "flops_per_step": flops_short(n, N, m), "flop_accounting": "6 n N m + 2 N m^2 (executed GEMM work)",
                           "l2": "inputs (%.1f GB/GPU) exceed L2" % (n_loc * n * 8 / 1e9)},
                "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "cuda_mallocs_in_timed_region": int(new_segments), "clocks": clocks,
                "eigenvalues_head": [float(x) for x in d_last[:3]]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
