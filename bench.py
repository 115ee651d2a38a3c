#!/usr/bin/env python
"""Benchmark of the reduced-basis hot path: M-weighted randomized POD eigensolve (BASELINE.json configs[1],
"applications/confusion output POD: 4096 snapshots x 263k-dof P1 field, rank 256, mass-matrix weighted").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

A step = one complete eigensolve through PODProjectorFromData.construct_subspace(method='randomized'):
mean shift, range finding (SpMM + two DMMA GEMMs + allreduce), M-orthonormalisation, Rayleigh-Ritz, lift and
encoder.  Samples are sharded by GPU (4096 per GPU, weak scaling); inputs (8.6 GB per GPU) exceed L2.
Prints ONE JSON line (rank 0).

Both arms report the SAME metric on the SAME config with the SAME flop accounting (executed GEMM work,
6 n N m + 2 N m^2).  The reference arm evaluates the reference's algorithm (oracle port; the reference itself needs
FEniCS/hIPPYlib and does not travel to the GPU box) with threaded BLAS-3 on the full config at N = 1; for N > 1 it runs
on rank 0 on one shard's worth of samples (cost is linear in the sample count) and says so.
"""
import argparse
import json
import os
import sys
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
FLOP_ACCOUNTING = "6 n N m + 2 N m^2 (executed GEMM work)"
DATA_SEED, OMEGA_SEED = 7, 1


def flops_short(n, N, m):
    """Executed-work accounting with the T = W^T W / N shortcut: 6 n N m + 2 N m^2."""
    return 6.0 * n * N * m + 2.0 * N * m * m


def config_dict(name, world):
    """The `config` object of the JSON line -- identical for both arms."""
    wl = WORKLOADS[name]
    n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
    N = n_loc * world
    return {"workload": wl["desc"], "name": name, "n": n, "samples_per_gpu": n_loc, "samples_total": N, "rank": k,
            "oversampling": p, "parallelism": "sample-sharded x%d" % world, "flops_per_step": flops_short(n, N, k + p),
            "flop_accounting": FLOP_ACCOUNTING, "l2": "inputs (%.1f GB/GPU) exceed L2" % (n_loc * n * 8 / 1e9)}


def blas_vendor():
    try:
        b = np.show_config(mode="dicts")["Build Dependencies"]["blas"]
        return "%s %s" % (b.get("name"), b.get("version"))
    except Exception:
        return "unknown"


def principal_angle(U, V, M=None):
    """Largest principal angle between span(U) and span(V) in the M inner product (NumPy, host)."""
    def orth(A):
        G = A.T @ (A if M is None else M @ A)
        L = np.linalg.cholesky((G + G.T) / 2)
        return np.linalg.solve(L, A.T).T
    Uo, Vo = orth(U), orth(V)
    R = Vo - Uo @ (Uo.T @ (Vo if M is None else M @ Vo))
    G = R.T @ (R if M is None else M @ R)
    s = np.sqrt(max(np.linalg.eigvalsh((G + G.T) / 2).max(), 0.0))
    return float(np.arcsin(min(s, 1.0)))


def eig_parity(d, d0, phi, phi0, M, floor=1e-5):
    """{max_rel_eig over the leading modes (lambda_i/lambda_1 > floor), max abs error / lambda_1 over all, angle}."""
    k = int(np.sum(d0 / d0[0] > floor))
    return {"leading_modes": k, "max_rel_eig": float(np.max(np.abs(d[:k] - d0[:k]) / d0[:k])),
            "max_abs_eig_over_lambda1": float(np.max(np.abs(d - d0)) / d0[0]),
            "angle": principal_angle(phi[:, :k], phi0[:, :k], M)}


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
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
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


# =========================================================================================== CPU legs (oracle = the checker)
def cpu_blocked_solve(u, M, rank, Om):
    """The reference's algorithm (M-weighted double pass, PODProjector.py:750-761 + KLEProjector.py:163-168 solver
    pattern) evaluated with threaded BLAS-3 on the host: oracle.projectors_np.pod_randomized_weighted_blocked.
    Returns (seconds, d, U)."""
    from oracle import projectors_np as P
    t0 = time.perf_counter()
    d, U, E, shift = P.pod_randomized_weighted_blocked(u, M, rank, Om, shifted=True)
    return time.perf_counter() - t0, d, U


def cpu_column_by_column_entry():
    """The column-by-column port (hIPPYlib-style doublePassG: one operator apply per column through
    CollectiveOperator(NullCollective), MGS B-orthonormalisation, sparse direct B-solves) -- the reference's BLAS-2 call
    structure -- on a reduced sample, reported with the SAME executed-work flop accounting as every other number."""
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P
    s = dict(n=66049, N=256, rank=256, oversampling=10, r0=256)
    M = syn.p1_mass_matrix_for(s["n"])
    u = syn.snapshots(s["n"], s["N"], r0=s["r0"], seed=0)
    m = s["rank"] + s["oversampling"]
    Om = syn.gaussian_omega(s["n"], m, seed=1)
    t0 = time.perf_counter()
    P.pod_randomized_weighted(u, M, s["rank"], Om, shifted=True)
    t = time.perf_counter() - t0
    return {"value": flops_short(s["n"], s["N"], m) / t * 1e-12, "unit": "TFLOP/s", "seconds": t,
            "flop_accounting": FLOP_ACCOUNTING + "; the port itself executes 8 n N m (no T shortcut)",
            "sample": "column-by-column oracle port (NumPy restatement of hIPPYlib doublePassG driven like "
                      "PODProjector.py:360-376) on n=%d dofs, N=%d snapshots, rank %d (+%d)"
                      % (s["n"], s["N"], s["rank"], s["oversampling"])}


def host_snapshots(wl, n_rows, seed):
    from hippyflow_b200 import synthetic as syn
    return syn.snapshots(wl["n"], n_rows, r0=wl["r0"], seed=seed)


def run_reference(args):
    """Reference arm: the reference's CPU algorithm on this box's host cores, same metric / config / flop accounting."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hippyflow_b200 import synthetic as syn
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    wl = WORKLOADS[args.workload]
    n, k, p = wl["n"], wl["rank"], wl["oversampling"]
    m = k + p
    n_rows = wl["n_loc"]                     # N = 1: the full config; N > 1: one shard's worth (bounded sample)
    cores = os.cpu_count()
    t_gen = time.perf_counter()
    M = syn.p1_mass_matrix_for(n)
    u = host_snapshots(wl, n_rows, DATA_SEED)
    Om = syn.gaussian_omega(n, m, seed=OMEGA_SEED)
    t_gen = time.perf_counter() - t_gen
    # time budget: every step is the full solve; when K + W full steps do not fit the budget the step count is cut and
    # the line says so (steps = what ran, steps_requested = K)
    budget = float(args.ref_budget_s)
    t_first, d, _ = cpu_blocked_solve(u, M, k, Om)
    want = args.steps + args.warmup
    nwarm, nsteps = args.warmup, args.steps
    if want * t_first > budget:
        nwarm = 1
        nsteps = int(max(1, min(args.steps, (budget - t_first) // t_first)))
    times = []
    for i in range(max(0, nwarm - 1)):        # the first solve above was warm-up step 1
        cpu_blocked_solve(u, M, k, Om)
    for _ in range(nsteps):
        t, d, _ = cpu_blocked_solve(u, M, k, Om)
        times.append(t)
    T = float(np.mean(times))
    val_sample = flops_short(n, n_rows, m) / T * 1e-12      # TFLOP/s: a rate, independent of the sample count
    sample = ("blocked BLAS-3 evaluation (oracle.projectors_np.pod_randomized_weighted_blocked) of the M-weighted double "
              "pass on n=%d dofs x N=%d snapshots, rank %d (+%d)" % (n, n_rows, k, p))
    if world > 1:
        sample += ("; N > 1: one process on rank 0, one shard's worth of samples (%d of %d) -- cost is linear in the "
                   "sample count, so the TFLOP/s rate is that of the full config" % (n_rows, n_rows * world))
    else:
        sample += " = the full config"
    line = {"impl": "reference", "metric": METRIC, "value": val_sample, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": nsteps, "warmup": nwarm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": T * 1e3 * (world if world > 1 else 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, world),
            "cpu_baseline": {"value": val_sample, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                             "seconds_per_step": T, "blas": blas_vendor(), "host_data_generation_s": t_gen,
                             "flop_accounting": FLOP_ACCOUNTING,
                             "step_cap": None if (nsteps == args.steps and nwarm == args.warmup) else
                             "cut to %d+%d steps to stay within %.0f s (first step took %.1f s)" % (nwarm, nsteps, budget, t_first),
                             "column_by_column": None if args.no_column_port else cpu_column_by_column_entry()},
            "e2e": {"value": val_sample, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "eigenvalues_head": [float(x) for x in d[:3]]}
    if world > 1:
        line["ms_per_step_note"] = "measured step (one shard) x world: the whole-job step time of a 1-process CPU run"
    print(json.dumps(line))


# =========================================================================================== CUDA arm
def _timed_steps(fn, steps, dist, world, dev, torch):
    """Time `steps` calls of fn on the device: barrier + synchronize on both sides, CUDA events, MAX over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    el = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    return float(el.item()) / steps, out


def sharded_parity_leg(hf, K, syn, coll, dev, rank, world, torch, dist):
    """Driver-visible multi-GPU correctness: a small weighted POD (n = 66,049, 256 snapshots per rank, rank 64 + 10)
    solved (a) sharded over all ranks through the collective and (b) by rank 0 alone on the concatenated data with
    NullCollective -- same Omega, data keyed by the global sample index."""
    n, per, k, p = 66049, 256, 64, 10
    M = syn.p1_mass_matrix_for(n)
    from hippyflow_b200.modeling.PODProjector import gaussian_omega
    Om = gaussian_omega(n, k + p, 5, dev)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    shard = syn.snapshots_device(n, per, dev, r0=128, seed=11, row_offset=rank * per)
    shard += 0.3                                                              # a non-trivial mean
    d, phi, Mphi, shift = proj.construct_subspace(shard, k, shifted=True, method="randomized", Omega=Om, collective=coll,
                                                  return_device=True)
    res = None
    if rank == 0:
        full = syn.snapshots_device(n, per * world, dev, r0=128, seed=11, row_offset=0)
        full += 0.3
        d1, phi1, _, shift1 = proj.construct_subspace(full, k, shifted=True, method="randomized", Omega=Om,
                                                      collective=hf.NullCollective(), return_device=True)
        res = eig_parity(np.asarray(d), np.asarray(d1), phi.cpu().numpy(), phi1.cpu().numpy(), M)
        res["shift_max_abs_diff"] = float((shift - shift1).abs().max())
        res.update(n=n, samples_per_rank=per, rank=k, oversampling=p, world=world,
                   what="sharded solve over the collective vs rank-0 solve of the concatenated data, same Omega")
    if world > 1:
        dist.barrier()
    return res


def target_cfg5_shard(hf, K, syn, coll, dev, rank, world, torch, dist, peak, steps=3):
    """North-star target 1: POD scaling sweep, n = 1,002,001 dofs x 4096 snapshots per GPU (32,768 at 8 GPUs), rank 256."""
    n, n_loc, k, p = 1002001, 4096, 256, 10
    m = k + p
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 60e9:
        return {"skipped": "needs ~60 GB of free device memory"}
    M = syn.p1_mass_matrix_for(n)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    Xt = syn.snapshots_device(n, n_loc, dev, r0=512, seed=DATA_SEED, row_offset=rank * n_loc)

    def step():
        return proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, collective=coll,
                                       return_device=True, overwrite_data=True)
    step()
    ms, out = _timed_steps(step, steps, dist, world, dev, torch)
    N = n_loc * world
    tf = flops_short(n, N, m) / (ms * 1e-3) * 1e-12
    d, phi, Mphi, _ = out
    G = K.dgemm(K.HFB_TN, phi, Mphi).cpu().numpy()
    res = {"workload": "POD scaling sweep: %d dofs x %d snapshots (%d per GPU), rank %d (+%d), M-weighted" % (n, N, n_loc, k, p),
           "ms_per_step": ms, "steps": steps, "tflops": tf, "tflops_per_gpu": tf / world, "frac_of_dmma_peak_per_gpu": tf / world / peak,
           "orthonormality_max_abs": float(np.abs(G - np.eye(k)).max()), "eigenvalues_head": [float(x) for x in np.asarray(d)[:3]]}
    del Xt, proj, out, phi, Mphi
    torch.cuda.empty_cache()
    return res


def jacobians_device(K, torch, n_loc, dQ, dM, dev, row_offset):
    """(n_loc*dQ, dM) stored-Jacobian shard generated in HBM, keyed by the global row (independent of the sharding),
    columns scaled smoothly so that mean J^T J has a decaying spectrum."""
    J = torch.empty((n_loc * dQ, dM), dtype=torch.float64, device=dev)
    sc = (1.0 + torch.arange(dM, device=dev, dtype=torch.float64)) ** -0.5
    for i0 in range(0, n_loc * dQ, 16384):
        K.fill_random_(J[i0:i0 + 16384], 31, row_offset=row_offset * dQ + i0)
        K.colscale_(J[i0:i0 + 16384], sc)
    return J


def target_cfg3_shard(hf, K, syn, coll, dev, rank, world, torch, dist, peak, steps=3):
    """North-star target 2: active subspace from stored Jacobians, 512 samples per GPU (4096 at 8 GPUs) x 100 x 65,536,
    rank 200 (+10), doublePass (prior_preconditioned=False)."""
    n_loc, dQ, dM, k, p = 512, 100, 65536, 200, 10
    m = k + p
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 45e9:
        return {"skipped": "needs ~45 GB of free device memory"}
    J = jacobians_device(K, torch, n_loc, dQ, dM, dev, rank * n_loc)
    params = hf.ActiveSubspaceParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = k, p, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J.view(n_loc, dQ, dM)), None, collective=coll, parameters=params,
                                      device=dev)

    def step():
        return proj.construct_input_subspace(prior_preconditioned=False)
    step()
    ms, out = _timed_steps(step, steps, dist, world, dev, torch)
    N = n_loc * world
    fl = 6.0 * dM * N * dQ * m + 2.0 * N * dQ * m * m
    tf = fl / (ms * 1e-3) * 1e-12
    V = out[1].tensor()
    G = K.dgemm(K.HFB_TN, V, V).cpu().numpy()
    res = {"workload": "active subspace from stored Jacobians: %d samples (%d per GPU) x %d x %d, rank %d (+%d)" % (N, n_loc, dQ, dM, k, p),
           "ms_per_step": ms, "steps": steps, "tflops": tf, "tflops_per_gpu": tf / world, "frac_of_dmma_peak_per_gpu": tf / world / peak,
           "flop_accounting": "6 dM N dQ m + 2 N dQ m^2 (executed)", "orthonormality_max_abs": float(np.abs(G - np.eye(k)).max()),
           "eigenvalues_head": [float(x) for x in np.asarray(out[0])[:3]]}
    del J, proj, out, V
    torch.cuda.empty_cache()
    return res


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
    from hippyflow_b200.linalg import CsrMatrix
    from hippyflow_b200.modeling.PODProjector import gaussian_omega

    wl = WORKLOADS[args.workload]
    n, n_loc, k, p = wl["n"], wl["n_loc"], wl["rank"], wl["oversampling"]
    m = k + p
    N = n_loc * world
    coll = hf.TorchCollective() if world > 1 else hf.NullCollective()

    M = syn.p1_mass_matrix_for(n)
    # one-time cost of a cold call: CSR upload + SpMM plan (row clustering, run records) -- outside the timed regions
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    probe = K.padded_zeros(n, m, dev)
    proj.M_device.matmat(probe)
    torch.cuda.synchronize()
    plan_build_ms = (time.perf_counter() - t0) * 1e3
    del probe
    Xt = syn.snapshots_device(n, n_loc, dev, r0=wl["r0"], seed=DATA_SEED, row_offset=rank * n_loc)
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
    peak = K.measure_dmma_peak(dev)

    # ---- roofline of the dominant kernel: the NN tall-skinny DMMA GEMM  W = Xt (M Omega)
    tagNN = (K.HFB_NN, n_loc, m, n)
    tagTN = (K.HFB_TN, n, m, n_loc)
    roof = None
    if tagNN in gemm_times:
        calls, tot = gemm_times[tagNN]
        avg_ms = tot / calls
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
                "peak_source": "measured live: register-resident DMMA.8x8x4 loop (hfb_measure_dmma_peak) = the arithmetic ceiling "
                               "148 SM x 64 FMA/clk x 1.965 GHz = 37.2; MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM 8192^3 "
                               "measured 35.5 TFLOP/s (profiles/r01_fp64_peaks.md)",
                "launches_timed": calls, "avg_launch_ms": avg_ms,
                "algorithmic_flops_per_launch": 2.0 * n * n_loc * m}
        if tagTN in gemm_times:
            c2, t2 = gemm_times[tagTN]
            roof["second_kernel"] = {"kernel": "dgemm_dmma_kernel<TN,17>", "achieved": 2.0 * n * n_loc * m / (t2 / c2 * 1e-3) * 1e-12,
                                     "avg_launch_ms": t2 / c2, "launches_timed": c2}
        roof["gemm_share_of_step"] = sum(t for tag, (_, t) in gemm_times.items() if tag[0] != "spmm") / (ms_per_step * args.steps)
        roof["ideal_step_ms_at_dmma_peak"] = flops_short(n, n_loc, m) / (peak * 1e12) * 1e3
        roof["step_frac_of_dmma_peak"] = value / world / peak

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
        if os.path.exists(tf):
            try:
                traffic_hbm = json.load(open(tf)).get(args.workload, {}).get("spmm_dram_bytes_per_launch", {}).get(tag[1])
            except Exception:
                traffic_hbm = None
        roof_hbm = {"bound": "hbm", "kernel": tag[1], "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
                    "traffic": traffic_hbm, "peak_source": src, "launches_timed": calls, "avg_launch_ms": avg_ms,
                    "algorithmic_bytes_per_launch": by,
                    "byte_accounting": "nnz*12 + (n+1)*4 + 2*n*m*8 (CSR once, dense block read once, result written once)",
                    "share_of_step": tot / (ms_per_step * args.steps)}

    # ---- end to end: host snapshots -> NumPy results, through the reference-facing API
    e2e = None
    host = None
    if not args.no_e2e:
        try:
            host = torch.empty((n_loc, n), dtype=torch.float64, pin_memory=True)
        except Exception:
            host = torch.empty((n_loc, n), dtype=torch.float64)
        host.copy_(Xt)
        torch.cuda.synchronize()
        n_e2e = max(1, min(args.steps, args.e2e_steps))

        def step_e2e(src):
            return proj.construct_subspace(src, k, shifted=True, method="randomized", oversampling=p, collective=coll)

        def time_e2e(src, reps, warm):
            for _ in range(warm):              # warm-up: fills torch's pinned-host cache used for the result arrays
                res = step_e2e(src)
                del res
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                res = step_e2e(src)
                chk = float(res[0][0])         # the eigenvalues / bases are NumPy arrays on the host at this point
                del res
            torch.cuda.synchronize()
            barrier()
            el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            return float(el.item()) / reps

        t_e2e = time_e2e(host, n_e2e, 2)
        # floor of this leg: the copies alone, all ranks at once (no compute): pinned host -> device of the snapshots and
        # device -> pinned host of the results
        dst = K.padded_empty(n_loc, n, dev)
        res_d = K.padded_empty(n, 2 * k + 1, dev)
        res_h = torch.empty((n, 2 * k + 1), dtype=torch.float64, pin_memory=True)
        floors = []
        for _ in range(2):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            res_h.copy_(res_d[:, :2 * k + 1], non_blocking=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            el = torch.tensor([t1 - t0, t2 - t1], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            floors.append(el.cpu().numpy())
        h2d_s, d2h_s = floors[-1]
        del dst, res_d, res_h
        e2e = {"value": flops_short(n, N, m) / t_e2e * 1e-12, "unit": "TFLOP/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(world * n_loc * n * 8),
               "d2h_bytes_per_step": int(world * (2 * n * k + n + k) * 8), "steps": n_e2e,
               "api": "PODProjectorFromData.construct_subspace(host array, method='randomized') -> NumPy (d, phi, Mphi, u_shift)",
               "input": "pinned host tensor (best case); see pageable_ms_per_step for a plain NumPy array",
               "h2d_only_ms": h2d_s * 1e3, "d2h_only_ms": d2h_s * 1e3,
               "h2d_GBps_per_rank": n_loc * n * 8 / h2d_s * 1e-9, "h2d_GBps_aggregate": world * n_loc * n * 8 / h2d_s * 1e-9,
               "e2e_floor_ms": (h2d_s + d2h_s) * 1e3,
               "e2e_floor_note": "concurrent copies alone on all ranks, max over ranks: the host-side ceiling of this leg",
               "plan_build_ms": plan_build_ms,
               "plan_build_note": "one-time per mass matrix (CSR upload + SpMM row clustering + run records); paid by a "
                                  "cold first call, not part of any timed region",
               "host_cpus_bound": (len(bound) if bound else None)}
        if not args.no_pageable:
            # the documented drop-in call passes a plain (pageable) NumPy array
            u_np = host.numpy().copy()
            e2e["pageable_ms_per_step"] = time_e2e(u_np, 2, 1) * 1e3
            del u_np

    # ---- driver-visible multi-GPU correctness (N > 1) and the north-star target shards
    sharded = None
    if world > 1 and not args.no_parity:
        sharded = sharded_parity_leg(hf, K, syn, coll, dev, rank, world, torch, dist)

    # ---- CPU baseline and full-size parity (rank 0, N = 1 only)
    cpu = None
    parity_full = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if bound and full_affinity:
            os.sched_setaffinity(0, full_affinity)      # the CPU baseline may use every host core
        from oracle import projectors_np as P
        # identical inputs on both sides: the device snapshots copied to the host, Omega generated once and shared
        u_host = host.numpy() if host is not None else Xt.cpu().numpy()
        Om_d = gaussian_omega(n, m, OMEGA_SEED, dev)
        Om_h = Om_d.to_dense()
        t_cpu, d_cpu, U_cpu = cpu_blocked_solve(u_host, M, k, Om_h)
        d_g, phi_g, Mphi_g, _ = proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, Omega=Om_d,
                                                        return_device=True)
        parity_full = eig_parity(np.asarray(d_g), d_cpu, phi_g.cpu().numpy(), U_cpu, M)
        parity_full["what"] = ("CUDA path vs blocked CPU oracle on the IDENTICAL full-size inputs (n=%d, N=%d, rank %d, same Omega)"
                               % (n, n_loc, k))
        cpu = {"value": flops_short(n, n_loc, m) / t_cpu * 1e-12, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
               "blas": blas_vendor(), "seconds": t_cpu, "flop_accounting": FLOP_ACCOUNTING,
               "sample": "the full config (n=%d dofs x N=%d snapshots, rank %d (+%d)): blocked BLAS-3 evaluation of the reference's "
                         "M-weighted double pass (oracle.projectors_np.pod_randomized_weighted_blocked), one solve" % (n, n_loc, k, p)}
        del U_cpu
        if not args.no_hep:
            # the reference's OWN method on the full config: PODProjectorFromData.construct_subspace(method='hep'),
            # PODProjector.py:812-833, restated in oracle.projectors_np.pod_from_data -- and the GPU 'hep' beside it
            t0 = time.perf_counter()
            d_h, _, _, _ = P.pod_from_data(u_host, M, k, shifted=True, method="hep")
            cpu["reference_hep_seconds"] = time.perf_counter() - t0
            torch.cuda.synchronize()
            for rep in range(2):
                t0 = time.perf_counter()
                d_gh, phi_gh, _, _ = proj.construct_subspace(Xt, k, shifted=True, method="hep", return_device=True)
                torch.cuda.synchronize()
                cpu["gpu_hep_seconds"] = time.perf_counter() - t0
            cpu["hep_max_rel_eig_diff"] = float(np.max(np.abs(np.asarray(d_gh)[:64] - d_h[:64]) / d_h[:64]))
            cpu["hep_note"] = ("method of snapshots on the full config: CPU = oracle restatement of PODProjector.py:812-833 "
                               "(Gram GEMM 2 n N^2 + LAPACK eigh(N) + lift); GPU = same method through this package "
                               "(device-resident input); eigenvalue difference over the 64 leading modes")
            del phi_gh
    del host

    # ---- north-star targets on this GPU count (secondary legs, 3 steps each)
    targets = None
    if not args.no_targets:
        del Xt, out
        torch.cuda.empty_cache()
        targets = {"cfg5_shard": target_cfg5_shard(hf, K, syn, coll, dev, rank, world, torch, dist, peak),
                   "cfg3_shard": target_cfg3_shard(hf, K, syn, coll, dev, rank, world, torch, dist, peak)}

    # ---- how the (n x m) sketch travelled between the GPUs
    exchange = None
    if world > 1:
        exs = {key: e for key, e in getattr(coll, "_peer_exchanges", {}).items()}
        mine = [e for key, e in exs.items() if key[0] == n]
        if mine and mine[0] is not None:
            e = mine[0]
            exchange = {"route": "peer",
                        "kernels": "dgemm_dmma_kernel<TN,*,PEER>: 256-bit st.global of every tile into the owner's slot (CUDA-IPC peer "
                                   "address) -> peer_barrier -> peer_reduce_bcast: fixed-rank-order sum, stored into every rank's "
                                   "result block over NVLink -> peer_barrier; the solver adopts the result block as its sketch",
                        "rows_per_owner": e.block, "verified_against_nccl_rel_err": getattr(e, "verify_err", None),
                        "nvlink_bytes_pushed_per_rank_per_step": 2 * e.n * e.ld * 8.0 * (world - 1) / world,
                        "nccl_on_data_path": "no (m x m Rayleigh matrix, one m-vector and one scalar only)"}
        else:
            exchange = {"route": "nccl", "note": "row blocks of the lift, asynchronous allreduce per block (peer route "
                                 "unavailable or disabled: HFB_PEER_LIFT=%s)" % os.environ.get("HFB_PEER_LIFT", "1")}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args.workload, world),
                "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "cuda_mallocs_in_timed_region": int(new_segments), "clocks": clocks,
                "parity_full_size": parity_full, "sharded_parity": sharded, "exchange": exchange, "targets": targets,
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
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hep", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-targets", action="store_true")
    ap.add_argument("--no-column-port", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=330.0,
                    help="reference arm: wall-clock budget for warm-up + timed steps; the step count is cut (and the line says so) beyond it")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
