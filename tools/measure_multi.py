"""Multi-GPU (torchrun) timing of the two BASELINE configs that need a full box:
   cfg5: POD scaling sweep, n = 1,002,001 dofs x 32,768 snapshots (263 GB fp64) sharded by sample, rank 256 + 10
   cfg3: active subspace from stored Jacobians, 4096 samples x 100 obs x 65,536 params (215 GB), rank 200 + 10
Each rank generates its own shard on the device; rank 0 prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
coll = hf.TorchCollective() if world > 1 else hf.NullCollective()
which = sys.argv[1:] or ["cfg5", "cfg3"]

def timed(f, n=3):
    f(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    return min(ts)

if "cfg5" in which:
    n, N, k, p = 1002001, 32768, 256, 10
    n_loc = N // world if world > 1 else 4096
    M = syn.p1_mass_matrix_for(n)
    Xt = syn.snapshots_device(n, n_loc, dev, r0=512, seed=7, row_offset=rank * n_loc)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    out = {}
    def solve():
        out["r"] = proj.construct_subspace(Xt, k, shifted=True, method="randomized", oversampling=p, collective=coll, return_device=True, overwrite_data=True)
    t = timed(solve)
    m = k + p; Ntot = n_loc * world
    fl = 6.0 * n * Ntot * m + 2.0 * Ntot * m * m
    d, phi, Mphi, _ = out["r"]
    G = K.dgemm(K.HFB_TN, phi, Mphi).cpu().numpy()
    if rank == 0:
        print(json.dumps({"config": "cfg5", "n_gpus": world, "n": n, "samples_total": Ntot, "samples_per_gpu": n_loc, "rank": k, "ms": t * 1e3,
                          "tflops_executed": fl / t * 1e-12, "tflops_per_gpu": fl / t * 1e-12 / world, "orth_err": float(np.abs(G - np.eye(k)).max()),
                          "d_head": [float(x) for x in d[:3]]}), flush=True)
    del Xt, proj, out, phi, Mphi
    torch.cuda.empty_cache()

if "cfg3" in which:
    N, dQ, dM, k, p = 4096, 100, 65536, 200, 10
    n_loc = N // world if world > 1 else 1024
    J = torch.empty((n_loc * dQ, dM), dtype=torch.float64, device=dev)
    for i0 in range(0, n_loc * dQ, 16384):
        K.fill_random_(J[i0:i0 + 16384], 31, row_offset=rank * n_loc * dQ + i0)
    sc = (1.0 + torch.arange(dM, device=dev, dtype=torch.float64)) ** -0.5
    for i0 in range(0, n_loc * dQ, 16384):
        K.colscale_(J[i0:i0 + 16384], sc)
    params = hf.ActiveSubspaceParameterList(); params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = k, p, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J.view(n_loc, dQ, dM)), None, collective=coll, parameters=params, device=dev)
    out = {}
    def solve():
        out["r"] = proj.construct_input_subspace(prior_preconditioned=False)
    t = timed(solve, n=2)
    m = k + p; Ntot = n_loc * world
    fl = 6.0 * dM * Ntot * dQ * m + 2.0 * Ntot * dQ * m * m
    if rank == 0:
        print(json.dumps({"config": "cfg3", "n_gpus": world, "samples_total": Ntot, "dQ": dQ, "dM": dM, "rank": k, "ms": t * 1e3,
                          "tflops_executed": fl / t * 1e-12, "tflops_per_gpu": fl / t * 1e-12 / world,
                          "d_head": [float(x) for x in out["r"][0][:3]]}), flush=True)
if world > 1:
    dist.destroy_process_group()
