"""Times the sketch lift + exchange at a benchmark shape under torchrun (one rank per GPU):
plain lift GEMM (no exchange) | lift in row blocks + NCCL allreduce (r1/r2 route) | fused lift + NVLink peer exchange (and its phases alone).  Prints one JSON line on rank 0 (times = max over ranks of the CUDA-event mean per call)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps, dev):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import hippyflow_b200 as hf
    from hippyflow_b200 import _lib as K
    from hippyflow_b200.linalg import SampleCovariance
    from hippyflow_b200.peer import PeerExchange
    n = int(os.environ.get("PL_N", 263169))
    R = int(os.environ.get("PL_ROWS", 4096))
    ncols = int(os.environ.get("PL_COLS", 267))
    reps = int(os.environ.get("PL_REPS", 5))
    X = K.padded_empty(R, n, dev)
    K.fill_random_(X, 7, row_offset=rank * R)
    W = K.padded_empty(R, ncols, dev)
    K.fill_random_(W, 8, row_offset=rank * R)
    coll = hf.MultipleSerialPDEsCollective()
    op = hf.SampleCovarianceOperator(SampleCovariance(X), coll, "avg")
    Y = hf.DeviceMultiVector(n, ncols, device=dev)
    Yt = Y.tensor()
    scale = 1.0 / (R * world)
    out = {"n": n, "rows_per_rank": R, "cols": ncols, "world": world, "reps": reps}
    out["gemm_only_ms"] = timed(lambda: K.dgemm(K.HFB_TN, X, W, out=Yt, alpha=scale), reps, dev)
    for nchunk in (1, 4):
        out["nccl_chunks%d_ms" % nchunk] = timed(lambda: op._lift_nccl(W, Y, Yt, scale, False, nchunk), reps, dev)
    ref = Yt.clone()
    ex = PeerExchange.create(None, dev, n, K._ld(Yt), ncols)
    if ex is None:
        out["peer_ms"] = None
    else:
        out["peer_ms"] = timed(lambda: ex.lift_allreduce(X, W, scale), reps, dev)
        out["peer_err_vs_nccl"] = float((ex.result_view() - ref).abs().max() / ref.abs().max())
        # phases alone (each rank times its own kernels; the barriers absorb the skew)
        out["peer_push_gemm_ms"] = timed(lambda: ex.push(X, W, scale), reps, dev)

        def red():
            ex.reduce_bcast()
            ex.barrier()
        out["peer_reduce_bcast_ms"] = timed(red, reps, dev)
        out["peer_barrier_ms"] = timed(lambda: ex.barrier(), reps, dev)
        from hippyflow_b200.peer import release_all
        release_all([ex], None)
    nbytes = n * K._ld(Yt) * 8
    out["exchange_bytes_per_rank_each_way"] = nbytes * (world - 1) / world
    if rank == 0:
        print("PEER_LIFT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
