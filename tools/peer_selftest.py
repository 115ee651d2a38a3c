"""Single-GPU self-test of the NVLink sketch exchange (hippyflow_b200/peer.py): P emulated ranks inside one process,
every 'peer' buffer local, one stream per rank.  Exercises the peer-store epilogue of the lift GEMM (hfb_dgemm_peer), the
flag barrier, the fixed-order slot reduction and the gather kernel against the sum of P plain lifts.  Run in a subprocess by
tests/test_gpu_kernels.py (a barrier time-out traps the kernel and would poison the caller's CUDA context)."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per emulated rank's stream

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hippyflow_b200 import _lib as K          # noqa: E402
from hippyflow_b200.peer import PeerExchange, plan_chunks  # noqa: E402


def run_case(P, n, R, ncols, nchunk, dev):
    g = torch.Generator(device="cpu").manual_seed(1000 * P + ncols)
    Xs = [K.to_padded(torch.randn((R, n), generator=g, dtype=torch.float64), dev) for _ in range(P)]
    Ws = [K.to_padded(torch.randn((R, ncols), generator=g, dtype=torch.float64), dev) for _ in range(P)]
    alpha = 0.37
    ref = None
    for X, W in zip(Xs, Ws):                       # rank order, like the owner's fixed-order sum
        part = K.dgemm(K.HFB_TN, X, W, alpha=alpha)
        ref = part.clone() if ref is None else ref + part
    ld = ((ncols + 15) // 16) * 16
    group = PeerExchange.local_group(P, dev, n, ld, ncols, nchunk)
    Ys = [K.padded_zeros(n, ncols, dev) for _ in range(P)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    torch.cuda.synchronize()
    for rep in range(2):                           # the second exchange reuses the slots and flags (growing epochs)
        for r in range(P):
            with torch.cuda.stream(streams[r]):
                if rep:
                    Ys[r].zero_()
                group[r].lift_allreduce(Xs[r], Ws[r], Ys[r], alpha)
        torch.cuda.synchronize()
        for r in range(P):
            assert torch.equal(Ys[r], Ys[0]), "ranks disagree bitwise"
            assert torch.equal(Ys[r], ref), (P, n, ncols, nchunk, float((Ys[r] - ref).abs().max()))
    for ex in group[1:]:
        ex.own = None                              # the buffers are shared: free each once
    chunks = len(group[0].chunks)
    for r, ex in enumerate(group):
        ex.base = None
        if ex.own is not None:
            K.peer_free(ex.own)
    return chunks


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    before = K.launch_count()
    cases = [(2, 5041, 64, 27, 1), (3, 66049, 96, 67, 1), (2, 66049, 64, 66, 3), (4, 40000, 48, 138, 2)]
    for P, n, R, ncols, nchunk in cases:
        got = run_case(P, n, R, ncols, nchunk, dev)
        print("case P=%d n=%d cols=%d chunks=%d ok" % (P, n, ncols, got))
    assert plan_chunks(263169, 267, 4, 8) and all(b % 128 == 0 for _, _, b in plan_chunks(263169, 267, 4, 8))
    print("PEER_SELFTEST_OK launches=%d" % (K.launch_count() - before))


if __name__ == "__main__":
    main()
