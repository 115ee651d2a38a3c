"""Single-GPU self-test of the NVLink sketch exchange (hippyflow_b200/peer.py): P emulated ranks inside one process, every
'peer' buffer local.  The phases run rank by rank on ONE stream (push of all ranks, signal of all ranks, wait of all ranks,
reduce, ...), so no kernel ever waits for work that is queued behind it.  Exercises the peer-store epilogue of the lift GEMM
(hfb_dgemm_peer), the flag words and epochs and the fixed-order slot reduction with its stores into every rank's result
block against the rank-ordered sum of P plain lifts.  The concurrent barrier between real ranks is covered by tests/multigpu_worker.py.  Run in a
subprocess by tests/test_gpu_kernels.py (a wait that times out traps the kernel and would poison the caller's context)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hippyflow_b200 import _lib as K          # noqa: E402
from hippyflow_b200.peer import PeerExchange  # noqa: E402


def run_case(P, n, R, ncols, dev):
    g = torch.Generator(device="cpu").manual_seed(1000 * P + ncols)
    Xs = [K.to_padded(torch.randn((R, n), generator=g, dtype=torch.float64), dev) for _ in range(P)]
    Ws = [K.to_padded(torch.randn((R, ncols), generator=g, dtype=torch.float64), dev) for _ in range(P)]
    alpha = 0.37
    ref = None
    for X, W in zip(Xs, Ws):                       # rank order, like the owner's fixed-order sum
        part = K.dgemm(K.HFB_TN, X, W, alpha=alpha)
        ref = part.clone() if ref is None else ref + part
    ld = ((ncols + 15) // 16) * 16
    group = PeerExchange.local_group(P, dev, n, ld, ncols)
    views = {}
    for rep in range(3):                           # later exchanges reuse the slots and flags (growing epochs), alternate blocks
        for r, ex in enumerate(group):
            ex.parity ^= 1
            ex.result_view().fill_(float("nan"))
            ex.push(Xs[r], Ws[r], alpha)
        for phase in (PeerExchange.reduce_bcast, None):
            for ex in group:
                ex.signal()
            for ex in group:
                ex.wait()
            if phase is not None:
                for ex in group:
                    phase(ex)
        torch.cuda.synchronize()
        for r, ex in enumerate(group):
            Y = ex.result_view()
            assert Y.shape == (n, ncols) and Y.stride(0) == ld
            assert torch.equal(Y, ref), (P, n, ncols, rep, r, float((Y - ref).abs().max()))
            views[(rep, r)] = Y
        if rep:                                    # the previous exchange's block is still intact
            for r in range(P):
                assert torch.equal(views[(rep - 1, r)], ref)
    del views, Y
    for ex in group:
        ex.close()


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    before = K.launch_count()
    for P, n, R, ncols in [(2, 5041, 64, 27), (3, 66049, 96, 67), (8, 40000, 48, 138), (4, 1000, 32, 266), (16, 3000, 16, 32)]:
        run_case(P, n, R, ncols, dev)
        print("case P=%d n=%d cols=%d ok" % (P, n, ncols))
    print("PEER_SELFTEST_OK launches=%d" % (K.launch_count() - before))


if __name__ == "__main__":
    main()
