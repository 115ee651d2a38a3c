"""Single-GPU self-test of the NVLink sketch exchange (hippyflow_b200/peer.py): P emulated ranks inside one process, every
'peer' buffer local.  The phases run rank by rank on ONE stream (push of all ranks, signal of all ranks, wait of all ranks,
reduce, ...), so no kernel ever waits for work that is queued behind it.  Exercises the peer-store epilogue of the lift GEMM
(hfb_dgemm_peer), the flag words and epochs, the fixed-order slot reduction and the gather kernel against the rank-ordered
sum of P plain lifts.  The concurrent barrier between real ranks is covered by tests/multigpu_worker.py.  Run in a
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
    Ys = [K.padded_zeros(n, ncols, dev) for _ in range(P)]
    for rep in range(2):                           # the second exchange reuses the slots and flags (growing epochs)
        for r in range(P):
            Ys[r].zero_()
            group[r].push(Xs[r], Ws[r], alpha)
        for phase in (PeerExchange.reduce, PeerExchange.gather):
            for ex in group:
                ex.signal()
            for ex in group:
                ex.wait()
            for r, ex in enumerate(group):
                phase(ex, Ys[r])
        torch.cuda.synchronize()
        for r in range(P):
            assert torch.equal(Ys[r], ref), (P, n, ncols, rep, r, float((Ys[r] - ref).abs().max()))
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
