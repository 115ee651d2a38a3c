import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from hippyflow_b200 import _lib as K, synthetic as syn
from hippyflow_b200.linalg import CsrMatrix
dev = torch.device("cuda:0")
for n, m in ((263169, 266), (1002001, 266), (251001, 138), (263169, 25)):
    M = syn.p1_mass_matrix_for(n); Md = CsrMatrix(M, dev)
    B = K.padded_empty(n, m, dev).normal_(); C = K.padded_empty(n, m, dev)
    for _ in range(3): Md.matmat(B, out=C)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); Md.matmat(B, out=C); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    by = Md.spmm_bytes(m)
    ref = torch.sparse_csr_tensor(Md.rowptr.long(), Md.colind.long(), Md.val, size=M.shape) @ B.contiguous()
    print(f"n={n} m={m}: {best:.3f} ms {by/best/1e6:.0f} GB/s ({by/best/1e6/6552.3*100:.1f}% of HBM peak) err {float((C-ref).abs().max()):.2e}", flush=True)
