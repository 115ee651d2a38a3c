"""Tuning sweep of the run-staged SpMM (csr_spmm_runs_kernel): ring depth at the benchmark
shapes, each with one clock64-profiled launch.  Usage (GPU box): python tools/tune_runs.py [--json out.json]"""
import json
import os
import sys
sys.path.insert(0, ".")
import torch
from hippyflow_b200 import _lib as K, synthetic as syn
from hippyflow_b200.linalg import CsrMatrix

dev = torch.device("cuda:0")
PEAK = 6552.3
res = []


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for n, m in ((263169, 266), (251001, 138), (263169, 74)):
    M = syn.p1_mass_matrix_for(n)
    Md = CsrMatrix(M, dev)
    B = K.padded_empty(n, m, dev).normal_()
    C = K.padded_empty(n, m, dev)
    ref = torch.sparse_csr_tensor(Md.rowptr.long(), Md.colind.long(), Md.val, size=M.shape) @ B.contiguous()
    rplan = CsrMatrix._runs_blobs(Md.plan, dev)
    for ng in (1,):
        for slots in (8, 4, 2):
            os.environ["HFB_RUNS_SLOTS"] = str(slots)
            os.environ.pop("HFB_RUNS_PROF", None)
            if K.csr_spmm_runs_slots(rplan, m, K._ld(B)) == 0:
                continue
            C.zero_()
            t = timeit(lambda: K.csr_spmm_runs(rplan, B, C))
            err = float((C - ref).abs().max())
            by = Md.spmm_bytes(m)
            print(f"m={m}  slots<={slots}: {t:.4f} ms {by / t / 1e6:.0f} GB/s ({by / t / 1e6 / PEAK * 100:.1f}%) err {err:.1e}", flush=True)
            res.append({"n": n, "m": m, "max_slots": slots, "ms": t, "gbs": by / t / 1e6,
                        "frac_hbm": by / t / 1e6 / PEAK, "err": err})
            os.environ["HFB_RUNS_PROF"] = "1"
            K.csr_spmm_runs(rplan, B, C)
            torch.cuda.synchronize()
if "--json" in sys.argv:
    json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
