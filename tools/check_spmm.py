"""SpMM micro-benchmark: the cluster-plan kernels ("dmma" DMMA panels, "frag" fragment records with whole-row or chunked
staging at 5 / 3 / 2 column groups per warp) and the generic panel kernel at the benchmark shapes, for a few cluster caps;
every result is checked against torch's sparse product.
Usage (GPU box): python tools/check_spmm.py [--quick] [--impls dmma,frag] [--json out.json]; HFB_CHECK_CAPS=16,32 pins the caps.
Operands (0.56-2.1 GB) exceed L2; times are the best of 20 launches (CUDA events)."""
import json
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from hippyflow_b200 import _lib as K, synthetic as syn
from hippyflow_b200.linalg import CsrMatrix

dev = torch.device("cuda:0")
PEAK = 6542.1   # MEASURED_PEAKS.json hbm_gbs
quick = "--quick" in sys.argv
shapes = ((263169, 266), (251001, 138), (263169, 74)) if quick else ((263169, 266), (1002001, 266), (251001, 138), (251001, 210), (263169, 74))
caps = ((16, 32),) if quick else ((16, 32), (12, 32), (8, 24))
if os.environ.get("HFB_CHECK_CAPS", "").replace(",", "").isdigit():
    caps = (tuple(int(v) for v in os.environ["HFB_CHECK_CAPS"].split(",")),)
impls = ("dmma", "frag", "ring", "runs")
if "--impls" in sys.argv:
    impls = tuple(sys.argv[sys.argv.index("--impls") + 1].split(","))
results = []


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


for n, m in shapes:
    M = syn.p1_mass_matrix_for(n)
    B = K.padded_empty(n, m, dev).normal_()
    C = K.padded_empty(n, m, dev)
    ref = None
    for rows, cols in caps:
        Md = CsrMatrix(M, dev, cluster_rows=False)
        Md.plan = CsrMatrix._build_plan(M.tocsr(), dev, max_rows=rows, max_cols=cols)
        Md.order = Md.plan["order"]
        if ref is None:
            ref = torch.sparse_csr_tensor(Md.rowptr.long(), Md.colind.long(), Md.val, size=M.shape) @ B.contiguous()
        for impl in impls:
            Md.impl = impl
            # tuning knobs swept per variant (None: leave the environment alone)
            knob = {"frag": "HFB_SPMM_FRAG_W"}.get(impl)
            depths = {"frag": (0, 192, 128, 96)}.get(impl, (None,))
            for dep in depths:
                if dep is not None:
                    os.environ.pop(knob, None)
                    if dep:
                        os.environ[knob] = str(dep)
                C.zero_()
                try:
                    t = timeit(lambda: Md.matmat(B, out=C))
                except Exception as e:   # e.g. shared-memory budget of the panel kernel at large caps
                    print(f"n={n} m={m} caps=({rows},{cols}) {impl}: {type(e).__name__}: {e}", flush=True)
                    continue
                by = Md.spmm_bytes(m)
                err = float((C - ref).abs().max())
                tag = impl + (f"/w{dep}" if dep else "")
                print(f"n={n} m={m} caps=({rows},{cols}) clusters={Md.plan['nclusters']} {tag}: {t:.3f} ms "
                      f"{by / t / 1e6:.0f} GB/s ({by / t / 1e6 / PEAK * 100:.1f}% of HBM peak) err {err:.2e}", flush=True)
                results.append({"n": n, "m": m, "caps": [rows, cols], "clusters": Md.plan["nclusters"], "impl": tag, "ms": t,
                                "gbs": by / t / 1e6, "frac_hbm": by / t / 1e6 / PEAK, "max_abs_err": err})
if "--json" in sys.argv:
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump(results, f, indent=1)
