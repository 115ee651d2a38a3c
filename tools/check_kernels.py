"""Quick on-GPU check + timing of the C-ABI kernels against torch fp64 (development aid; the parity
tests proper live in tests/)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from hippyflow_b200 import _lib as K

dev = torch.device("cuda:0")
torch.manual_seed(0)

def rel(a, b):
    return float((a - b).norm() / b.norm())

def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

ok = True
# ---- correctness over odd shapes, all layouts
for layout, name in ((K.HFB_NN, "NN"), (K.HFB_TN, "TN"), (K.HFB_NT, "NT")):
    for (M, N, Kd) in ((100, 25, 289), (289, 25, 100), (513, 266, 1000), (130, 138, 77), (64, 74, 6400), (1000, 210, 333), (257, 4, 50), (300, 300, 300)):
        for splits in (0, 1, 3):
            A = torch.randn(M, Kd, dtype=torch.float64, device=dev)
            B = torch.randn(Kd, N, dtype=torch.float64, device=dev)
            ref = A @ B
            Ad = K.to_padded(A.t().contiguous() if layout == K.HFB_TN else A, dev)
            Bd = K.to_padded(B.t().contiguous() if layout == K.HFB_NT else B, dev)
            C = K.dgemm(layout, Ad, Bd, alpha=0.5, splits=splits)
            torch.cuda.synchronize()
            e = rel(C, 0.5 * ref)
            good = e < 1e-13
            ok &= good
            if not good or splits == 0:
                print(f"dgemm {name} M={M} N={N} K={Kd} splits={splits}: rel err {e:.2e} {'OK' if good else 'FAIL'}")
print("dgemm correctness:", "PASS" if ok else "FAIL")

# ---- perf at the cfg2 shapes
if "--perf" in sys.argv:
    n, R, m = 263169, 4096, 266
    Xt = K.padded_empty(R, n, dev); Xt.normal_()
    B = K.padded_empty(n, m, dev); B.normal_()
    W = K.padded_empty(R, m, dev)
    Y = K.padded_empty(n, m, dev)
    fl = 2.0 * n * R * m
    res = {}
    for s in (0,):
        t = timeit(lambda: K.dgemm(K.HFB_NN, Xt, B, out=W, splits=s))
        res[f"NN_W=XtB_splits{s}"] = (t, fl / t * 1e-9)
    t = timeit(lambda: K.dgemm(K.HFB_TN, Xt, W, out=Y))
    res["TN_Y=XW"] = (t, fl / t * 1e-9)
    ref = Xt @ B
    print("perf-shape err NN", rel(W, ref)); 
    K.dgemm(K.HFB_TN, Xt, W, out=Y); print("perf-shape err TN", rel(Y, Xt.t() @ W))
    for k, v in res.items(): print(f"{k}: {v[0]:.3f} ms  {v[1]:.2f} TFLOP/s")
    print("auto splits NN:", K.lib().hfb_dgemm_auto_splits(0, R, m, n))
    # Gram TN: G = Q^T Z (m x m, K = n)
    t = timeit(lambda: K.dgemm(K.HFB_TN, B, Y, splits=0)); print(f"TN Gram m x m K=n: {t:.3f} ms {2.0*n*m*m/t*1e-9:.2f} TF/s")
    # U = Q V (n x m @ m x 256)
    V = K.padded_empty(m, 256, dev); V.normal_()
    t = timeit(lambda: K.dgemm(K.HFB_NN, Y, V)); print(f"NN U=QV: {t:.3f} ms {2.0*n*m*256/t*1e-9:.2f} TF/s")
    json.dump(res, open("gpurun_out/check_perf.json", "w"))
sys.exit(0 if ok else 1)
