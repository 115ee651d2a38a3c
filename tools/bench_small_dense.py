"""Timing of the round-2 kernels on one B200 (CUDA events, best of n): device Cholesky-QR factor, batched Jacobi SVD, the
strided-batch DMMA GEMM on the per-sample products of cfg4 (dQ = 200, dM = 251,001, r = 128) beside the stacked GEMM on the
same data, the batched randomized SVD of cfg3-shaped Jacobians and the output-subspace forms.
Writes gpurun_out/<tag>_small_dense.json.  Development / reporting aid; bench.py is the contract benchmark."""
import json
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
only = sys.argv[2] if len(sys.argv) > 2 else "all"          # "chol": just the Cholesky-QR factor timings
dev = torch.device("cuda:0")
res = {}


def timed(f, n=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ---- device Cholesky-QR factor
for m in (138, 210, 266, 522, 1024):
    Y = torch.randn(4 * m, m, dtype=torch.float64, device=dev)
    G = K.to_padded(Y.t() @ Y, dev)
    res["chol_inverse_m%d_ms" % m] = timed(lambda: K.chol_inverse(G), n=10)
    print("chol_inverse m=%d: %.3f ms" % (m, res["chol_inverse_m%d_ms" % m]), flush=True)
    if m in (138, 266):
        K.chol_inverse_profile(G)
        cyc = K.chol_inverse_profile(G)
        names = ["scaling", "copy", "panel load", "diag block", "panel solve", "trailing update", "zero S", "inv panel load",
                 "inv solve+stage", "inv eager update", "final scaling"]
        res["chol_inverse_m%d_section_kcycles" % m] = {nm: float(c) / 1e3 for nm, c in zip(names, cyc)}
        print("   sections (k cycles):", res["chol_inverse_m%d_section_kcycles" % m], flush=True)

if only == "chol":
    json.dump(res, open("gpurun_out/%s_small_dense_chol.json" % tag, "w"), indent=1)
    sys.exit(0)

# ---- batched Jacobi SVD
for (batch, rows, cols) in ((592, 100, 60), (592, 100, 110), (296, 200, 138), (592, 138, 138)):
    A0 = torch.randn(batch, rows, cols, dtype=torch.float64, device=dev) * (torch.arange(1, cols + 1, device=dev, dtype=torch.float64) ** -1.0)
    A = K.batched_empty(batch, rows, cols, dev)

    def run():
        A.copy_(A0)
        return K.jacobi_svd_batched_(A)
    t_copy = timed(lambda: A.copy_(A0))
    t = timed(run) - t_copy
    sig, info = run()
    res["jacobi_%dx%dx%d" % (batch, rows, cols)] = {"ms": t, "us_per_matrix_per_sm": t * 1e3 / (batch / 148.0), "sweeps_mean": float(info.float().mean())}
    print("jacobi", (batch, rows, cols), res["jacobi_%dx%dx%d" % (batch, rows, cols)], flush=True)

# ---- cfg4 per-sample products: strided-batch launch vs the stacked GEMM on the same resident chunk
N, dQ, dM, r = 32, 200, 251001, 128
J = torch.empty((N * dQ, dM + (dM % 2)), dtype=torch.float64, device=dev)[:, :dM]
K.fill_random_(J, 5)
J3 = J.as_strided((N, dQ, dM), (dQ * J.stride(0), J.stride(0), 1))
MPhi = K.padded_empty(dQ, r, dev).normal_()
V = K.padded_empty(dM, r, dev).normal_()
out = K.batched_empty(N, dM, r, dev)
t_b = timed(lambda: K.dgemm_batched(K.HFB_TN, J3, MPhi, out=out), n=5)
fl = 2.0 * N * dQ * dM * r
JV = K.padded_empty(N * dQ, r, dev)
t_s = timed(lambda: K.dgemm(K.HFB_NN, J, V, out=JV), n=5)
res["cfg4_jacobian_transpose_action"] = {"samples": N, "dQ": dQ, "dM": dM, "rQ": r, "ms": t_b, "tflops": fl / t_b * 1e-9,
                                         "launches": 1, "bytes_GB": (N * dQ * dM + N * dM * r) * 8 / 1e9,
                                         "GBps": (N * dQ * dM + N * dM * r) * 8 / t_b * 1e-6}
res["cfg4_stacked_gemm_same_data"] = {"what": "J_all V (N dQ x dM)(dM x r), one stacked NN GEMM", "ms": t_s, "tflops": fl / t_s * 1e-9}
res["cfg4_batched_over_stacked"] = (fl / t_b) / (fl / t_s)
print("cfg4 JstarPhi batched:", res["cfg4_jacobian_transpose_action"], "stacked:", res["cfg4_stacked_gemm_same_data"], flush=True)
# the round-1 form for comparison: one launch per sample
def loop():
    for i in range(N):
        K.dgemm(K.HFB_TN, J[i * dQ:(i + 1) * dQ], MPhi, out=out[i])
t_l = timed(loop, n=3)
res["cfg4_jacobian_transpose_action_per_sample_loop"] = {"ms": t_l, "tflops": fl / t_l * 1e-9, "launches": N}
print("per-sample loop:", res["cfg4_jacobian_transpose_action_per_sample_loop"], flush=True)
Phi = K.padded_empty(dQ, r, dev).normal_()
t_r = timed(lambda: hf.reduced_jacobians(J3, Phi, V, dev), n=3)
res["cfg4_reduced_jacobians"] = {"ms": t_r, "samples_per_s": N / (t_r * 1e-3), "tflops": fl / t_r * 1e-9}
print("reduced jacobians:", res["cfg4_reduced_jacobians"], flush=True)
del J, J3, out, JV
torch.cuda.empty_cache()

# ---- cfg3-shaped Jacobians: batched randomized SVD and the output subspace
N, dQ, dM, k, l = 256, 100, 65536, 50, 60
J = torch.empty((N * dQ, dM), dtype=torch.float64, device=dev)
K.fill_random_(J, 31)
K.colscale_(J, (1.0 + torch.arange(dM, device=dev, dtype=torch.float64)) ** -0.5)
J3 = J.view(N, dQ, dM)
Om = K.padded_empty(dM, l, dev).normal_()
t = timed(lambda: hf.accuracyEnhancedSVD_batched(J3, Om, k, s=1, device=dev), n=3, warm=1)
fl = N * (2.0 * dQ * dM * l * 4 + 2.0 * dM * l * l + 2.0 * dM * l * k)      # J Om, J^T Y, J Z, J^T Q, B B^T, B^T W
res["cfg3_randomized_svd_batched"] = {"samples": N, "dQ": dQ, "dM": dM, "rank": k, "sketch": l, "power_iterations": 1, "ms": t,
                                      "samples_per_s": N / (t * 1e-3), "tflops": fl / t * 1e-9,
                                      "J_read_passes": 4, "GBps_J": 4 * N * dQ * dM * 8 / t * 1e-6}
print("cfg3 rSVD:", res["cfg3_randomized_svd_batched"], flush=True)
op = hf.MeanJJTfromDataOperator(J3, device=dev)
t_d = timed(lambda: op.dense(), n=5)
res["cfg3_output_dense_JJT"] = {"ms": t_d, "tflops": 2.0 * N * dQ * dQ * dM / t_d * 1e-9, "GBps_J": N * dQ * dM * 8 / t_d * 1e-6}
X = hf.DeviceMultiVector(dQ, 60, device=dev)
X.tensor().normal_()
Y = hf.DeviceMultiVector(dQ, 60, device=dev)
t_o = timed(lambda: op.matMvMult(X, Y), n=3)
res["cfg3_output_operator_form"] = {"ms": t_o, "tflops": 4.0 * N * dQ * dM * 60 / t_o * 1e-9}
print("output subspace:", res["cfg3_output_dense_JJT"], res["cfg3_output_operator_form"], flush=True)
json.dump(res, open("gpurun_out/%s_small_dense.json" % tag, "w"), indent=1)
