"""Throughput of the other BASELINE.json configs on ONE B200 (device-resident synthetic data):
  cfg3  active subspace from stored Jacobians  (N x 100 x 65,536, rank 200 + 10, doublePass)
  cfg4  KLE from stored draws + projection of the draws onto the basis (n = 251,001 = 501^2, N = 16,384, rank 128 + 10)
  spmm  CSR SpMM at cfg2 / cfg5 widths (HBM roofline)
Writes gpurun_out/configs_r01.json.  Development / reporting aid; bench.py is the contract benchmark."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import hippyflow_b200 as hf
from hippyflow_b200 import _lib as K, synthetic as syn
from hippyflow_b200.linalg import CsrMatrix

dev = torch.device("cuda:0")
res = {}
HBM = 6552.3


def timed(f, n=3, warm=1):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)

which = sys.argv[1:] or ["spmm", "cfg4", "cfg3"]

if "spmm" in which:
    for name, n, m in (("cfg2", 263169, 266), ("cfg5", 1002001, 266), ("cfg4", 251001, 138)):
        M = syn.p1_mass_matrix_for(n)
        Md = CsrMatrix(M, dev)
        B = K.padded_empty(n, m, dev).normal_()
        C = K.padded_empty(n, m, dev)
        t = timed(lambda: Md.matmat(B, out=C), n=10, warm=3)
        by = Md.spmm_bytes(m)
        res["spmm_" + name] = {"n": n, "m": m, "nnz": int(M.nnz), "ms": t * 1e3, "algorithmic_GB": by / 1e9, "GBps": by / t / 1e9,
                               "frac_of_hbm_peak": by / t / 1e9 / HBM}
        print("spmm", name, res["spmm_" + name], flush=True)
        del B, C, Md

if "cfg4" in which:
    n, N, k, p = 251001, 16384, 128, 10
    M = syn.p1_mass_matrix_for(n)
    m_data = syn.snapshots_device(n, N, dev, r0=512, decay=1.0, eps=1e-6, seed=21)
    params = hf.KLEParameterList(); params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = k, p, False, False
    proj = hf.KLEProjector(hf.SampleCovariancePrior(m_data, M, device=dev), parameters=params)
    out = {}
    def solve():
        out["r"] = proj.construct_input_subspace("mass")
    t = timed(solve)
    m = k + p
    fl = 6.0 * n * N * m + 2.0 * N * m * m
    d, V, E = out["r"]
    res["cfg4_kle"] = {"n": n, "N": N, "rank": k, "ms": t * 1e3, "tflops_executed": fl / t * 1e-12, "d_head": [float(x) for x in d[:3]]}
    print("cfg4 kle", res["cfg4_kle"], flush=True)
    red = K.padded_empty(N, k, dev)
    tp = timed(lambda: hf.project_data(m_data, E, dev, out=red), n=5)
    by = N * n * 8 + n * k * 8 + N * k * 8
    res["cfg4_projection"] = {"ms": tp * 1e3, "tflops": 2.0 * N * n * k / tp * 1e-12, "algorithmic_GB": by / 1e9, "GBps": by / tp / 1e9,
                              "frac_of_hbm_peak": by / tp / 1e9 / HBM}
    print("cfg4 projection", res["cfg4_projection"], flush=True)
    # size-independent property at full size: V^T (M V) = I and the projection of the basis onto itself is I
    G = K.dgemm(K.HFB_TN, V.tensor(), E.tensor()).cpu().numpy()
    res["cfg4_orth_err"] = float(np.abs(G - np.eye(k)).max())
    # reduced Jacobians Phi^T J_i V on a resident chunk of 64 samples: J (64, 200, 251001)
    dQ, chunk, rQ = 200, 64, 128
    J = torch.empty((chunk * dQ, n + (n % 2)), dtype=torch.float64, device=dev)[:, :n]
    K.fill_random_(J, 5)
    Phi = K.padded_empty(dQ, rQ, dev).normal_()
    J3 = J.as_strided((chunk, dQ, n), (dQ * J.stride(0), J.stride(0), 1))
    tj = timed(lambda: hf.reduced_jacobians(J3, Phi, V, dev), n=3)
    byj = chunk * dQ * n * 8
    res["cfg4_reduced_jacobians"] = {"samples": chunk, "ms": tj * 1e3, "samples_per_s": chunk / tj, "tflops": 2.0 * chunk * dQ * n * k / tj * 1e-12,
                                     "GBps_J": byj / tj / 1e9}
    print("cfg4 reduced jacobians", res["cfg4_reduced_jacobians"], flush=True)
    del m_data, proj, J, J3, out, V, E
    torch.cuda.empty_cache()

if "cfg3" in which:
    N, dQ, dM, k, p = 2048, 100, 65536, 200, 10          # 107 GB on one GPU (the 4096-sample config is 215 GB: 8 GPUs)
    J = torch.empty((N * dQ, dM), dtype=torch.float64, device=dev)
    for i0 in range(0, N * dQ, 16384):
        K.fill_random_(J[i0:i0 + 16384], 31, row_offset=i0)
    # give the Jacobians a decaying spectrum: scale columns smoothly (keeps the run well conditioned)
    sc = (1.0 + torch.arange(dM, device=dev, dtype=torch.float64)) ** -0.5
    for i0 in range(0, N * dQ, 16384):
        K.colscale_(J[i0:i0 + 16384], sc)
    J3 = J.view(N, dQ, dM)
    params = hf.ActiveSubspaceParameterList(); params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = k, p, False, False
    proj = hf.ActiveSubspaceProjector(hf.StoredJacobians(J3), None, parameters=params, device=dev)
    out = {}
    def solve():
        out["r"] = proj.construct_input_subspace(prior_preconditioned=False)
    t = timed(solve, n=2)
    m = k + p
    fl = 6.0 * dM * N * dQ * m + 2.0 * N * dQ * m * m
    res["cfg3_as"] = {"N": N, "dQ": dQ, "dM": dM, "rank": k, "ms": t * 1e3, "tflops_executed": fl / t * 1e-12,
                      "d_head": [float(x) for x in out["r"][0][:3]]}
    print("cfg3", res["cfg3_as"], flush=True)

json.dump(res, open("gpurun_out/configs_r01.json", "w"), indent=1)
