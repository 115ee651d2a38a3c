"""Measure the FP64 roofline denominators on the box: cuBLAS DGEMM (torch.matmul fp64) burst and
sustained, and a device copy for HBM. Writes gpurun_out/fp64_peaks.json. Same method the driver
used for MEASURED_PEAKS.json (bf16), applied to fp64."""
import json, time, subprocess, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {"gpu": torch.cuda.get_device_name(0)}
def clocks():
    try:
        o = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], text=True)
        return o.strip().splitlines()[0]
    except Exception as e:
        return str(e)
for N in (4096, 8192):
    a = torch.randn(N, N, dtype=torch.float64, device=dev); b = torch.randn(N, N, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(3): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{N}_burst_tflops"] = 2 * N**3 / best * 1e-9
    if N == 8192:
        t0 = time.time(); n = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 4.0:
            for _ in range(5): torch.matmul(a, b, out=c)
            n += 5; torch.cuda.synchronize()
        clk = clocks()
        e1.record(); torch.cuda.synchronize()
        res["dgemm_8192_sustained_tflops"] = 2 * N**3 * n / e0.elapsed_time(e1) * 1e-9
        res["clocks_under_dgemm"] = clk
# tall-skinny shapes of the path (cfg2): W = Xt[4096 x 263169] @ B[263169 x 266];  Y = Xt^T @ W
n, R, m = 263169, 4096, 266
Xt = torch.randn(R, n, dtype=torch.float64, device=dev); B = torch.randn(n, m, dtype=torch.float64, device=dev)
for name, f, fl in (("xtb", lambda: Xt @ B, 2.0 * n * R * m), ("xw", None, 2.0 * n * R * m)):
    if name == "xw":
        W = Xt @ B; f = lambda: Xt.t() @ W
    for _ in range(2): f()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res[f"cublas_tallskinny_{name}_tflops"] = fl / best * 1e-9
    res[f"cublas_tallskinny_{name}_ms"] = best
del Xt, B
x = torch.empty(1 << 30, dtype=torch.float64, device=dev); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
res["hbm_copy_gbs"] = 2 * x.numel() * 8 / best * 1e-6
print(json.dumps(res, indent=1))
import os; os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
