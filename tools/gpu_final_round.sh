#!/bin/bash
# Round-end GPU pass (about 4 minutes): full GPU suite, smoke, the default bench line, one ncu --set full capture of the
# run-staged SpMM at cfg2 and the ncu launch list of a short bench run.  Outputs under gpurun_out/.
# Usage: gpurun --timeout 420 -- 'bash tools/gpu_final_round.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest_gpu.log 2>&1
stamp "gpu suite rc=$?: $(tail -1 gpurun_out/r3_pytest_gpu.log)"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1
stamp "smoke rc=$?: $(tail -1 gpurun_out/r3_smoke.log)"
timeout 150 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r3_bench_1gpu.json 2> gpurun_out/r3_bench_1gpu.err
stamp "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r3_bench_1gpu.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "hbm", d["roofline_hbm"]["kernel"], d["roofline_hbm"]["avg_launch_ms"], d["roofline_hbm"]["frac"],
          "e2e_ms", d["e2e"]["ms_per_step"], "parity", d.get("parity_full_size"))
except Exception as e:
    print("bench parse failed", e)
P
HFB_CHECK_CAPS=16,32 timeout 60 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_runs_kernel --launch-skip 5 -c 1 \
    -f -o gpurun_out/r3_spmm_runs_cfg2 python tools/check_spmm.py --quick --impls runs > gpurun_out/r3_ncu_runs.log 2>&1
stamp "ncu set full rc=$?"
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r3_bench_under_ncu.log 2>&1
stamp "ncu launch list rc=$?"
