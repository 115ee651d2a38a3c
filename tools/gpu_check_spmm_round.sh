#!/bin/bash
# One GPU-box pass (about 3 minutes): full GPU suite, smoke, bench, the SpMM variant table at the benchmark shapes and the
# ncu launch list of a short bench run.  Outputs under gpurun_out/.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_check_spmm_round.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }

timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1
stamp "gpu suite rc=$?: $(tail -1 gpurun_out/t_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?: $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
stamp "bench rc=$?"
head -c 400 gpurun_out/bench_1gpu.json; echo
HFB_CHECK_CAPS=16,32 timeout 300 python tools/check_spmm.py --impls staged,tma,regblock,dmma,frag,pipe --json gpurun_out/spmm_variants.json \
    > gpurun_out/spmm_variants.log 2>&1
stamp "variant table rc=$?"
grep "GB/s" gpurun_out/spmm_variants.log | cut -c1-130
HFB_CHECK_CAPS=8,24 timeout 200 python tools/check_spmm.py --impls frag,pipe --json gpurun_out/spmm_variants_8_24.json \
    > gpurun_out/spmm_variants_8_24.log 2>&1
grep "GB/s" gpurun_out/spmm_variants_8_24.log | cut -c1-130
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
stamp "ncu launch list rc=$?"
