#!/bin/bash
# One short GPU-box pass for the run-staged SpMM: correctness tests of the new kernel, then the timing table against the ring
# kernel at the benchmark shapes.  Outputs under gpurun_out/.  Usage: gpurun --timeout 400 -- 'bash tools/gpu_runs_round.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "runs" > gpurun_out/r3_runs_tests.log 2>&1
stamp "runs tests rc=$?: $(tail -1 gpurun_out/r3_runs_tests.log)"
for impls in ${RUNS_IMPLS:-frag,ring,runs}; do
    HFB_CHECK_CAPS=16,32 timeout 150 python tools/check_spmm.py --quick --impls $impls --json gpurun_out/r3_spmm_$impls.json \
        > gpurun_out/r3_spmm_$impls.log 2>&1
    stamp "table $impls rc=$?"
    grep -E "GB/s|Error" gpurun_out/r3_spmm_$impls.log | cut -c1-150
done
