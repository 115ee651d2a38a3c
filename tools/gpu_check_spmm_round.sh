#!/bin/bash
# One GPU-box pass: SpMM kernel tests -> SpMM sweep (picks the fastest variant/caps at cfg2) -> full GPU suite, smoke and
# bench under that choice -> ncu of the chosen SpMM kernel and the bench launch list.  Outputs under gpurun_out/.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check_spmm_round.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }

timeout 400 python -m pytest tests/test_gpu_kernels.py -k spmm -x -q > gpurun_out/t_spmm.log 2>&1
SPMM_RC=$?
stamp "spmm kernel tests rc=$SPMM_RC: $(tail -1 gpurun_out/t_spmm.log)"

IMPLS=staged,regblock
[ $SPMM_RC -ne 0 ] && IMPLS=staged
timeout 500 python tools/check_spmm.py --impls $IMPLS --json gpurun_out/spmm_sweep.json > gpurun_out/spmm_sweep.log 2>&1
stamp "sweep rc=$?"
python - > gpurun_out/spmm_choice.env <<'PY'
import json
try:
    res = [r for r in json.load(open("gpurun_out/spmm_sweep.json")) if r["n"] == 263169 and r["max_abs_err"] < 1e-12]
    best = min(res, key=lambda r: r["ms"])
    impl, _, dep = best["impl"].partition("/depth")
    print("export HFB_SPMM_IMPL=%s HFB_SPMM_ROWS=%d HFB_SPMM_COLS=%d" % (impl, best["caps"][0], best["caps"][1]))
    if dep:
        print("export HFB_SPMM_RB_DEPTH=%s" % dep)
    print("# best: %r" % (best,))
except Exception as e:
    print("# no choice: %r" % (e,))
PY
cat gpurun_out/spmm_choice.env
source gpurun_out/spmm_choice.env

timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1
stamp "gpu suite rc=$?: $(tail -1 gpurun_out/t_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?: $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
stamp "bench rc=$?"
head -c 600 gpurun_out/bench_1gpu.json; echo

KREGEX=csr_spmm_staged
[ "$HFB_SPMM_IMPL" = regblock ] && KREGEX=csr_spmm_regblock
HFB_CHECK_CAPS="$HFB_SPMM_ROWS,$HFB_SPMM_COLS" timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 \
    -f -o gpurun_out/spmm_chosen python tools/check_spmm.py --quick --impls ${HFB_SPMM_IMPL:-staged} > gpurun_out/ncu_spmm.log 2>&1
stamp "ncu spmm rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
stamp "ncu launch list rc=$?"
