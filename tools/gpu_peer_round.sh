#!/bin/bash
# One multi-GPU gpurun call: emulated-rank self-test, NCCL/peer parity worker, lift+exchange timings, bench line (with targets).
W=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 200 python tools/peer_selftest.py > gpurun_out/peer_selftest.log 2>&1; echo "selftest rc=$?"; tail -2 gpurun_out/peer_selftest.log
timeout 600 $TR --master-port 29511 tests/multigpu_worker.py > gpurun_out/peer_mgw$W.log 2>&1; echo "worker rc=$?"
tail -3 gpurun_out/peer_mgw$W.log
timeout 400 $TR --master-port 29512 tools/measure_peer_lift.py > gpurun_out/peer_lift$W.log 2>&1; echo "lift rc=$?"
grep PEER_LIFT gpurun_out/peer_lift$W.log || tail -20 gpurun_out/peer_lift$W.log
timeout 900 $TR --master-port 29513 bench.py --gpus $W --steps 20 --warmup 3 $2 > gpurun_out/peer_bench${W}.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/peer_bench${W}.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], l['value'], l.get('exchange'), l.get('sharded_parity')); print(json.dumps(l.get('targets'))[:1500]); print(l['e2e'])" || tail -30 gpurun_out/peer_bench${W}.log
