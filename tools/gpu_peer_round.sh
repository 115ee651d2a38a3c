#!/bin/bash
# One gpurun call (2+ GPUs): emulated-rank self-test, NCCL/peer parity worker, lift+exchange timings, bench with both routes.
W=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/peer_topo.txt 2>&1
timeout 300 python tools/peer_selftest.py > gpurun_out/peer_selftest.log 2>&1; echo "selftest rc=$?"
tail -3 gpurun_out/peer_selftest.log
timeout 600 $TR --master-port 29511 tests/multigpu_worker.py > gpurun_out/peer_mgw$W.log 2>&1; echo "worker rc=$?"
tail -5 gpurun_out/peer_mgw$W.log
timeout 400 $TR --master-port 29512 tools/measure_peer_lift.py > gpurun_out/peer_lift$W.log 2>&1; echo "lift rc=$?"
grep PEER_LIFT gpurun_out/peer_lift$W.log || tail -20 gpurun_out/peer_lift$W.log
for chunks in 4 1; do
HFB_PEER_CHUNKS=$chunks timeout 400 $TR --master-port 29513 bench.py --gpus $W --steps 10 --warmup 3 --no-e2e --no-targets > gpurun_out/peer_bench${W}_peer$chunks.log 2>&1; echo "bench peer$chunks rc=$?"
tail -1 gpurun_out/peer_bench${W}_peer$chunks.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], l['value'], l.get('exchange',{}).get('route'), l.get('sharded_parity'))" || tail -20 gpurun_out/peer_bench${W}_peer$chunks.log
done
HFB_PEER_LIFT=0 timeout 400 $TR --master-port 29514 bench.py --gpus $W --steps 10 --warmup 3 --no-e2e --no-targets > gpurun_out/peer_bench${W}_nccl.log 2>&1; echo "bench nccl rc=$?"
tail -1 gpurun_out/peer_bench${W}_nccl.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], l['value'], l.get('exchange',{}).get('route'))" || tail -20 gpurun_out/peer_bench${W}_nccl.log
