#!/bin/bash
# 2-GPU trip: sharded parity (NCCL broadcast exchange and fused P2P exchange), then the multi-GPU bench
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=index,name --format=csv,noheader
  echo "== pytest sharded"
  timeout 400 python -m pytest tests/test_sharded_gpu.py -x -q --timeout 180 2>&1 | tail -15
  for mode in "" "--no-p2p"; do
    echo "== bench --gpus 2 (100M edges) $mode"
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --nodes 5000000 --edges 100000000 --scale 23 --steps 3 --warmup 3 $mode 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('value',d['value'],'ms/step',d['ms_per_step'],'iters',d['config']['iterations_per_step'],d['config']['parallelism'][:60])"
  done
  echo "== bench --gpus 2 full size (fused p2p)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -3 | cut -c1-1500
} > gpurun_out/trip8.log 2>&1
tail -40 gpurun_out/trip8.log | cut -c1-1500
