#!/bin/bash
mkdir -p gpurun_out
{
  echo "== pytest sharded + path1"
  timeout 500 python -m pytest tests/test_sharded_gpu.py tests/test_hyperball_gpu.py -x -q --timeout 180 2>&1 | tail -8
  for mode in "" "--no-p2p"; do
  echo "== bench --gpus 2 full size $mode"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus 2 --steps 3 --warmup 3 $mode 2>&1 | grep '^{' > gpurun_out/r01_scale_n2b.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r01_scale_n2b.json"))
print('N=2 value',d['value'],'ms/step',d['ms_per_step'],'iters',d['config']['iterations_per_step'])
for r in d['per_iter']: print('   rank',r['rank'],'edges',r['edges_local'],'iter_ms',r['iter_ms'],'modes',r['modes'])
PY
  done
} > gpurun_out/trip12.log 2>&1
tail -30 gpurun_out/trip12.log | cut -c1-500
