#!/bin/bash
mkdir -p gpurun_out
{
  echo "== pytest sharded"
  timeout 300 python -m pytest tests/test_sharded_gpu.py -x -q --timeout 180 2>&1 | tail -4
  echo "== bench --gpus 2 full size (fused)"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | grep '^{' > gpurun_out/r01_scale_n2c.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r01_scale_n2c.json"))
print('N=2 value',d['value'],'ms/step',d['ms_per_step'])
for r in d['per_iter']: print('   rank',r['rank'],'iter_ms',r['iter_ms'])
PY
} > gpurun_out/trip15.log 2>&1
tail -12 gpurun_out/trip15.log | cut -c1-300
