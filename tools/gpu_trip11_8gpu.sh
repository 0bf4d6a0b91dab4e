#!/bin/bash
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
  for n in 8 4; do
    echo "== bench --gpus $n full size (fused p2p)"
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) \
        bench.py --gpus $n --steps 3 --warmup 3 2>&1 | grep '^{' > gpurun_out/r01_scale_n$n.json
    python - <<PY
import json
d=json.load(open("gpurun_out/r01_scale_n$n.json"))
print('N=$n value',d['value'],'ms/step',d['ms_per_step'],'iters',d['config']['iterations_per_step'],'stage_ms',d['config']['stage_ms'])
for r in d['per_iter']: print('   rank',r['rank'],'rows',r['rows'],'edges',r['edges_local'],'iter_ms',r['iter_ms'],'modes',r['modes'])
PY
  done
} > gpurun_out/trip11.log 2>&1
tail -60 gpurun_out/trip11.log | cut -c1-600
