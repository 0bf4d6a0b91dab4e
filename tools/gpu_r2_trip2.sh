#!/bin/bash
# Second GPU trip of round 2 (2 GPUs): the multi-GPU pieces prepared at the end of round 1.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_r2_trip2.sh; tail -60 gpurun_out/r2_trip2.log'
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip2.log 2>&1
echo "== sharded parity: nccl, p2p, symmetric memory unicast, NVSwitch multicast"
SB200_TEST_SYMM=1 timeout 400 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -6
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-bm25 --no-cpu"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ms/step %.2f" % d["ms_per_step"], d["config"]["parallelism"][:60], [(r["rank"], r["iter_ms"], r["modes"]) for r in d.get("per_iter", [])][:1])'
for v in "" "SB200_SHARDED_PUSH=1"; do
  for x in p2p multicast; do
    echo; echo "== N=2 exchange=$x [$v]"; env $v timeout 200 $T --exchange $x 2> gpurun_out/r2_n2.err | python -c "$pick" || tail -5 gpurun_out/r2_n2.err
  done
done
