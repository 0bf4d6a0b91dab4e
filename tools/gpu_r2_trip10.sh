#!/bin/bash
# Round 2, trip 10 (8 GPUs, one short call): exchange variants at N = 8 in one process (bench.py --sweep).
N=${1:-8}
cd /root/repo
mkdir -p gpurun_out
timeout ${2:-80} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --sweep > gpurun_out/r2_sweep_n$N.json 2> gpurun_out/r2_sweep_n$N.err
echo "rc=$?"; grep "^\[sweep\]" gpurun_out/r2_sweep_n$N.err | cut -c1-600
tail -3 gpurun_out/r2_sweep_n$N.err | cut -c1-300
