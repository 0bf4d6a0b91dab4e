#!/bin/bash
# Round 2, trip 8 (2 GPUs): short-row kernel on a side stream beside the long-row kernel (SB200_QUAD_SIDE_CTAS = CTAs per SM
# of the persistent k_pull_quad_owned; 0 = the serial layout of trips 3 / 6b), per-kernel times at N = 2.
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip8.log 2>&1
nvidia-smi -L | head -4
echo "== 2-GPU sharded tests"
timeout 400 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -4
pick='import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1]); print("N=%d ms/step %.2f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), "parity", d["parity"]["c2"].get("green")); print("  rank0", d["per_iter"][0]["iter_ms"], d["per_iter"][0]["modes"]); print("  kernels", [(k["kernel"], k["launches"], round(k["avg_launch_ms"],3)) for k in d["kernels"]])'
for c in 2 0 1 3 4; do
  echo; echo "== N=2 p2p, SB200_QUAD_SIDE_CTAS=$c"
  SB200_QUAD_SIDE_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$c bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2> gpurun_out/r2_t8_$c.err | tee gpurun_out/r2_t8_n2_c$c.json | python -c "$pick" || tail -8 gpurun_out/r2_t8_$c.err
done
