#!/bin/bash
# Round 2, trip 9 (N GPUs, default 2): sharded ingest (each rank copies 1/N of the edge stream, NCCL all-gather over NVLink) in the
# N > 1 e2e, persistent finalize + owned item list + side-stream short rows in the loop.  Usage: gpu_r2_trip9.sh [N]
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip9_n$N.log 2>&1
nvidia-smi -L | head -8
if [ "$N" -le 4 ]; then
echo "== sharded tests"
timeout 400 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -4
fi
pick='import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1]); print("N=%d ms/step %.2f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), "parity", d["parity"]["c2"].get("green"), d["parity"]["c2"].get("equal")); print("  e2e", d.get("e2e")); print("  rank0", d["per_iter"][0]["iter_ms"], d["per_iter"][0]["modes"]); print("  kernels", [(k["kernel"], k["launches"], round(k["avg_launch_ms"],3)) for k in d["kernels"]]); print("  stage_ms", d["run"]["stage_ms"])'
run() { # name, env..., extra args
  name=$1; shift
  echo; echo "== N=$N $name"
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA 2> gpurun_out/r2_t9_$N.err | tee gpurun_out/r2_t9_n${N}_$name.json | python -c "$pick" || tail -12 gpurun_out/r2_t9_$N.err
}
EXTRA=""          run default SB200_X=0
if [ "$N" -le 4 ]; then
EXTRA="--no-e2e"  run side0 SB200_QUAD_SIDE_CTAS=0
EXTRA="--no-e2e"  run side4 SB200_QUAD_SIDE_CTAS=4
fi
if [ "$N" -le 2 ]; then
EXTRA="--no-e2e"  run nolist SB200_OWNED_ITEMS=0
fi
cp gpurun_out/r2_t9_n${N}_default.json gpurun_out/r02_scale_n$N.json
