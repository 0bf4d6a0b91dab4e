#!/bin/bash
# Round 2, trip 6b (4 GPUs of one box; 8-GPU leases were not available): 4-GPU process test, strong scaling N = 2 / 4 with the
# parity check against the C2 fingerprint and the new N > 1 e2e, subscriber filter on / off at N = 4.
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip6b.log 2>&1
nvidia-smi -L | head -8; python -c "import bench; print('host_threads', bench.host_threads())"; free -g | head -2
echo "== 4-GPU sharded test"
timeout 250 python -m pytest tests/test_sharded_gpu.py -x -q -k "multi_gpu and 4" 2>&1 | tail -4
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("N=%d ms/step %.2f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), "parity", d["parity"]["c2"].get("green"), d["parity"]["c2"].get("equal")); print("  e2e", d.get("e2e")); print("  rank0", d["per_iter"][0]["iter_ms"], d["per_iter"][0]["modes"]); print("  slowest rank per iteration", [max(r["iter_ms"][i] for r in d["per_iter"]) for i in range(len(d["per_iter"][0]["iter_ms"]))]); print("  roofline", {k: d["roofline"][k] for k in ("kernel","avg_launch_ms","frac")} if d["roofline"] else None, "stage_ms", d["run"]["stage_ms"])'
for n in 4 2; do
  echo; echo "== N=$n p2p (default)"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/r2_n$n.err | tee gpurun_out/r02_scale_n$n.json | python -c "$pick" || tail -8 gpurun_out/r2_n$n.err
done
echo; echo "== N=4 p2p, publish to all peers (no subscriber filter)"
SB200_PUBLISH_ALL=1 timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e 2> gpurun_out/r2_n4b.err | python -c "$pick" || tail -8 gpurun_out/r2_n4b.err
echo; echo "== N=4 NCCL fallback"
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 5 --warmup 3 --no-p2p --no-e2e 2> gpurun_out/r2_n4c.err | python -c "$pick" || tail -8 gpurun_out/r2_n4c.err
