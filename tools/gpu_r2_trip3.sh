#!/bin/bash
# Round 2, trip 3 (2 GPUs): sharded path behind the ABI (device barrier, subscriber-filtered publish), exchanges compared.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_trip3.sh; tail -100 gpurun_out/r2_trip3.log'
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip3.log 2>&1
echo "== host"; nproc; python -c "import os,bench; print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), 'host_threads', bench.host_threads())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2
nvidia-smi topo -m | head -6
echo "== sharded + group tests"
SB200_TEST_SYMM=1 timeout 500 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ms/step %.2f value %.3e" % (d["ms_per_step"], d["value"]), d["run"]["parallelism"][:70]); print("  parity", json.dumps(d["parity"])[:400]); print("  rank0", d["per_iter"][0]); print("  roofline", {k: d["roofline"][k] for k in ("kernel","avg_launch_ms","frac")} if d["roofline"] else None)'
for x in "--exchange p2p" "--exchange multicast" "--exchange symm" "--no-p2p"; do
  echo; echo "== N=2 $x"; timeout 300 $T $x 2> gpurun_out/r2_n2.err | python -c "$pick" || tail -8 gpurun_out/r2_n2.err
done
echo; echo "== N=2 p2p, publish to all peers (no subscriber filter)"; SB200_PUBLISH_ALL=1 timeout 300 $T --exchange p2p 2> gpurun_out/r2_n2.err | python -c "$pick" || tail -8 gpurun_out/r2_n2.err
echo; echo "== N=2 p2p, sharded push off"; SB200_SHARDED_PUSH=0 timeout 300 $T --exchange p2p 2> gpurun_out/r2_n2.err | python -c "$pick" || tail -8 gpurun_out/r2_n2.err
echo; echo "== oracle staging phases at 100M edges (CPU only)"
ORC_TIMING=1 timeout 300 python - <<'PY'
import numpy as np, time, oracle, bench
th = bench.host_threads()
N, E = 5_000_000, 100_000_000
t = time.time(); d = oracle.synth_edges(1, N, E, seed=42, scale=23, threads=th); print("gen", round(time.time() - t, 2), "threads", th)
a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
for threads in (th, 32, 128):
    t = time.time(); o = oracle.DenseHyperBall(*a, threads=threads, mt=True); ts = time.time() - t
    t = time.time(); it = o.run(); print("threads", threads, "stage", round(ts, 2), "run", round(time.time() - t, 2), it); o.close()
PY
