#!/bin/bash
# First GPU trip of round 2 (1 GPU, ~6-7 min): validate everything that was prepared without a GPU at the end of
# round 1, then measure each switch against the default.
#   gpurun --timeout 900 -- 'bash tools/gpu_r2_trip1.sh; tail -80 gpurun_out/r2_trip1.log'
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip1.log 2>&1
run() { echo; echo "== $*"; timeout "${T:-200}" "$@" 2>&1 | tail -${N:-6}; echo "rc=${PIPESTATUS[0]}"; }
T=200 run python -m pytest tests -m gpu -x -q
echo; echo "##### gated parity tests"
T=200 run env SB200_TEST_AND3=1 python -m pytest tests/test_bm25_gpu.py -q -k and3
T=300 run env SB200_TEST_OR3=1 python -m pytest tests/test_bm25_gpu.py -q -k or3
T=200 run env SB200_ARENA=1 python -m pytest tests/test_hyperball_gpu.py -x -q
T=200 run env SB200_STAGE_ROWPERM=1 python -m pytest tests/test_hyperball_gpu.py -x -q
T=200 run env SB200_L2_PERSIST_MB=64 python -m pytest tests/test_hyperball_gpu.py -x -q -k "random_graph or c1"
echo; echo "##### measurements (path 1, no CPU / BM25 legs)"
B="python bench.py --no-bm25 --no-cpu --steps 3 --warmup 3 --e2e-steps 4"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]; print("ms/step %.2f  top kernel %.3f ms  e2e %.0f ms  walls %s" % (d["ms_per_step"], r["avg_launch_ms"], e["ms_per_step"], e["step_wall_ms"]))'
for v in "" "SB200_ARENA=1" "SB200_STAGE_ROWPERM=1" "SB200_ARENA=1 SB200_STAGE_ROWPERM=1"; do
  echo; echo "== e2e with [$v]"; env $v timeout 150 $B 2> gpurun_out/r2_e2e.err | python -c "$pick" || tail -3 gpurun_out/r2_e2e.err
done
echo; echo "== SB200_L2_HINTS=1"; env SB200_L2_HINTS=1 timeout 150 $B --no-e2e 2> gpurun_out/r2_l2.err | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ms/step %.2f" % d["ms_per_step"], [(k["name"], round(k["ms"]/k["launches"],3)) for k in d["kernels"]])' || tail -3 gpurun_out/r2_l2.err
for mb in 32 64 96; do
  echo; echo "== SB200_L2_PERSIST_MB=$mb"; env SB200_L2_PERSIST_MB=$mb timeout 150 $B --no-e2e 2> gpurun_out/r2_l2.err | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ms/step %.2f" % d["ms_per_step"], [(k["name"], round(k["ms"]/k["launches"],3)) for k in d["kernels"]])' || tail -3 gpurun_out/r2_l2.err
done
echo; echo "##### BM25 signal combine (C4 at 1/10 scale): default vs k_or3"
for v in "" "SB200_BM25_OR3=1"; do
  echo "== [$v]"; env $v timeout 200 python - <<'PY'
from stract_b200 import bm25_bench
import bench
peaks, _ = bench._peaks()
r = bm25_bench.run_signal(0, peaks, max_doc=10_000_000, df_scale=2.0e6, cpu=False)
print({k: r[k] for k in ("value", "kernel_ms_per_batch", "docs_scored")}, "e2e", r["e2e"]["ms_per_batch"])
PY
done
echo; echo "##### BM25 AND: default vs unit kernel"
for v in "" "SB200_BM25_AND3=1"; do
  echo "== [$v]"; env $v timeout 200 python - <<'PY'
import json
from stract_b200 import bm25_bench
import bench
peaks, _ = bench._peaks()
r = bm25_bench.run_and(0, peaks, cpu=False)
print({k: r[k] for k in ("value", "kernel_ms_per_batch", "docs_scored", "blocks_decoded")}, "e2e", r["e2e"]["ms_per_batch"])
PY
done

# ---- 2-GPU follow-up (separate call: gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_r2_trip2.sh') is in tools/gpu_r2_trip2.sh
