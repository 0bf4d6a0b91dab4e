#!/bin/bash
mkdir -p gpurun_out
{
  echo "== pytest path1 (staging rewrite)"
  timeout 600 python -m pytest tests/test_hyperball_gpu.py -x -q --timeout 180 2>&1 | tail -6
  echo "== full bench path1 with e2e"
  SB200_STAGE_TIMING=1 timeout 900 python bench.py --steps 3 --e2e-steps 2 --no-bm25 --no-cpu 2> gpurun_out/r01_stage_timing_c.txt | tail -1 > gpurun_out/r01_bench_path1_c.json
  python - <<'PY'
import json
d=json.load(open("gpurun_out/r01_bench_path1_c.json"))
print('value',d['value'],'ms/step',d['ms_per_step'],'stage_ms',d['config']['stage_ms'])
print('e2e',d['e2e'])
PY
  tail -22 gpurun_out/r01_stage_timing_c.txt
} > gpurun_out/trip14.log 2>&1
tail -40 gpurun_out/trip14.log | cut -c1-700
