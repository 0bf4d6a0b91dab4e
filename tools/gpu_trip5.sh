#!/bin/bash
# trip 5: verify BM25 restructure, ncu evidence for path 1, full-size bench of both paths
mkdir -p gpurun_out
{
  echo "== pytest bm25"
  timeout 300 python -m pytest tests/test_bm25_gpu.py -x -q --timeout 120 2>&1 | tail -5
  echo "== ncu --set full, k_pull_warp<dense> (100M-edge graph)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pull_warp -s 4 -c 2 -f -o gpurun_out/r01_pull_warp \
    python bench.py --nodes 5000000 --edges 100000000 --scale 23 --steps 1 --warmup 1 --no-cpu --no-e2e --no-bm25 2>&1 | tail -3
  echo "== ncu --set full, k_pull_quad<dense>"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pull_quad -s 4 -c 1 -f -o gpurun_out/r01_pull_quad \
    python bench.py --nodes 5000000 --edges 100000000 --scale 23 --steps 1 --warmup 1 --no-cpu --no-e2e --no-bm25 2>&1 | tail -3
  echo "== ncu launch list, full-size path-1 bench command"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_launches_path1.csv \
    python bench.py --steps 3 --no-cpu --no-e2e --no-bm25 2>&1 | tail -2
  echo "== full bench (both paths)"
  SB200_STAGE_TIMING=1 timeout 1500 python bench.py --steps 5 --e2e-steps 2 2> gpurun_out/r01_stage_timing.txt | tail -1 > gpurun_out/r01_bench_full.json
  tail -c 3000 gpurun_out/r01_bench_full.json
  tail -40 gpurun_out/r01_stage_timing.txt
} > gpurun_out/trip5.log 2>&1
tail -50 gpurun_out/trip5.log | cut -c1-2500
