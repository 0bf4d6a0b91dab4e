#!/bin/bash
mkdir -p gpurun_out
{
  for st in tail one onetail two many; do
    echo "== diag $st"; timeout 60 python tools/diag_bm25.py $st 2>&1 | tail -15
  done
  echo "== pytest path2"
  timeout 400 python -m pytest tests/test_bm25_gpu.py -x -q --timeout 60 2>&1 | tail -15
} > gpurun_out/trip3.log 2>&1
tail -60 gpurun_out/trip3.log
