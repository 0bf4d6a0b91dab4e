#!/bin/bash
mkdir -p gpurun_out
{
  for st in one onetail two; do
    echo "== diag $st"; timeout 60 python tools/diag_bm25.py $st 2>&1 | tail -8
  done
  echo "== pytest gpu (all)"
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15
  echo "== smoke"
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  echo "== path1 bench 5M/100M with stage timing"
  SB200_STAGE_TIMING=1 timeout 300 python bench.py --nodes 5000000 --edges 100000000 --scale 23 --steps 2 --warmup 1 --no-cpu --no-e2e --no-bm25 2>&1 | tail -30
  echo "== bm25 bench scale 0.1"
  timeout 400 python - <<'PY'
import json, sys
sys.path.insert(0, ".")
from stract_b200 import bm25_bench
import bench
peaks, src = bench._peaks()
print(json.dumps(bm25_bench.run(0, peaks, src, scale=0.1)))
PY
} > gpurun_out/trip4.log 2>&1
tail -80 gpurun_out/trip4.log | cut -c1-1500
