#!/bin/bash
# trip 2: all GPU parity tests (path 1 + path 2), BM25 bench at 1/10 scale then full
mkdir -p gpurun_out
{
  echo "== pytest gpu"
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
  echo "== bm25 bench scale 0.1"
  timeout 900 python - <<'PY'
import json, sys
sys.path.insert(0, ".")
from stract_b200 import bm25_bench
import bench
peaks, src = bench._peaks()
print(json.dumps(bm25_bench.run(0, peaks, src, scale=0.1)))
PY
  echo "== bm25 bench full"
  timeout 1800 python - <<'PY'
import json, sys
sys.path.insert(0, ".")
from stract_b200 import bm25_bench
import bench
peaks, src = bench._peaks()
print(json.dumps(bm25_bench.run(0, peaks, src, scale=1.0)))
PY
} > gpurun_out/trip2.log 2>&1
tail -40 gpurun_out/trip2.log
