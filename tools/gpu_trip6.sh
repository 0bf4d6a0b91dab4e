#!/bin/bash
mkdir -p gpurun_out
{
  echo "== pytest gpu (all)"
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -12
  echo "== path1 bench full, no e2e"
  timeout 300 python bench.py --steps 5 --no-e2e --no-cpu --no-bm25 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms/step',d['ms_per_step'])
for k in d['kernels']: print('  ',k['name'],k['launches'],round(k['ms'],2),'GB/s alg',round(k['alg_bytes']/k['ms']/1e6,1),'share',round(k['share_of_step'],3))
for p in d['per_iter']: print('  ',p)
print(d['roofline'])"
  echo "== bm25 bench full"
  timeout 900 python - <<'PY'
import json, sys
sys.path.insert(0, ".")
from stract_b200 import bm25_bench
import bench
peaks, src = bench._peaks()
r = bm25_bench.run(0, peaks, src, scale=1.0)
for k, v in r.items():
    if isinstance(v, dict):
        print(k, {x: v[x] for x in ("value", "kernel_ms_per_batch", "postings_per_batch", "docs_scored", "blocks_decoded") if x in v})
        print("   e2e", v["e2e"]); print("   roofline", v["roofline"]); print("   cpu", v.get("cpu_baseline"))
PY
} > gpurun_out/trip6.log 2>&1
tail -60 gpurun_out/trip6.log | cut -c1-1200
