#!/bin/bash
# e2e wall-clock breakdown: page-locked vs pageable result buffers (path 1 only, no CPU/BM25 legs)
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/trip18.log 2>&1
export SB200_RESULT_TIMING=1
echo "== pinned outputs"
timeout 110 python bench.py --no-bm25 --no-cpu --steps 1 --warmup 3 --e2e-steps 4 > gpurun_out/r01_e2e_pinned.json 2> gpurun_out/r01_e2e_pinned.err
echo "rc=$?"; grep "sb200 result" gpurun_out/r01_e2e_pinned.err | tail -12
python -c "
import json
d=json.loads(open('/root/repo/gpurun_out/r01_e2e_pinned.json').read().strip().splitlines()[-1]); print(d['e2e'])"
