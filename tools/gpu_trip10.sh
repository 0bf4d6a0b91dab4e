#!/bin/bash
mkdir -p gpurun_out
{
  echo "== pytest gpu (all)"
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -12
  echo "== full bench (both paths, e2e, cpu)"
  SB200_STAGE_TIMING=1 timeout 1200 python bench.py --steps 5 --e2e-steps 2 2> gpurun_out/r01_stage_timing_b.txt | tail -1 > gpurun_out/r01_bench_full_b.json
  python - <<'PY'
import json
d=json.load(open("gpurun_out/r01_bench_full_b.json"))
print('value',d['value'],'ms/step',d['ms_per_step'])
for k in d['kernels']: print('  ',k['name'],k['launches'],round(k['ms'],2),'GB/s alg',round(k['alg_bytes']/k['ms']/1e6,1),'share',round(k['share_of_step'],3))
print([ (p['mode'],p['ms']) for p in d['per_iter']])
print('roofline',d['roofline']); print('e2e',d['e2e']); print('cpu',d['cpu_baseline']); print('clocks',d['clocks'],'launches',d['gpu_launches'])
for k,v in d['bm25'].items():
    if isinstance(v,dict): print(k,{x:v[x] for x in ('value','kernel_ms_per_batch','blocks_decoded') if x in v}); print('   e2e',v['e2e']['value'],'cpu',v['cpu_baseline']['value'])
PY
  tail -12 gpurun_out/r01_stage_timing_b.txt
} > gpurun_out/trip10.log 2>&1
tail -45 gpurun_out/trip10.log | cut -c1-900
