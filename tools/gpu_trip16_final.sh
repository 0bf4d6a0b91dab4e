#!/bin/bash
# final round-1 evidence: all GPU tests, smoke, default bench (as the driver runs it), ncu memory sections at full size
mkdir -p gpurun_out
{
  echo "== pytest gpu (all)"
  timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -5
  echo "== smoke"
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  echo "== default bench"
  timeout 900 python bench.py > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err
  tail -c 600 gpurun_out/r01_bench_final.err
  python - <<'PY'
import json
d=json.load(open("gpurun_out/r01_bench_final.json"))
print('value',d['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'clocks',d['clocks'])
for k in d['kernels']: print('  ',k['name'],k['launches'],round(k['ms'],2),'GB/s alg',round(k['alg_bytes']/k['ms']/1e6,1),'share',round(k['share_of_step'],3))
print('roofline',d['roofline']); print('e2e',d['e2e']); print('cpu',d['cpu_baseline'])
for k,v in d['bm25'].items():
    if isinstance(v,dict): print(k,{x:v[x] for x in ('value','kernel_ms_per_batch','blocks_decoded') if x in v}); print('   e2e',v['e2e']['value'],'cpu',v['cpu_baseline']['value'],'roof',v['roofline']['frac'])
PY
  echo "== ncu memory sections, k_pull_warp<dense> at full size"
  timeout 240 ncu --section MemoryWorkloadAnalysis --section SpeedOfLight --section Occupancy --clock-control none -k regex:k_pull_warp -s 8 -c 1 -f -o gpurun_out/r01_pull_warp_full \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-bm25 2>&1 | tail -2 | cut -c1-300
} > gpurun_out/trip16.log 2>&1
tail -40 gpurun_out/trip16.log | cut -c1-900
