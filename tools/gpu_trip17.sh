#!/bin/bash
# round-1 last validation: GPU parity suite + default bench after the pooled-state / page-locked result change
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/trip17.log 2>&1
echo "== pytest gpu"
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== default bench"
timeout 200 python bench.py > gpurun_out/r01_bench_e2e2.json 2> gpurun_out/r01_bench_e2e2.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('/root/repo/gpurun_out/r01_bench_e2e2.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'clocks', d['clocks'])
    print('roofline', d['roofline'])
    print('e2e', d['e2e'])
    for k, v in (d.get('bm25') or {}).items():
        if isinstance(v, dict): print(k, v.get('value'), v.get('kernel_ms_per_batch'), 'e2e', v.get('e2e'))
except Exception as ex:
    print('parse failed', ex)
PY
tail -5 gpurun_out/r01_bench_e2e2.err
