#!/bin/bash
# Round 2, trip 7 (1 GPU): final validation of the tree (all GPU tests, smoke, racecheck), the QUAD2 experiment, the
# multi-field bench leg at full size.
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip7.log 2>&1
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== compute-sanitizer racecheck over smoke()"; timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
B="python bench.py --no-e2e --no-cpu --no-bm25 --no-c1 --steps 5 --warmup 3"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ms/step %.2f" % d["ms_per_step"], [(k["kernel"], round(k["avg_launch_ms"],3)) for k in d["kernels"]], d["parity"])'
echo; echo "== path 1 default"; timeout 200 $B 2> gpurun_out/r2_q.err | python -c "$pick" || tail -3 gpurun_out/r2_q.err
echo; echo "== path 1 SB200_QUAD2=1"; SB200_QUAD2=1 timeout 200 $B 2> gpurun_out/r2_q.err | python -c "$pick" || tail -3 gpurun_out/r2_q.err
echo; echo "== multi-field bench leg (full C4 size) + AND / signal legs quick"
timeout 900 python - <<'PY'
import json, bench, bench_bm25
peaks, src = bench._peaks()
r = bench_bm25.run_multi(0, peaks)
print(json.dumps({k: v for k, v in r.items() if k != "workload"}))
PY
echo; echo "== full default bench (driver arguments)"
timeout 1500 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2_full.err | tee gpurun_out/r02_bench_n1_final2.json | python -c "$pick; print({k: d[k] for k in ('value','ms_per_step','e2e','cpu_baseline','gpu_launches') if k in d}); print(json.dumps(d.get('bm25'))[:3000])" || tail -5 gpurun_out/r2_full.err
