#!/bin/bash
# Round 2, trip 2 (1 GPU): new defaults + group tests, full bench with the full-size oracle parity, reference arm, ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_trip2.sh; tail -150 gpurun_out/r2_trip2.log'
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip2.log 2>&1
run() { echo; echo "== $*"; timeout "${T:-300}" "$@" 2>&1 | tail -${N:-8}; echo "rc=${PIPESTATUS[0]}"; }
T=400 run python -m pytest tests -m gpu -x -q
echo; echo "##### bench N=1 (full-size oracle parity, writes the C2 golden)"
timeout 900 python bench.py --steps 5 --warmup 3 --write-golden > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench_n1.err
cp tests/golden/path1_c2.json gpurun_out/ 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"] and d["e2e"].get("ms_min_median_max"))
print("parity", json.dumps(d["parity"])[:1500])
print("cpu", d.get("cpu_baseline"))
print("c1", d.get("c1"))
for k in d["kernels"]: print(k["kernel"], round(k["avg_launch_ms"], 3), round(k["frac"], 3), round(k["share_of_step"], 3))
b = d.get("bm25") or {}
for k, v in b.items():
    if isinstance(v, dict): print(k, v.get("value"), v.get("kernel_ms_per_batch"), v.get("e2e", {}).get("ms_per_batch"), v.get("parity"), v.get("cpu_baseline", {}).get("value"), v.get("cpu_baseline", {}).get("runs_s"))
PY
echo; echo "##### reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_ref.err; cut -c1-1500 gpurun_out/r02_bench_ref.json
echo; echo "##### ncu: launch list + full captures (path 1)"
B="python bench.py --no-e2e --no-cpu --no-bm25 --no-c1 --steps 1 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_path1.csv $B > /dev/null 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name "regex:k_pull_warp|k_pull_quad|k_finalize" --launch-count 3 -f -o gpurun_out/r02_path1_iter0 $B > gpurun_out/r02_ncu_path1.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02_path1_iter0.ncu-rep --page raw --csv > gpurun_out/r02_path1_iter0.ncu_raw.csv 2>/dev/null
echo; echo "##### ncu: BM25"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:k_and3" --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_bm25_and python tools/ncu_bm25.py and > gpurun_out/r02_ncu_and.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_ncu_and.log
ncu -i gpurun_out/r02_bm25_and.ncu-rep --page raw --csv > gpurun_out/r02_bm25_and.ncu_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:k_or3" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_bm25_signal python tools/ncu_bm25.py signal > gpurun_out/r02_ncu_signal.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_ncu_signal.log
ncu -i gpurun_out/r02_bm25_signal.ncu-rep --page raw --csv > gpurun_out/r02_bm25_signal.ncu_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
echo; echo "##### compute-sanitizer (memcheck) over smoke()"
T=400 N=15 run compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
