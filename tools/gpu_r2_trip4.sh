#!/bin/bash
# Round 2, trip 4 (1 GPU): new GPU tests (multi-field signals, graph searches, group run), BM25 kernel switches, ncu of the
# union kernel at full size, host CPU probe.
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip4.log 2>&1
echo "== host"; nproc; python -c "import os,bench; print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), 'host_threads', bench.host_threads())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2
echo "== tests"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo; echo "== AND occupancy variants"
for o in 5 6 8; do SB200_AND3_OCC=$o timeout 200 python tools/bm25_variants.py and 2>&1 | tail -1; done
echo; echo "== union kernel: TMA staging on/off (full C5 size)"
timeout 900 python tools/bm25_variants.py signal 2>&1 | tail -8
echo; echo "== ncu union kernel (full size, reduced sections)"
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats \
  --clock-control none --kernel-name "regex:k_or3" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_bm25_signal python tools/ncu_bm25.py signal > gpurun_out/r02_ncu_signal.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_ncu_signal.log
ncu -i gpurun_out/r02_bm25_signal.ncu-rep --page raw --csv > gpurun_out/r02_bm25_signal.ncu_raw.csv 2>/dev/null; ls -la gpurun_out/r02_bm25_signal*
echo; echo "== oracle staging phases at 100M edges (CPU only)"
ORC_TIMING=1 timeout 400 python - <<'PY'
import numpy as np, time, oracle, bench
th = bench.host_threads()
N, E = 5_000_000, 100_000_000
t = time.time(); d = oracle.synth_edges(1, N, E, seed=42, scale=23, threads=th); print("gen", round(time.time() - t, 2), "threads", th)
a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
for threads in sorted({th, 16, 32, 128}):
    t = time.time(); o = oracle.DenseHyperBall(*a, threads=threads, mt=True); ts = time.time() - t
    t = time.time(); it = o.run(); print("threads", threads, "stage", round(ts, 2), "run", round(time.time() - t, 2), it, flush=True); o.close()
PY
