#!/bin/bash
# first GPU trip: environment probe, path-1 parity tests, smoke, scaled + full bench
mkdir -p gpurun_out
{
  echo "== env"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  echo "== pytest gpu"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
  echo "== bench small (5M/100M)"
  timeout 600 python bench.py --nodes 5000000 --edges 100000000 --scale 23 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 2>&1 | tail -3
  echo "== bench full"
  timeout 1500 python bench.py --steps 3 --e2e-steps 1 2>&1 | tail -3
} > gpurun_out/trip1.log 2>&1
tail -60 gpurun_out/trip1.log
