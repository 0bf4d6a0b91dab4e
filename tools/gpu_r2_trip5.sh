#!/bin/bash
# Round 2, trip 5 (1 GPU): union-kernel occupancy variants, then the full default bench + reference arm (final numbers).
cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2_trip5.log 2>&1
echo "== tests (bm25 + sharded group only; the full suite ran in trip 4)"; timeout 600 python -m pytest tests/test_bm25_gpu.py tests/test_sharded_gpu.py tests/test_hyperball_gpu.py -m gpu -x -q 2>&1 | tail -4
echo; echo "== union kernel occupancy variants (C5 at full size, TMA on)"
for o in 5 6 8; do SB200_OR3_OCC=$o timeout 400 python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, ".")
import bench_bm25
from stract_b200 import bm25
max_doc = 100_000_000
ix = bench_bm25.synth_index(max_doc, 2.0e7)
rng = np.random.default_rng(99)
cols = [rng.random(max_doc) ** 8, rng.random(max_doc), rng.random(max_doc), 1.0 / (1.0 + rng.integers(0, 1000, max_doc).astype(np.float64))]
seg = bm25.SegmentReader(ix["postings"], ix["infos"], ix["fieldnorm_ids"], total_num_tokens=ix["total_num_tokens"])
comp = bm25.SignalComputer(seg, bm25.SignalTable(cols), [2.0, 0.02, 2.0, 0.001], coeff_text=0.005)
terms = bench_bm25.log_uniform_queries(10_000, 5, seed=2)
comp.top_docs_batch(terms, 1000)
ks = []
for _ in range(3):
    d, tot, n, st = comp.top_docs_batch(terms, 1000, return_stats=True); ks.append(st["kernel_ms"])
print("SIGNAL occ", os.environ["SB200_OR3_OCC"], "kernel_ms", [round(x, 1) for x in ks], "checksum", int(d.astype(np.uint64).sum()), float(tot.sum()))
PY
done
echo; echo "== AND e2e after the packed copy"
timeout 200 python tools/bm25_variants.py and 2>&1 | tail -1
SB200_BM25_DENSE_OUT=1 timeout 200 python tools/bm25_variants.py and 2>&1 | tail -1
echo; echo "##### bench N=1"
timeout 1200 python bench.py --steps 10 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"] and d["e2e"].get("ms_min_median_max"))
print("parity", {k: v.get("green") for k, v in d["parity"].items()})
print("cpu", d.get("cpu_baseline"))
print("c1", {k: d["c1"][k] for k in ("ms_per_step", "e2e_ms_per_step", "parity")} if "c1" in d else None)
for k in d["kernels"]: print(k["kernel"], round(k["avg_launch_ms"], 3), "frac", round(k["frac"], 3), "dram_frac", k["dram_frac"] and round(k["dram_frac"], 3), "share", round(k["share_of_step"], 3))
b = d.get("bm25") or {}
for k, v in b.items():
    if isinstance(v, dict): print(k, "value %.3e" % v["value"], v.get("kernel_ms_per_batch"), "e2e", v["e2e"]["ms_per_batch"], v.get("parity", {}).get("green"), "cpu %.3e" % v["cpu_baseline"]["value"], v["cpu_baseline"].get("runs_s"), v.get("max_docs_250k"))
PY
echo; echo "##### reference arm"
timeout 900 python bench.py --impl reference --steps 10 --warmup 5 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_ref_final.err; cut -c1-1200 gpurun_out/r02_bench_ref_final.json
