"""One BM25 batch at bench size for an ncu capture (tools/gpu_r2_trip*.sh): `and` = configs[3] (10M docs, 10k x 2-term AND,
top-1000), `signal` = configs[4] (100M docs, 10k x 5-term OR + 4 signals).  One warm-up batch, one captured batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import bench_bm25  # noqa: E402
from stract_b200 import bm25  # noqa: E402

what = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if what == "and":
    ix = bench_bm25.synth_index(int(10_000_000 * scale), 2.0e6 * scale)
    seg = bm25.SegmentReader(ix["postings"], ix["infos"], ix["fieldnorm_ids"], total_num_tokens=ix["total_num_tokens"])
    terms = bench_bm25.log_uniform_queries(10_000, 2)
    top = bm25.TopDocs.with_limit(1000)
    for _ in range(2):
        d, s, n, st = top.search_batch(seg, terms, bm25.MODE_AND, return_stats=True)
    print("and kernel_ms", st["kernel_ms"], "postings", st["postings_scored"])
else:
    max_doc = int(100_000_000 * scale)
    ix = bench_bm25.synth_index(max_doc, 2.0e7 * scale)
    rng = np.random.default_rng(99)
    cols = [rng.random(max_doc) ** 8, rng.random(max_doc), rng.random(max_doc), 1.0 / (1.0 + rng.integers(0, 1000, max_doc).astype(np.float64))]
    seg = bm25.SegmentReader(ix["postings"], ix["infos"], ix["fieldnorm_ids"], total_num_tokens=ix["total_num_tokens"])
    comp = bm25.SignalComputer(seg, bm25.SignalTable(cols), [2.0, 0.02, 2.0, 0.001], coeff_text=0.005)
    terms = bench_bm25.log_uniform_queries(10_000, 5, seed=2)
    for _ in range(2):
        d, tot, n, st = comp.top_docs_batch(terms, 1000, return_stats=True)
    print("signal kernel_ms", st["kernel_ms"], "postings", st["postings_scored"])
