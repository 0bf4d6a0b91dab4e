"""Step-by-step diagnostic of the BM25 GPU path (run each stage under a short timeout)."""
import faulthandler, sys, time
faulthandler.dump_traceback_later(45, exit=True)
import numpy as np
sys.path.insert(0, ".")
stage = sys.argv[1]
from stract_b200 import bm25
from stract_b200.bm25 import SegmentReader, TopDocs, MODE_AND, MODE_OR
rng = np.random.default_rng(11)
max_doc = 60000
DFS = {"tail": [100], "one": [128], "onetail": [300], "two": [255, 6000], "many": [1, 3, 100, 127, 128, 129, 255, 256, 257, 300, 511, 512, 1000, 1024, 2500, 6000, 15000, 40000]}[stage if stage in ("tail", "one", "onetail", "two") else "many"]
lens = np.maximum(1, rng.lognormal(4.0, 0.8, max_doc)).astype(np.uint32)
ids = bm25.fieldnorms_to_ids(lens)
td = [np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32) for df in DFS]
tt = [np.minimum(rng.geometric(0.6, df), 255).astype(np.uint32) for df in DFS]
avg = np.float32(np.float32(bm25.fieldnorm_table()[ids].astype(np.uint64).sum()) / np.float32(max_doc))
data, infos = bm25.encode_postings(td, tt, ids, avg)
print(stage, "encoded", len(data), flush=True)
t0 = time.time(); seg = SegmentReader(data, infos, ids); print("segment created", seg.info(), round(time.time() - t0, 2), flush=True)
q = list(range(min(len(DFS), 2)))
for mode, name in ((MODE_AND, "AND"), (MODE_OR, "OR")):
    for k in (1, 10, 1000):
        t0 = time.time()
        r = TopDocs.with_limit(k).search(seg, q, mode)
        print(name, "k", k, "->", len(r), r[:2], round(time.time() - t0, 3), flush=True)
print("done", stage, flush=True)
