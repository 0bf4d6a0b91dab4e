"""Host-side (CPU) pieces of path 2 against the oracle: fieldnorm code, BM25 weights, posting writer bytes."""
import numpy as np

import oracle
from stract_b200 import bm25


def test_fieldnorm_code_matches():
    assert all(bm25.id_to_fieldnorm(i) == oracle.id_to_fieldnorm(i) for i in range(256))
    for v in list(range(3000)) + [10**6, 2**31, 2**32 - 1]:
        assert bm25.fieldnorm_to_id(v) == oracle.fieldnorm_to_id(v)


def test_weights_bit_equal():
    for df, n, avg in [(3, 6, 10.0), (300, 1024, 10.0), (10, 129, 20.0), (12345, 10**7, 247.3), (1, 2, 1.5)]:
        w, c = oracle.tv_bm25_weight(df, n, avg)
        m = bm25.Bm25Weight.for_one_term(df, n, avg)
        assert w == m.weight and np.array_equal(c, m.cache)
        w, c = oracle.stract_bm25_weight(df, n, avg)
        m = bm25.StractBm25Weight.for_one_term(df, n, avg)
        assert w == m.weight and np.array_equal(c, m.cache)


def test_writer_bytes_equal_oracle_writer():
    rng = np.random.default_rng(1)
    max_doc = 50_000
    lens = np.maximum(1, rng.lognormal(4, 0.8, max_doc)).astype(np.uint32)
    ids = bm25.fieldnorms_to_ids(lens)
    seg = oracle.Segment(ids)
    td, tt = [], []
    for df in [1, 5, 127, 128, 129, 255, 256, 300, 1000, 20000, 3]:
        d = np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32)
        t = np.minimum(rng.geometric(0.5, df), 300).astype(np.uint32)
        td.append(d); tt.append(t); seg.add_term(d, t)
    data, infos = bm25.encode_postings(td, tt, ids, seg.avg_fieldnorm)
    off, ln, df = seg.term_infos()
    assert np.array_equal(data, seg.postings_bytes())
    assert all(infos[i].postings_off == off[i] and infos[i].postings_len == ln[i] and infos[i].doc_freq == df[i] for i in range(len(df)))


def test_score_rank():
    assert bm25.score_rank(0) == 10.0
    assert abs(bm25.score_rank(7) - 9.0) < 1e-12
    assert bm25.score_rank(8**11) == 0.0
