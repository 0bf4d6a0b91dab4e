"""GPU tests added after the round's last GPU trip.  They passed on the CPU SIMT emulator (tests/emu) but have not run
on hardware yet, so they live in a file that sorts behind the hardware-validated suites: `pytest -x` reaches them last."""
import numpy as np
import pytest

import test_golden
from test_bm25_gpu import DFS, MODE_AND, MODE_OR, TopDocs, check_query, random_index
from test_hyperball_gpu import HarmonicCentrality, _graph, synth

pytestmark = pytest.mark.gpu


def test_cuda_path1_reproduces_golden():
    test_golden.check_path1_against_golden()


def test_cuda_c1_reproduces_golden():
    test_golden.check_c1_against_golden()


def test_cuda_path2_reproduces_golden():
    test_golden.check_path2_against_golden()


def test_positions_record_option_skip_entries():
    """IndexRecordOption::WithFreqsAndPositions, the option of Stract's position-bearing text fields
    (core/src/schema/text_field.rs:124-130): 12-byte skip entries with the block's tf sum (skip.rs:217-232).  Same
    results as the WithFreqs file of the same postings, and bit-exact against the oracle reading the 12-byte entries
    (AND, OR incl. block-max pruning in the oracle, signal combine)."""
    (o1, s1), _ = random_index(17, 60_000, DFS, record_option=1)
    (o2, s2), rng = random_index(17, 60_000, DFS, record_option=2)
    assert o2.postings_bytes().size > o1.postings_bytes().size      # 4 more bytes per full block
    nt = len(DFS)
    for _ in range(25):
        q = [int(x) for x in rng.choice(nt, int(rng.integers(1, 4)), replace=False)]
        for mode in (MODE_AND, MODE_OR):
            if mode == MODE_OR and len(q) > 2:
                continue
            check_query(o2, s2, q, mode, 100)
            a = TopDocs.with_limit(100).search(s1, q, mode); b = TopDocs.with_limit(100).search(s2, q, mode)
            assert a == b
    check_query(o2, s2, [nt - 1, nt - 2], MODE_AND, 1000)
    check_query(o2, s2, [nt - 1, nt - 2], MODE_OR, 1000)
    s1.close(); s2.close()



def test_rank_assignment_matches_store_harmonic_order():
    """store_harmonic's rank pass and top_nodes (webgraph/centrality/mod.rs:17-37,88-108) on the
    device, against the numpy restatement: order (centrality desc, id asc) resp. (centrality, id) desc, incl. the many
    exact ties a HyperLogLog-estimated centrality has."""
    from oracle import harmonic_ranks
    d = synth.rmat_graph(20_000, 120_000, seed=5)
    got = HarmonicCentrality.calculate(_graph(d), with_ranks=True, top=500)
    order = harmonic_ranks(got.ids_lo, got.ids_hi, got.values, ties_desc=False)
    assert len(np.unique(got.values)) < len(got.values)          # the tie rule is exercised
    rlo, rhi = got.rank_ids
    assert np.array_equal(rlo, got.ids_lo[order]) and np.array_equal(rhi, got.ids_hi[order])
    ranks = got.harmonic_rank()
    assert len(ranks) == len(got.values) and ranks[(int(rhi[0]) << 64) | int(rlo[0])] == 0
    torder = harmonic_ranks(got.ids_lo, got.ids_hi, got.values, ties_desc=True)[:500]
    tlo, thi, tc = got.top
    assert np.array_equal(tlo, got.ids_lo[torder]) and np.array_equal(thi, got.ids_hi[torder]) and np.array_equal(tc, got.values[torder])
    assert [c for _, c in got.top_nodes(10)] == sorted(got.values, reverse=True)[:10]


def test_term_info_store_decoded_on_device():
    """SURVEY 8(f) rank 2, the ordinal -> TermInfo half: the device decoder of tantivy's TermInfoStore against the store
    bytes the oracle's TermInfoStoreWriter produces -- the reference's own test_pack case (term_info_store.rs:330-357)
    and the TermInfo table of a real postings file, which must open the same segment."""
    import oracle
    from stract_b200 import bm25
    off = lambda i: i * 13 + i * i   # noqa: E731
    n = 1000
    ps = np.array([off(i) for i in range(n)], np.uint64); pe = np.array([off(i + 1) for i in range(n)], np.uint64)
    store = oracle.term_info_store_write(np.arange(n, dtype=np.uint32), ps, pe, ps * 3, pe * 3)
    infos, cnt = bm25.decode_term_info_store(store)
    assert cnt == n
    for i in range(n):
        assert (infos[i].doc_freq, infos[i].postings_off, infos[i].postings_len) == (i, off(i), off(i + 1) - off(i)), i
    (oseg, seg), rng = random_index(19, 60_000, DFS)
    data = oseg.postings_bytes()
    t_off, t_len, t_df = oseg.term_infos()
    store = oracle.term_info_store_write(t_df, t_off, t_off + t_len)
    infos, cnt = bm25.decode_term_info_store(store)
    assert cnt == len(DFS)
    seg2 = bm25.SegmentReader(data, infos, oseg.fieldnorm_ids)
    for q in ([len(DFS) - 1, len(DFS) - 2], [3, len(DFS) - 1], [len(DFS) - 4]):
        for mode in (MODE_AND, MODE_OR):
            assert TopDocs.with_limit(100).search(seg, q, mode) == TopDocs.with_limit(100).search(seg2, q, mode)
    seg.close(); seg2.close()


def test_searcher_over_three_segments_matches_one_big_segment():
    """tantivy Searcher semantics for path A: index-wide BM25 statistics (bm25.rs:98-134) and merge_fruits
    (top_collector.rs:109-129).  The same collection as ONE oracle segment and as THREE device segments must give the
    same (doc, score) lists, bit for bit -- (segment_ord, doc) order equals global doc order for contiguous splits."""
    import oracle
    from stract_b200 import bm25
    from stract_b200.bm25 import NO_TERM, Searcher, SegmentReader
    rng = np.random.default_rng(27)
    max_doc, cuts = 45_000, [0, 12_000, 30_000, 45_000]
    dfs = [40, 300, 2_000, 9_000, 20_000]
    lens = np.maximum(1, rng.lognormal(4.0, 0.8, max_doc)).astype(np.uint32)
    ids = bm25.fieldnorms_to_ids(lens)
    td = [np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32) for df in dfs]
    td[0] = td[0][td[0] >= cuts[1]]                      # the rarest term does not occur in segment 0 at all
    tt = [np.minimum(rng.geometric(0.6, len(d)), 255).astype(np.uint32) for d in td]
    whole = oracle.Segment(ids)
    for d, t in zip(td, tt):
        whole.add_term(d, t)
    segs, ords = [], []
    for s in range(3):
        lo, hi = cuts[s], cuts[s + 1]
        sd, st, present = [], [], []
        for d, t in zip(td, tt):
            m = (d >= lo) & (d < hi)
            if m.any():
                present.append(len(sd)); sd.append(d[m] - lo); st.append(t[m])
            else:
                present.append(NO_TERM)
        data, infos = bm25.encode_postings(sd, st, ids[lo:hi], 1.0)
        segs.append(SegmentReader(data, infos, ids[lo:hi]))
        ords.append(present)
    searcher = Searcher(segs)
    assert searcher.total_num_docs == max_doc and np.float32(searcher.average_fieldnorm) == np.float32(whole.avg_fieldnorm)
    queries = [[4, 3], [0, 4], [2, 1], [0, 1], [3], [4, 2]]
    cache = bm25.compute_tf_cache(searcher.average_fieldnorm)
    full_df = np.array([len(d) for d in td])
    for mode in (MODE_AND, MODE_OR):
        for q in queries:
            per_seg = [np.array([[o[t] for t in q]], np.uint32) for o in ords]
            for limit, offset in ((50, 0), (20, 7)):
                sg, dd, sc, n = searcher.search_batch(TopDocs.with_limit(limit).and_offset(offset), per_seg, mode)
                w = np.array([bm25.Bm25Weight.for_one_term(int(full_df[t]), max_doc, searcher.average_fieldnorm).weight for t in q], np.float32)
                od, os_, _ = whole.topk(np.array(q, np.uint32), w, np.tile(cache, (len(q), 1)), mode, limit + offset)
                od, os_ = od[offset:], os_[offset:]
                glob = np.array(cuts, np.uint32)[sg[0, :n[0]]] + dd[0, :n[0]]
                assert np.array_equal(glob, od), (mode, q, limit, offset)
                assert np.array_equal(sc[0, :n[0]], os_), (mode, q, limit, offset)
    for s in segs:
        s.close()


def test_signal_searcher_over_segments_matches_one_big_segment():
    """Path B with searcher-wide BM25 statistics (core/src/ranking/bm25.rs:52-92) and the fruit merge: three device
    segments with their own signal tables against one oracle segment holding everything."""
    import oracle
    from stract_b200 import bm25
    from stract_b200.bm25 import NO_TERM, SegmentReader, SignalComputer, SignalSearcher, SignalTable
    rng = np.random.default_rng(29)
    max_doc, cuts = 36_000, [0, 9_000, 25_000, 36_000]
    dfs = [60, 500, 3_000, 8_000, 15_000]
    lens = np.maximum(1, rng.lognormal(4.0, 0.8, max_doc)).astype(np.uint32)
    ids = bm25.fieldnorms_to_ids(lens)
    td = [np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32) for df in dfs]
    td[0] = td[0][td[0] < cuts[2]]                       # the rarest term is missing from the last segment
    tt = [np.minimum(rng.geometric(0.6, len(d)), 255).astype(np.uint32) for d in td]
    cols = [rng.random(max_doc) ** 8, rng.random(max_doc), rng.random(max_doc), 1.0 / (1.0 + rng.integers(0, 1000, max_doc))]
    coeffs = [2.0, 0.02, 2.0, 0.001]
    whole = oracle.Segment(ids)
    for d, t in zip(td, tt):
        whole.add_term(d, t)
    comps, ords, keep = [], [], []
    for s in range(3):
        lo, hi = cuts[s], cuts[s + 1]
        sd, st, present = [], [], []
        for d, t in zip(td, tt):
            m = (d >= lo) & (d < hi)
            if m.any():
                present.append(len(sd)); sd.append(d[m] - lo); st.append(t[m])
            else:
                present.append(NO_TERM)
        data, infos = bm25.encode_postings(sd, st, ids[lo:hi], 1.0)
        seg = SegmentReader(data, infos, ids[lo:hi])
        table = SignalTable([c[lo:hi] for c in cols])
        comps.append(SignalComputer(seg, table, coeffs, coeff_text=0.005)); ords.append(present); keep.append((seg, table))
    searcher = SignalSearcher(comps)
    queries = np.array([[4, 3, 2, 1, 0], [0, 1, 2, 3, 4], [2, 4, 1, 3, 0]], np.uint32)
    per_seg = [np.array([[o[t] for t in q] for q in queries], np.uint32) for o in ords]
    sg, dd, tot, n = searcher.top_docs_batch(per_seg, 200)
    cache = bm25.compute_tf_cache(searcher.average_fieldnorm)
    full_df = np.array([len(d) for d in td])
    w = np.array([[bm25.StractBm25Weight.for_one_term(int(full_df[t]), max_doc, searcher.average_fieldnorm).weight for t in q] for q in queries], np.float32)
    od, ot, on, _ = whole.signal_topk_batch(queries, w, np.tile(cache, (queries.size, 1)), 1.2, 0.005, cols, coeffs, 200)
    assert np.array_equal(n, on)
    for q in range(len(queries)):
        glob = np.array(cuts, np.uint32)[sg[q, :n[q]]] + dd[q, :n[q]]
        assert np.array_equal(glob, od[q, :on[q]]) and np.array_equal(tot[q, :n[q]], ot[q, :on[q]]), q
    for seg, table in keep:
        table.close(); seg.close()
