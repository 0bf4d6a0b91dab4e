"""Pins the path-2 (BM25) oracle against the reference's known-answer tests.  CPU only.
assert_nearly_equals! in the reference is |a-b| < 5e-4 relative-ish (crates/tantivy/src/lib.rs); we use 1e-6."""
import math

import numpy as np
import pytest

import oracle
from oracle import Segment, fieldnorm_to_id, id_to_fieldnorm, fieldnorms_to_ids, tv_bm25_weight, stract_bm25_weight

TERMINATED = 0x7FFFFFFF


def near(a, b, tol=1e-6):
    return abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b)))


def test_fieldnorm_code():
    # tantivy/src/fieldnorm/code.rs:276-296
    assert fieldnorm_to_id(0) == 0 and fieldnorm_to_id(1) == 1
    for i in range(41):
        assert fieldnorm_to_id(i) == i
    assert fieldnorm_to_id(41) == 40 and fieldnorm_to_id(42) == 41
    for i in range(43, 256):
        fn = id_to_fieldnorm(i)
        assert fieldnorm_to_id(fn) == i and fieldnorm_to_id(fn - 1) == i - 1 and fieldnorm_to_id(fn + 1) == i
    assert fieldnorm_to_id(0xFFFFFFFF) == 255
    assert id_to_fieldnorm(254) == 1_879_048_216 and id_to_fieldnorm(255) == 2_013_265_944  # code.rs:268-269


def test_bitpacker4x_roundtrip_and_size():
    # compression/mod.rs:268-330: round trips; compressed size = num_bits * 16 bytes
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for nb in range(0, 33):
        hi = (1 << nb) - 1
        vals = rng.integers(0, hi + 1, 128, dtype=np.uint64).astype(np.uint32) if nb else np.zeros(128, np.uint32)
        packed = np.zeros(512, np.uint8); back = np.zeros(128, np.uint32)
        n = L.orc_bp4_roundtrip(vals, nb, packed, back)
        assert n == nb * 16 and np.array_equal(vals, back)
    # documented layout: value k sits in lane k%4 of a 4-lane interleaved stream
    vals = np.zeros(128, np.uint32); vals[5] = 1   # lane 1, position 1, nb=1 -> bit 1 of u32 word index 1
    packed = np.zeros(512, np.uint8); back = np.zeros(128, np.uint32)
    L.orc_bp4_roundtrip(vals, 1, packed, back)
    words = packed[:16].view(np.uint32)
    assert list(words) == [0, 2, 0, 0]


def test_vint_and_bitwidth():
    L = oracle.lib()
    assert L.orc_encode_bitwidth(2, 1) == 0b01000010  # skip.rs:314-320
    # compression/mod.rs:332-372: vint-encoded sorted block of the given input is <= 154 bytes ... and round trips
    vals = np.arange(0, 128 * 7, 7, dtype=np.uint32)[:100] * 11 + 3
    out = np.zeros(1024, np.uint8)
    n = L.orc_vint_sorted_encode(vals, vals.size, 0, out)
    # decode by hand: 7-bit groups, stop bit on the last byte
    got, cur, sh, acc = [], 0, 0, 0
    for b in out[:n]:
        acc += (int(b) & 127) << sh
        if b & 128:
            cur += acc; got.append(cur); acc = 0; sh = 0
        else:
            sh += 7
    assert got == list(vals)


def test_idf():
    assert near(oracle.lib().orc_tv_idf(1, 2), math.log(2.0))  # bm25.rs:238-243


def _seg_from(doc_tfs, fieldnorms):
    ids = fieldnorms_to_ids(fieldnorms)
    avg = (sum(fieldnorms) / len(fieldnorms)) if len(fieldnorms) else 0.0  # create_from_docs_and_tfs segment_postings.rs:85-96
    seg = Segment(ids, avg_fieldnorm=np.float32(avg))
    docs = np.array([d for d, _ in doc_tfs], np.uint32); tfs = np.array([t for _, t in doc_tfs], np.uint32)
    t = seg.add_term(docs, tfs)
    return seg, t


def test_term_scorer_max_score():
    # term_scorer.rs:142-164
    w, cache = tv_bm25_weight(3, 6, 10.0)
    seg, t = _seg_from([(2, 3), (3, 12), (7, 8)], [0, 0, 10, 12, 0, 0, 0, 100])
    c = seg.cursor(t, w, cache)
    assert near(c.max_score(), 1.3990127)
    assert c.doc() == 2 and c.tf() == 3
    assert near(c.block_max_score(), 1.3676447)
    assert near(c.score(), 1.0892314)
    assert c.advance() == 3 and c.doc() == 3 and c.tf() == 12
    assert near(c.score(), 1.3676447)
    assert c.advance() == 7 and c.tf() == 8
    assert near(c.score(), 0.72015285)
    assert c.advance() == TERMINATED


def test_term_scorer_shallow_advance():
    # term_scorer.rs:166-182
    w, cache = tv_bm25_weight(300, 1024, 10.0)
    doc_tfs = [(i * 10, 1 + (i * 10) % 3) for i in range(300)]
    seg, t = _seg_from(doc_tfs, [10] * 3000)
    c = seg.cursor(t, w, cache)
    assert c.doc() == 0
    c.shallow_seek(1289)
    assert c.doc() == 0
    assert c.seek(1289) == 1290 and c.doc() == 1290


def test_block_wand_block_max_kat():
    # term_scorer.rs:229-253
    doc_tfs = [(d, 1) for d in range(128)] + [(d, 2 if d == 200 else 1) for d in range(128, 256)] + [(256, 1), (257, 3), (258, 1)]
    w, cache = tv_bm25_weight(10, 129, 20.0)
    seg, t = _seg_from(doc_tfs, [20] * 300)
    c = seg.cursor(t, w, cache)
    assert near(c.block_max_score(), 2.5161593)
    c.shallow_seek(135)
    assert near(c.block_max_score(), 3.4597192)
    c.shallow_seek(256)
    assert near(c.block_max_score(), 5.2971773)  # block not loaded -> max_score()
    assert c.seek(256) == 256
    assert near(c.block_max_score(), 3.9539647)


@pytest.mark.parametrize("seed", range(8))
def test_block_max_score_property(seed):
    # term_scorer.rs:185-225 (proptest): the stored block max equals the max of the scores in the block
    rng = np.random.default_rng(seed)
    n = int(rng.integers(80, 300))
    tf = rng.integers(1, 10, n); extra = rng.integers(0, 100, n)
    fieldnorms = list((tf + extra).astype(int))
    avg = np.float32(np.float32(sum(fieldnorms)) / np.float32(n))
    w, cache = tv_bm25_weight(n, n * 10, avg)
    seg, t = _seg_from([(d, int(tf[d])) for d in range(n)], fieldnorms)
    c = seg.cursor(t, w, cache)
    for b in range(0, n, 128):
        bm = c.block_max_score(); best = 0.0
        for d in range(b, min(b + 128, n)):
            assert c.doc() == d
            best = max(best, c.score()); c.advance()
        assert near(bm, best, 1e-5)


def _index(docs_tokens):
    """Tiny in-RAM index of one text field: returns (segment, {term: ord}, {term: df}, avg_fieldnorm)."""
    lens = [len(t) for t in docs_tokens]
    seg = Segment(fieldnorms_to_ids(lens))
    vocab = sorted({w for t in docs_tokens for w in t})
    ords, dfs = {}, {}
    for w in vocab:
        docs = [d for d, t in enumerate(docs_tokens) if w in t]
        tfs = [docs_tokens[d].count(w) for d in docs]
        ords[w] = seg.add_term(docs, tfs); dfs[w] = len(docs)
    return seg, ords, dfs, np.float32(np.float32(sum(lens)) / np.float32(len(lens)))


def test_top_docs_droopy_tax_kat():
    # collector/top_score_collector.rs:590-602,676-700: query `droopy tax` (OR) over 3 docs
    docs = ["hello happy tax payer".split(), "droopy says hello happy tax payer".split(), "i like droopy".split()]
    seg, ords, dfs, avg = _index(docs)
    q = ["droopy", "tax"]
    wc = [tv_bm25_weight(dfs[t], 3, avg) for t in q]
    weights = [w for w, _ in wc]; caches = np.stack([c for _, c in wc])
    for mode in (1, 2):
        d, s, _ = seg.topk([ords[t] for t in q], weights, caches, mode, 4)
        assert list(d) == [1, 2, 0]
        assert near(s[0], 0.81221175) and near(s[1], 0.5376842) and near(s[2], 0.48527452)
    d, s, _ = seg.topk([ords[t] for t in q], weights, caches, 1, 2)  # limit 2 (:702-715)
    assert list(d) == [1, 2]
    d, s, _ = seg.topk([ords[t] for t in q], weights, caches, 0, 4)  # AND: only doc 1 has both
    assert list(d) == [1] and near(s[0], 0.81221175)


def test_topn_order_ties_by_doc():
    # top_collector.rs:50-66 / top_score_collector.rs:754-779: equal scores -> ascending doc
    n = 700
    seg = Segment(fieldnorms_to_ids([5] * n))
    t = seg.add_term(np.arange(n, dtype=np.uint32), np.ones(n, np.uint32))
    w, cache = tv_bm25_weight(n, n, 5.0)
    d, s, _ = seg.topk([t], [w], cache[None], 1, 300)
    assert list(d) == list(range(300)) and len(set(s.tolist())) == 1


def _random_index(rng, max_doc, n_terms, max_df):
    lens = np.maximum(1, rng.lognormal(3.0, 0.8, max_doc)).astype(np.uint32)
    seg = Segment(fieldnorms_to_ids(lens))
    dfs = []
    for _ in range(n_terms):
        df = int(rng.integers(1, max_df))
        docs = np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32)
        tfs = np.minimum(rng.geometric(0.6, df), 255).astype(np.uint32)
        seg.add_term(docs, tfs); dfs.append(df)
    return seg, np.array(dfs), np.float32(np.float32(lens.sum()) / np.float32(max_doc))


@pytest.mark.parametrize("seed", range(6))
def test_block_wand_equals_exhaustive_union(seed):
    # block_wand.rs:336-508 (proptests + explicit case): pruning must return the exhaustive top-k.
    # 1- and 2-term sums are order independent (f32 + is commutative) -> bit-exact.
    rng = np.random.default_rng(100 + seed)
    seg, dfs, avg = _random_index(rng, 4000, 12, 1500)
    for _ in range(12):
        nt = int(rng.integers(1, 3))
        q = rng.choice(len(dfs), nt, replace=False)
        wc = [tv_bm25_weight(dfs[t], 4000, avg) for t in q]
        weights = [w for w, _ in wc]; caches = np.stack([c for _, c in wc])
        for k in (1, 10, 200):
            d1, s1, sc1 = seg.topk(q, weights, caches, 1, k)
            d2, s2, sc2 = seg.topk(q, weights, caches, 2, k)
            assert np.array_equal(d1, d2) and np.array_equal(s1, s2)
            assert sc1 <= sc2


def test_intersection_matches_bruteforce():
    # intersection.rs:161-248 semantics + score association left + right + others
    rng = np.random.default_rng(7)
    seg, dfs, avg = _random_index(rng, 6000, 10, 3000)
    off, ln, df = seg.term_infos()
    for nt in (2, 3, 4):
        q = rng.choice(len(dfs), nt, replace=False)
        wc = [tv_bm25_weight(dfs[t], 6000, avg) for t in q]
        weights = [w for w, _ in wc]; caches = np.stack([c for _, c in wc])
        d, s, _ = seg.topk(q, weights, caches, 0, 5000)
        # brute force through single-term exhaustive scans
        per = []
        for i, t in enumerate(q):
            dd, ss, _ = seg.topk([t], [weights[i]], caches[i][None], 2, 6000)
            per.append(dict(zip(dd.tolist(), ss.tolist())))
        common = set(per[0])
        for p in per[1:]:
            common &= set(p)
        order = sorted(range(nt), key=lambda i: (dfs[q[i]], i))  # sort_by_key(size_hint), stable
        exp = {}
        for doc in common:
            others = np.float32(0.0)
            for i in order[2:]:
                others = np.float32(others + np.float32(per[i][doc]))
            exp[doc] = np.float32(np.float32(np.float32(per[order[0]][doc]) + np.float32(per[order[1]][doc])) + others)
        ranked = sorted(exp.items(), key=lambda kv: (-kv[1], kv[0]))
        assert list(d) == [k for k, _ in ranked]
        assert np.array_equal(s, np.array([v for _, v in ranked], np.float32))


def test_stract_bm25_and_linear_combine():
    # core/src/ranking/bm25.rs:136-150 expression, tf==0 -> 0; initial.rs:79-93 f64 combine order
    w, cache = stract_bm25_weight(10, 1000, 20.0)
    L = oracle.lib()
    assert L.orc_stract_bm25_score(w, cache, 1.2, 5, 0) == 0.0
    tf, idn = 3, 7
    exp = np.float32(w) * (np.float32(np.float32(tf) * np.float32(np.float32(1.2) + np.float32(1.0))) / np.float32(np.float32(tf) + cache[idn]))
    assert L.orc_stract_bm25_score(w, cache, 1.2, idn, tf) == exp
    rng = np.random.default_rng(3)
    seg, dfs, avg = _random_index(rng, 3000, 8, 900)
    q = np.array([0, 3, 5, 6, 7], np.uint32)
    wc = [stract_bm25_weight(dfs[t], 3000, avg) for t in q]
    weights = np.array([x for x, _ in wc], np.float32); caches = np.stack([c for _, c in wc])
    sig = [rng.random(3000) ** 8, rng.random(3000)]
    coeffs = [2.0, 0.02]
    d, tot, scored = seg.signal_topk(q, weights, caches, 1.2, 0.005, sig, coeffs, 50)
    # recompute by brute force in numpy with the same rounding sequence
    tfs = np.zeros((len(q), 3000), np.uint32)
    for i, t in enumerate(q):
        dd, _, _ = seg.topk([t], [np.float32(1)], np.ones((1, 256), np.float32), 2, 3000)
        c = seg.cursor(t, 1.0, np.ones(256, np.float32))
        while c.doc() != TERMINATED:
            tfs[i, c.doc()] = c.tf(); c.advance()
    cand = np.nonzero(tfs.sum(0))[0]
    assert scored == len(cand)
    totals = {}
    for doc in cand:
        bm = np.float32(0.0)
        for i in range(len(q)):
            bm = np.float32(bm + np.float32(L.orc_stract_bm25_score(weights[i], caches[i], 1.2, seg.fieldnorm_ids[doc], int(tfs[i, doc]))))
        total = 0.0 + 0.005 * float(bm)
        for j in range(2):
            total = total + coeffs[j] * sig[j][doc]
        totals[int(doc)] = total
    ranked = sorted(totals.items(), key=lambda kv: (-kv[1], kv[0]))[:50]
    assert list(d) == [k for k, _ in ranked]
    assert np.array_equal(tot, np.array([v for _, v in ranked]))


def test_term_info_store_reference_kats():
    """tantivy/src/termdict/fst_termdict/term_info_store.rs:293-308 (test_bitpacked) and :330-357 (test_pack), the
    compute_num_bits values quoted there, and the fixed 47-byte TermInfoBlockMeta (:50-53)."""
    from oracle import bitpack, extract_bits, term_info_store_get, term_info_store_write
    buf = bitpack([321, 2, 51], [9, 2, 6])
    assert len(buf) == 3
    assert extract_bits(buf, 0, 9) == 321 and extract_bits(buf, 9, 2) == 2 and extract_bits(buf, 11, 6) == 51
    off = lambda i: i * 13 + i * i   # noqa: E731
    n = 1000
    df = np.arange(n, dtype=np.uint32)
    ps = np.array([off(i) for i in range(n)], np.uint64); pe = np.array([off(i + 1) for i in range(n)], np.uint64)
    store = term_info_store_write(df, ps, pe, ps * 3, pe * 3)
    meta_len = int(store[:8].view(np.uint64)[0])
    assert int(store[8:16].view(np.uint64)[0]) == n and meta_len == 47 * ((n + 255) // 256)
    for i in range(n):
        assert term_info_store_get(store, i) == (i, off(i), off(i + 1), off(i) * 3, off(i + 1) * 3), i


def test_stract_bm25_idf_scaling_reference_case():
    """core/src/ranking/bm25.rs:156-177 `test_bm25_idf_scaling`: MultiBm25Weight over ('the': df 98, 'end': df 20) of 100 docs,
    avg_fieldnorm 1.0; a doc with tf (8, 13) must outscore one with tf (15, 10), fieldnorm id 0 on both."""
    L = oracle.lib()
    ws = [stract_bm25_weight(98, 100, 1.0), stract_bm25_weight(20, 100, 1.0)]

    def score(stats):   # MultiBm25Weight::score: f32 sum over the terms in order
        s = np.float32(0.0)
        for (idn, tf), (w, cache) in zip(stats, ws):
            s = np.float32(s + np.float32(L.orc_stract_bm25_score(w, cache, 1.2, idn, tf)))
        return s
    high_the, high_end = score([(0, 15), (0, 10)]), score([(0, 8), (0, 13)])
    assert high_end > high_the
