"""Multi-field recall-stage signals through the C ABI against the CPU oracle: BM25 per field, Bm25F over
the fields, coverage, idf_sum, numeric columns, n-gram dampening chains, in SignalComputeOrder order; f64 totals and doc
order bit-exact.  Also the host-side order / coefficient quirks of the reference (computer/order.rs, mod.rs:300-389)."""
import numpy as np
import pytest

import oracle
from stract_b200 import bm25
from stract_b200.bm25 import NO_TERM, MultiFieldSignalComputer, SignalComputeOrder, SignalTable

pytestmark = pytest.mark.gpu

FIELDS = ["Title", "CleanBody", "Url", "TitleBigrams", "TitleTrigrams"]
ENABLED = {"Bm25F", "Bm25Title", "TitleCoverage", "Bm25TitleBigrams", "Bm25TitleTrigrams", "Bm25CleanBody", "CleanBodyCoverage", "IdfSumUrl"}


def _field_index(seed, max_doc, dfs, mean_len):
    import test_bm25_gpu as T
    rng = np.random.default_rng(seed)
    lens = np.maximum(1, rng.lognormal(mean_len, 0.7, max_doc)).astype(np.uint32)
    td, tt = [], []
    for df in dfs:
        td.append(np.sort(rng.choice(max_doc, df, replace=False)).astype(np.uint32))
        tt.append(np.minimum(rng.geometric(0.6, df), 255).astype(np.uint32))
    return T.build(td, tt, lens)


def make_segment(seed=5, max_doc=30_000):
    dfs = {"Title": [40, 300, 129, 2500, 7], "CleanBody": [900, 6000, 128, 15000, 3000, 1], "Url": [20, 500, 4000],
           "TitleBigrams": [30, 260, 1000], "TitleTrigrams": [12, 200]}
    mean = {"Title": 2.0, "CleanBody": 5.0, "Url": 1.5, "TitleBigrams": 1.8, "TitleTrigrams": 1.6}
    pairs = {f: _field_index(seed + i, max_doc, dfs[f], mean[f]) for i, f in enumerate(FIELDS)}
    return pairs, dfs


def random_queries(rng, dfs, names, nq, ns):
    """Slots in the order prepare_textfields builds them per field; unknown terms mixed in, padding behind."""
    sf = np.full((nq, ns), 0xFF, np.uint8); st = np.full((nq, ns), NO_TERM, np.uint32)
    for q in range(nq):
        x = 0
        for fi, name in enumerate(names):
            n_terms = int(rng.integers(0, 4))
            for _ in range(n_terms):
                if x >= ns:
                    break
                sf[q, x] = fi
                st[q, x] = NO_TERM if rng.random() < 0.15 else int(rng.integers(0, len(dfs[name])))
                x += 1
    return sf, st


def check_against_oracle(comp, pairs, cols, sf, st, k, numeric, boost=None, dfa=None):
    docs, totals, n_out = comp.top_docs_batch(sf, st, k, slot_boost=boost, doc_freq_all_body=dfa)
    names = comp.names
    # the vectorised idf of the host mirror against the scalar f32 expression (tantivy bm25.rs:52-56), slot by slot
    from stract_b200.bm25 import idf
    for q in range(sf.shape[0]):
        for x in range(sf.shape[1]):
            f = int(sf[q, x])
            if f == 0xFF or f & 0x80:
                assert comp.last_inputs["idf"][q, x] == 0 and comp.last_inputs["idf_f"][q, x] == 0
                continue
            r = comp.readers[f]
            df = int(r.doc_freq[st[q, x]]) if st[q, x] != NO_TERM and st[q, x] < r.n_terms else 0
            assert comp.last_inputs["idf"][q, x] == idf(df, r.max_doc)
            assert comp.last_inputs["idf_f"][q, x] == idf(df if dfa is None else int(dfa[q][x]), r.max_doc)
    osegs = [pairs[n][0] for n in names]
    caches = comp.last_inputs["caches"]
    coefs = [np.float32(comp.field_coefficient(n)) for n in names]
    ops = [(kind, names.index(field) if field is not None else 0, chain, col, comp.coefficient(name, coef))
           for name, kind, field, chain, col, coef in comp.order.entries]
    for q in range(sf.shape[0]):
        od, ot = oracle.multi_signal_topk(osegs, caches, [1.2] * len(names), coefs, sf[q], st[q], comp.last_inputs["idf"][q],
                                          comp.last_inputs["idf_f"][q], ops, cols, k, None if boost is None else boost[q])
        n = int(n_out[q])
        assert n == len(od), (q, n, len(od))
        assert np.array_equal(docs[q, :n], od), (q, docs[q, :8], od[:8])
        assert np.array_equal(totals[q, :n], ot), (q, totals[q, :4], ot[:4])


def test_multi_field_signals_bit_exact():
    pairs, dfs = make_segment()
    max_doc = 30_000
    rng = np.random.default_rng(11)
    cols = [rng.random(max_doc) ** 6, 1.0 / (1.0 + rng.integers(0, 1000, max_doc).astype(np.float64))]
    numeric = [("HostCentrality", 0, 2.5), ("FetchTimeMs", 1, 0.001)]
    comp = MultiFieldSignalComputer({n: pairs[n][1] for n in FIELDS}, ENABLED, SignalTable(cols), numeric,
                                    coefficients={"Bm25Title": 0.02, "Bm25F": 0.3})
    assert comp.names == ["Title", "CleanBody", "Url", "TitleBigrams", "TitleTrigrams"]
    # <= 8 slots -> TMAX 8 kernel; 9..16 slots -> TMAX 16 kernel
    for ns, nq, k in ((8, 24, 50), (14, 16, 200)):
        sf, st = random_queries(rng, dfs, comp.names, nq, ns)
        check_against_oracle(comp, pairs, cols, sf, st, k, numeric)
    # WeightCache: the Bm25F idf from the AllBody doc_freq of the token instead of the field's own
    sf, st = random_queries(rng, dfs, comp.names, 8, 8)
    check_against_oracle(comp, pairs, cols, sf, st, 50, numeric, dfa=rng.integers(1, 20_000, sf.shape))
    # doc-range work items + merge: a batch dominated by one heavy query
    sf = np.full((12, 6), 0xFF, np.uint8); st = np.full((12, 6), NO_TERM, np.uint32)
    sf[0, :3] = [0, 1, 1]; st[0, :3] = [3, 3, 1]
    for q in range(1, 12):
        sf[q, :2] = [0, 2]; st[q, :2] = [4, 0]
    check_against_oracle(comp, pairs, cols, sf, st, 100, numeric)


def test_all_numeric_signals_from_raw_columns_in_the_program():
    """The 13 numeric CoreSignals built from raw fast-field columns (sb200_signals_create_raw) behind the text signals, default
    coefficients: totals bit-equal to the oracle program over the oracle's own transforms of the same raw columns."""
    import test_numeric_signals_gpu as N
    from stract_b200.bm25 import RawSignalTable
    pairs, dfs = make_segment(seed=33)
    max_doc = 30_000
    raw = N.raw_columns(max_doc, seed=4)
    rc = ([400, 10, None, 3, 90], 503)
    tab = RawSignalTable(raw, current_timestamp=N.NOW, region_count=rc, selected_region=3)
    comp = MultiFieldSignalComputer({n: pairs[n][1] for n in FIELDS}, ENABLED, tab, tab.numeric)
    assert [e[0] for e in comp.order.entries][-13:] == [n for n, _, _ in tab.numeric]
    cols = [N.want_column(name, raw[name], N.NOW, rc, 3) for name, _, _ in tab.numeric]
    rng = np.random.default_rng(8)
    sf, st = random_queries(rng, dfs, comp.names, 16, 8)
    check_against_oracle(comp, pairs, cols, sf, st, 120, None)
    tab.close()


def test_optic_rule_boosts_bit_exact():
    """SignalComputer::boosts (computer/mod.rs:471-497): rule docsets as probe-only slots, boosts and downranks, the
    1/(1+diff) branch, a rule that matches nothing, documents that only a rule holds (never candidates)."""
    pairs, dfs = make_segment(seed=21)
    max_doc = 30_000
    rng = np.random.default_rng(5)
    cols = [rng.random(max_doc)]
    comp = MultiFieldSignalComputer({n: pairs[n][1] for n in FIELDS}, ENABLED, SignalTable(cols), [("HostCentrality", 0, 1.0)])
    names = comp.names
    nq, ns = 20, 12
    sf = np.full((nq, ns), 0xFF, np.uint8); st = np.full((nq, ns), NO_TERM, np.uint32); bo = np.zeros((nq, ns), np.float64)
    for q in range(nq):
        x = 0
        for fi, name in enumerate(names[:3]):
            for _ in range(int(rng.integers(1, 3))):
                sf[q, x] = fi; st[q, x] = int(rng.integers(0, len(dfs[name]))); x += 1
        # rules on the Url / CleanBody fields; written BEFORE some text slots on purpose in every other query (the library reorders)
        rules = [(2, 2, 3.0), (1, 1, -1.5), (2, 1, -4.0), (1, 4, 0.5), (0, NO_TERM, 9.0)][: int(rng.integers(1, 6))]
        for f, t, b in rules:
            sf[q, x] = 0x80 | f; st[q, x] = t; bo[q, x] = b; x += 1
        if q % 2:
            perm = rng.permutation(x)
            sf[q, :x], st[q, :x], bo[q, :x] = sf[q, perm], st[q, perm], bo[q, perm]
            # the f32 sums follow slot order per field, the rule sums rule order: tell the oracle the same (stable) order
    check_against_oracle(comp, pairs, cols, sf, st, 100, None, boost=bo)
    # without rules the factor is exactly 1: passing slot_boost must not change a bit
    sf2 = sf.copy(); sf2[sf2 >= 0x80] = 0xFF; sf2[sf == 0xFF] = 0xFF
    d0, t0, n0 = comp.top_docs_batch(sf2, st, 100)
    d1, t1, n1 = comp.top_docs_batch(sf2, st, 100, slot_boost=bo)
    assert np.array_equal(d0, d1) and np.array_equal(t0, t1) and np.array_equal(n0, n1)


def test_coefficient_precedence_mirror():
    """SignalComputer::coefficient: query coefficients (entry or default) shadow the linear model entirely; without a query
    the model's weight, else the default.  Host logic only (no kernel launch)."""
    class _R:   # a reader stand-in: the constructor only stores it
        pass
    mk = lambda **kw: MultiFieldSignalComputer({"Title": _R()}, {"Bm25Title"}, **kw)
    lm = {"Bm25Title": 0.5, "TitleCoverage": 0.25}
    c = mk(coefficients={"Bm25Title": 0.02}, linear_model=lm)
    assert c.coefficient("Bm25Title", 0.0063) == 0.02 and c.coefficient("TitleCoverage", 0.01) == 0.01     # model ignored
    c = mk(linear_model=lm, has_query=False)
    assert c.coefficient("Bm25Title", 0.0063) == 0.5 and c.coefficient("Bm25F", 0.1) == 0.1
    c = mk(has_query=False)
    assert c.coefficient("Bm25Title", 0.0063) == 0.0063
    # TextFieldData.signal_coefficient follows the same rule: Title's last signal is TitleCoverage
    assert mk(linear_model=lm, has_query=False).field_coefficient("Title") == 0.25
    assert mk(coefficients={"TitleCoverage": 0.7}, linear_model=lm).field_coefficient("Title") == 0.7


def test_signal_compute_order_mirror():
    """SignalComputeOrder::new and the signal_coefficient quirk of prepare_textfields, host side only."""
    o = SignalComputeOrder(ENABLED | {"Bm25CleanBodyBigrams"}, numeric=[("HostCentrality", 0, 1.0)])
    assert [e[0] for e in o.entries] == ["Bm25TitleTrigrams", "Bm25TitleBigrams", "Bm25Title", "Bm25CleanBodyBigrams", "Bm25CleanBody",
                                         "Bm25F", "TitleCoverage", "CleanBodyCoverage", "IdfSumUrl", "HostCentrality"]
    assert [e[3] for e in o.entries][:5] == [1, 2, 2, 1, 2]
