"""Committed golden fixtures (tests/golden/*.json; generator: tests/golden/make_golden.py).

CPU: the oracle still reproduces every frozen number (so the checker cannot drift unnoticed), and the list of reference
known-answer values stays in step with the oracle tests.  GPU (tests/test_round1_late_gpu.py): the CUDA path reproduces the frozen numbers
directly through the C ABI -- no oracle call at run time.  The same two functions also run on the CPU SIMT emulator
(tests/test_hyperball_emulated.py, tests/test_bm25_emulated.py)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)


def _load(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def test_oracle_reproduces_golden_fixtures():
    import make_golden as M
    assert json.loads(json.dumps(M.path1())) == _load("path1_small.json")
    assert json.loads(json.dumps(M.path2())) == _load("path2_small.json")
    assert json.loads(json.dumps(M.path1_c1())) == _load("path1_c1.json")


def test_reference_kat_list_matches_oracle_tests():
    k = _load("reference_kats.json")
    src = open(os.path.join(HERE, "test_oracle_path1.py")).read() + open(os.path.join(HERE, "test_oracle_path2.py")).read()
    assert repr(k["kahan_sum"]["sum"]) in src
    for v in k["term_scorer_f32"]["values"] + k["top_docs_droopy_tax"]["scores"]:
        assert repr(v) in src, v
    import oracle
    s = oracle.KahanSum()
    for x in k["kahan_sum"]["inputs"]:
        s.add(x)
    assert s.sum == k["kahan_sum"]["sum"]


# ---- the CUDA path against the frozen numbers (no oracle involved) ----------------------------------------------
def check_path1_against_golden():
    import make_golden as M
    from stract_b200.webgraph import DeviceGraph, HarmonicCentrality, Webgraph
    g = _load("path1_small.json")
    d = M.path1_inputs()
    a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    assert M.sha(np.stack(a)) == g["input_sha256"]
    graph = Webgraph.from_arrays(*a)
    dg = DeviceGraph(graph)
    try:
        assert dg.info()["n_nodes"] == g["n_nodes"]
        for it in g["per_iteration"]:
            st = dg.step()
            assert M.sha(dg.registers()) == it["registers_sha256"]
            assert (st["n_changed"] > 0) == bool(it["changed"])
    finally:
        dg.close()
    r = HarmonicCentrality.calculate(graph, with_ranks=True, top=64)
    assert r.iterations == g["iterations"] and len(r.values) == g["n_positive"]
    assert M.sha(r.ids_lo) == g["ids_lo_sha256"] and M.sha(r.ids_hi) == g["ids_hi_sha256"]
    assert M.f64hex(r.values[:16]) == g["centrality_head"] and M.sha(r.values) == g["centrality_sha256"]
    rlo, rhi = r.rank_ids
    order = np.array(g["rank_order_head"])
    assert np.array_equal(rlo[:32], r.ids_lo[order]) and np.array_equal(rhi[:32], r.ids_hi[order])


def check_c1_against_golden():
    """BASELINE configs[0] at full size against the frozen hashes."""
    import make_golden as M
    from stract_b200 import synth
    from stract_b200.webgraph import DeviceGraph, Webgraph
    g = _load("path1_c1.json")
    d = synth.uniform_graph(100_000, 1_000_000, 42)
    a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    assert M.sha(np.stack(a)) == g["input_sha256"]
    dg = DeviceGraph(Webgraph.from_arrays(*a))
    try:
        iters, _ = dg.run(20)
        lo, hi, c = dg.result()
        assert iters == g["iterations"] and len(c) == g["n_positive"]
        assert M.sha(dg.registers()) == g["registers_sha256"]
        assert M.sha(lo) == g["ids_lo_sha256"] and M.sha(hi) == g["ids_hi_sha256"]
        assert M.f64hex(c[:8]) == g["centrality_head"] and M.sha(c) == g["centrality_sha256"]
    finally:
        dg.close()


def check_path2_against_golden():
    import make_golden as M
    from stract_b200 import bm25
    from stract_b200.bm25 import MODE_AND, MODE_OR, SegmentReader, SignalComputer, SignalTable, TopDocs
    g = _load("path2_small.json")
    lens, td, tt, cols = M.path2_inputs()
    ids = bm25.fieldnorms_to_ids(lens)
    avg = np.array([int(g["avg_fieldnorm"], 16)], np.uint32).view(np.float32)[0]
    for rec in (2, 1):
        data, infos = bm25.encode_postings(td, tt, ids, avg, record_option=rec)
        assert M.sha(data) == g[f"postings_sha256_record{rec}"]
    seg = SegmentReader(data, infos, ids)
    try:
        assert np.float32(seg.average_fieldnorm) == avg
        comp = SignalComputer(seg, SignalTable(cols), [2.0, 0.5], coeff_text=0.005)
        for qi, q in enumerate(g["queries"]):
            for name, mode in (("and", MODE_AND), ("or", MODE_OR)):
                want = g["results"][name][qi]
                if want is None:
                    continue
                got = TopDocs.with_limit(50).search(seg, q, mode)
                assert [d_ for _, d_ in got] == want["docs"], (name, q)
                assert M.f32hex([s for s, _ in got]) == want["scores"], (name, q)
            want = g["results"]["signal"][qi]
            dd, tot, n = comp.top_docs_batch(np.array([q], np.uint32), 50)
            assert [int(x) for x in dd[0, :n[0]]] == want["docs"] and M.f64hex(tot[0, :n[0]]) == want["totals"], ("signal", q)
    finally:
        seg.close()
