"""Pins the path-1 oracle against every known-answer test the reference holds for it
(SURVEY.md 8c).  CPU only."""
import math
import os

import numpy as np

import oracle
from oracle import Bloom, DenseHyperBall, Hll, KahanSum, hyperball_faithful
from stract_b200.webgraph import RelFlags, SKIPPED_REL


def test_kahan_kat():
    # crates/core/src/kahan_sum.rs:86-104
    k = KahanSum()
    assert k.sum == 0.0
    for x in [10000.0, math.pi, math.e, math.pi, math.e, math.pi, math.e]:
        k.add(x)
    assert k.sum == 10017.579623446147


def test_bloom_kat():
    # crates/bloom/src/lib.rs:197-216
    bf = Bloom(100, 0.01)
    for x in (1, 2, 3, 4, 5):
        bf.insert(x)
    assert all(bf.contains(x) for x in (1, 2, 3, 4, 5))
    assert not any(bf.contains(x) for x in (6, 7, 8, 9, 10))


def test_bloom_estimate_card_truncation():
    # lib.rs:108-123: `.ln() as i64` binds before the multiplication -> 0 until fill >= 63.2 %
    bf = Bloom(1000, 0.05)
    nbits = oracle.lib().orc_bloom_num_bits(1000, 0.05)
    assert nbits == math.ceil(1000 * math.log(0.05) / (-8.0 * math.log(2.0) ** 2))
    assert bf.estimate_card() == 0
    for x in range(200):
        bf.insert(x)
    assert bf.estimate_card() == 0


def _size_bounds(h):
    size = h.size()
    delta = int((1.04 / math.sqrt(h.n)) * 2.0 * size)
    return size - delta, size + delta


def test_hll128_many_sizes_and_merge():
    # crates/core/src/hyperloglog.rs:4566-4599 (N=128)
    h = Hll(128)
    for i in range(10_000):
        h.add(i)
    lo, hi = _size_bounds(h)
    assert lo < h.size() < hi
    wm, a, b = Hll(128), Hll(128), Hll(128)
    for i in range(10_000):
        wm.add(i); a.add(i)
    for i in range(10_001, 20_000):
        wm.add(i); b.add(i)
    a.merge(b)
    assert np.array_equal(a.registers, wm.registers)


def test_hll128_ten_million():
    # hyperloglog.rs:4553-4564: 10M inserts stay inside size_bounds
    L = oracle.lib()
    regs = np.zeros(128, np.uint8)
    L.orc_hll_add_range(regs, 128, 0, 10_000_000)
    size = int(L.orc_hll_size(regs, 128))
    delta = int((1.04 / math.sqrt(128)) * 2.0 * size)
    assert size > 0 and size - delta < size < size + delta


def test_hll64_small_exact():
    # linear counting with distinct registers truncates to the exact count for tiny sets
    h = Hll(64)
    seen = set()
    n = 0
    for i in range(1, 200):
        hsh = (i * 11400714819323198549) & ((1 << 64) - 1)
        j = hsh >> 58
        if j in seen:
            continue
        seen.add(j); h.add(i); n += 1
        if n <= 4:
            assert h.size() == n
    assert h.size() > 4


def _ids(names):
    # arbitrary distinct u128 ids (the reference hashes names with xxh3; orderings do not depend on it)
    return {n: (0xABCDEF0000 + 7919 * (i + 1)) | ((i + 1) << 64) for i, n in enumerate(names)}


def _soa(edges):
    n = len(edges)
    a = [np.zeros(n, np.uint64) for _ in range(5)]
    m = (1 << 64) - 1
    for i, (f, t, r) in enumerate(edges):
        a[0][i] = f & m; a[1][i] = f >> 64; a[2][i] = t & m; a[3][i] = t >> 64; a[4][i] = r
    return a


def _as_map(res):
    return {(int(h) << 64) | int(l): float(c) for l, h, c in zip(res["ids_lo"], res["ids_hi"], res["centrality"])}


def _test_edges(ids, flag=0):
    return [(ids["A"], ids["B"], flag), (ids["B"], ids["C"], flag), (ids["A"], ids["C"], flag),
            (ids["C"], ids["A"], flag), (ids["D"], ids["C"], flag)]


def test_harmonic_orderings_and_hand_kat():
    # harmonic.rs:478-493: C > A > B, D absent; hand-derived values
    ids = _ids("ABCD")
    for impl in ("faithful", "dense"):
        if impl == "faithful":
            m = _as_map(hyperball_faithful(*_soa(_test_edges(ids))))
        else:
            d = DenseHyperBall(*_soa(_test_edges(ids)))
            d.run()
            m = _as_map(d.result())
        assert m[ids["C"]] > m[ids["A"]] > m[ids["B"]]
        assert ids["D"] not in m
        assert m[ids["C"]] == 3.0 / 3.0
        assert abs(m[ids["A"]] - (1 + 0.5 + 0.5) / 3) < 1e-15
        assert abs(m[ids["B"]] - (1 + 0.5 + 1.0 / 3) / 3) < 1e-15


def test_host_harmonic_centrality():
    # harmonic.rs:359-476: B.com (2 in-links from distinct hosts) > A.com (only self links)
    ids = _ids(["A.com", "B.com", "C.com", "D.com"])
    e = [(ids["A.com"], ids["A.com"], 0)] * 12 + [(ids["C.com"], ids["B.com"], 0), (ids["D.com"], ids["B.com"], 0)]
    m = _as_map(hyperball_faithful(*_soa(e)))
    assert m[ids["B.com"]] > m.get(ids["A.com"], 0.0)


def test_additional_edges_ignored():
    # harmonic.rs:495-553: duplicates of an edge across commits give an identical map
    ids = _ids("ABCD")
    base = _as_map(hyperball_faithful(*_soa(_test_edges(ids))))
    extra = _test_edges(ids) + [(ids["A"], ids["B"], 0)] * 8
    assert _as_map(hyperball_faithful(*_soa(extra))) == base


def test_rel_flags_ignored():
    # harmonic.rs:555-602: all-TAG / all-SAME_ICANN_DOMAIN edges => no positive centrality
    ids = _ids("ABCD")
    for flag in (RelFlags.TAG, RelFlags.SAME_ICANN_DOMAIN):
        assert flag & SKIPPED_REL
        res = hyperball_faithful(*_soa(_test_edges(ids, flag)))
        assert res["n_nodes"] == 4 and len(res["centrality"]) == 0


def test_first_occurrence_decides_flags():
    # store.rs:313 unique_by keeps the first (from,to): a skipped first copy hides a later clean one
    ids = _ids("AB")
    skipped_first = [(ids["A"], ids["B"], RelFlags.NOFOLLOW), (ids["A"], ids["B"], 0)]
    clean_first = [(ids["A"], ids["B"], 0), (ids["A"], ids["B"], RelFlags.NOFOLLOW)]
    for impl in (hyperball_faithful, lambda *a: (lambda d: (d.run(), d.result())[1])(DenseHyperBall(*a))):
        assert len(impl(*_soa(skipped_first))["centrality"]) == 0
        assert len(impl(*_soa(clean_first))["centrality"]) == 1


def test_faithful_equals_dense_on_random_graphs():
    from stract_b200 import synth
    for n, e, seed in ((50, 200, 1), (300, 2000, 2), (2000, 6000, 3)):
        g = synth.uniform_graph(n, e, seed)
        a = (g["from_lo"], g["from_hi"], g["to_lo"], g["to_hi"], g["rel_flags"])
        f = hyperball_faithful(*a)
        d = DenseHyperBall(*a, threads=2)
        d.run()
        r = d.result()
        assert f["iters"] == r["iters"] and f["n_nodes"] == r["n_nodes"]
        assert np.array_equal(f["ids_lo"], r["ids_lo"]) and np.array_equal(f["ids_hi"], r["ids_hi"])
        assert np.array_equal(f["centrality"], r["centrality"])  # bit-exact


def test_harmonic_rank_order_restatement():
    """store_harmonic sorts (Reverse(SortableFloat(c)), node_id): centrality descending, equal centralities by
    ascending id; top_nodes takes the largest (c, id) pairs, i.e. equal centralities by DESCENDING id
    (crates/core/src/webgraph/centrality/mod.rs:17-37,88-108; SortableFloat = f64::total_cmp, core/src/lib.rs:259-262)."""
    import numpy as np
    from oracle import harmonic_ranks
    lo = np.array([5, 1, 9, 3, 7], np.uint64); hi = np.array([0, 2, 0, 0, 2], np.uint64)
    c = np.array([0.5, 0.25, 0.5, 1.0, 0.25])
    ids = [(int(h) << 64) | int(l) for l, h in zip(lo, hi)]
    want = sorted(range(5), key=lambda i: (-c[i], ids[i]))
    assert list(harmonic_ranks(lo, hi, c)) == want == [3, 0, 2, 1, 4]
    want_top = sorted(range(5), key=lambda i: (c[i], ids[i]), reverse=True)
    assert list(harmonic_ranks(lo, hi, c, ties_desc=True)) == want_top == [3, 2, 0, 4, 1]


def test_hll64_bias_does_not_depend_on_the_std_binary_search():
    """estimate_bias binary-searches the precision-5 raw table, which is not sorted (inversions at 127/128 and 130/131),
    so the index it returns depends on the std implementation: Rust >= 1.82 (branch-free, what the oracle and the GPU
    follow) and Rust 1.52-1.81 (three-way compare) can land on different neighbours.  The value that matters is the mean
    bias of the 6 nearest entries, and a dense sweep over [20, 320] -- extra fine around the two inversions -- shows it is
    the same under both searches, so HyperLogLog<64>::size() does not depend on the toolchain."""
    import re
    t = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "sb200_hll_tables.h")).read()

    def arr(name):
        m = re.search(name + r"\[\w*\]\s*=\s*\{(.*?)\};", t, re.S)
        return np.array([float(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()])
    raw, bias = arr("SB200_HLL_RAW_P5"), arr("SB200_HLL_BIAS_P5")
    n = len(raw)
    assert [i for i in range(n - 1) if raw[i] > raw[i + 1]] == [127, 130]

    def new_search(e):
        size, base = n, 0
        while size > 1:
            half = size // 2; mid = base + half
            base = base if raw[mid] > e else mid
            size -= half
        r = base if raw[base] == e else base + (1 if raw[base] < e else 0)
        return min(r, n - 1)

    def old_search(e):
        left, right, size = 0, n, n
        while left < right:
            mid = left + size // 2
            if raw[mid] < e: left = mid + 1
            elif raw[mid] > e: right = mid
            else: return mid
            size = right - left
        return min(left, n - 1)

    def mean_bias(idx, e):
        il, ir, acc = idx, (idx + 1 if idx < n - 1 else -1), 0.0
        for _ in range(6):
            if il >= 0 and ir >= 0:
                right = abs(raw[ir] - e) < abs(raw[il] - e)
            else:
                right = il < 0
            i = ir if right else il
            acc += bias[i]
            if right: ir = i + 1 if i < n - 1 else -1
            else: il = i - 1 if i > 0 else -1
        return acc / 6.0
    sweep = np.concatenate([np.linspace(20, 320, 60001), np.linspace(128.2, 128.5, 60001), np.linspace(130.9, 131.2, 60001),
                            raw, np.nextafter(raw, 0), np.nextafter(raw, 1e9)])
    differing_index = 0
    for e in sweep:
        a, b = new_search(e), old_search(e)
        if a != b:
            differing_index += 1
            assert mean_bias(a, e) == mean_bias(b, e), e
    assert differing_index > 0     # the searches do disagree on the index somewhere -- just never on the result
