"""BucketCollector mirror (stract_b200/collector.py) pinned on the reference's own scenarios
(crates/core/src/collector/top_docs.rs:493-751), and the top-K harvest against the insert-everything harvest."""
import numpy as np

from stract_b200.collector import BucketCollector, CollectorConfig, Hashes, SimhashTable, harvest_top_k


def run(top_n, docs):
    c = BucketCollector(top_n, CollectorConfig())
    for (site, title, url, uwt, sim), doc, score in docs:
        c.insert(score, Hashes(site, title, url, uwt, sim), (score, doc))
    return c.into_sorted_vec(True)


def test_all_different():        # top_docs.rs:518-583
    docs = [((i, i, i, i, s), 122 + i, float(i)) for i, s in zip(range(1, 6), (12, 123, 1234, 12345, 123456))]
    assert run(3, docs) == [(5.0, 127), (4.0, 126), (3.0, 125)]


def test_less_than_topn():       # top_docs.rs:585-627
    docs = [((3, 3, 3, 3, 12), 125, 3.0), ((4, 4, 4, 4, 123), 126, 4.0), ((5, 5, 5, 5, 1234), 127, 5.0)]
    assert run(10, docs) == [(5.0, 127), (4.0, 126), (3.0, 125)]


def test_same_key_de_prioritised():   # top_docs.rs:629-705
    docs = [((1, 1, 1, 1, 12), 125, 3.0), ((2, 2, 2, 2, 123), 126, 3.1), ((2, 2, 2, 2, 1234), 127, 5.0)]
    assert run(10, docs) == [(5.0, 127), (3.0, 125), (3.1, 126)]
    assert run(2, docs) == [(5.0, 127), (3.0, 125)]


def test_simhash_dedup():        # top_docs.rs:707-750
    docs = [((1, 1, 1, 1, 1234), 125, 3.0), ((2, 2, 2, 2, 1234), 126, 3.1), ((3, 3, 3, 3, 1), 127, 5.0)]
    assert run(10, docs) == [(5.0, 127), (3.1, 126), (3.0, 125)]


def test_simhash_table_is_hamming_ball():   # simhash.rs:69-135: K = 3
    t = SimhashTable()
    t.insert(0b1111_0000)
    assert t.contains(0b1111_0000) and t.contains(0b1111_0111) and not t.contains(0b1111_1111 ^ 0b1111_0000 ^ 0b1_0000_0000)


def test_harvest_over_top_k_equals_harvest_over_everything():
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = 5000
        totals = rng.random(n) ** 3 * 10.0
        site = rng.integers(0, 40 if trial % 2 else 2000, n)      # few sites: heavy de-ranking, K has to grow
        url = rng.integers(0, 1 << 40, n); title = rng.integers(0, 3000, n) + (1 << 41)
        sim = rng.integers(1, 1 << 62, n)
        hashes = [Hashes(int(site[i]) + (1 << 50), int(title[i]), int(url[i]), int(url[i]) ^ 1, int(sim[i])) for i in range(n)]
        order = np.lexsort((np.arange(n), -totals))

        def search(K):
            return order[:K], totals[order[:K]]
        top_n = 100
        full = BucketCollector(top_n, CollectorConfig())
        for d in order:                      # the reference's collector sees every document (ascending doc order there;
            full.insert(float(totals[d]), hashes[d], (int(d), float(totals[d])))   # scores are distinct, order is immaterial)
        want = full.into_sorted_vec(True)
        got, K, proven = harvest_top_k(top_n, search, lambda d: hashes[d], k_max=8192)
        assert proven and got == want, (trial, K)
