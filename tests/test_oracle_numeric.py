"""Hand-derived known answers for the oracle's numeric-signal transforms (core/src/ranking/signals/core/non_text.rs has no
test of its own for them: parity of these transforms is pinned on the expressions, not on reference vectors)."""
import numpy as np

import oracle

NOW = 1_700_000_000


def one(which, x, **kw):
    return float(oracle.numeric_scores(which, [x], **kw)[0])


def test_score_rank_points():
    # (10 - log_8(1 + rank)).max(0): rank 0 -> 10, 7 -> 9, 63 -> 8 (ln 64 / ln 8 == 2 in f64), far beyond 8^10 -> clamped to 0
    assert one(1, 0) == 10.0 and one(1, 7) == 9.0 and one(1, 63) == 8.0
    assert one(1, 8 ** 11) == 0.0
    r = oracle.numeric_scores(1, np.arange(0, 5000, dtype=np.uint64))
    assert np.all(np.diff(r) <= 0) and r[-1] > 5.0          # monotone, log-slow


def test_inverse_fetch_time_and_link_density_points():
    assert one(4, 0) == 1.0 and one(4, 1) == 0.5 and one(4, 3) == 0.25
    assert one(5, 0) == 1.0 and one(5, 999) == 1.0 / 1000.0 and one(5, 1000) == 0.0 and one(5, 10 ** 9) == 0.0   # cache of 1000 entries
    assert one(7, 0.0) == 1.0 and one(7, 0.5) == 0.5 and one(7, 0.5000001) == 0.0 and one(7, 0.25) == 0.75
    assert one(2, 1) == 1.0 and one(2, 0) == 0.0 and one(3, 1) == 0.0 and one(3, 0) == 1.0                         # IsHomepage / HasAds


def test_update_timestamp_points():
    # future or equal -> 0; < 1 h old -> hours 0 -> 72/72; exactly 72 h -> 0.5; the cache ends at 3 * 365 * 24 hours; no clock -> 0
    assert one(6, NOW, now=NOW) == 0.0 and one(6, NOW + 5, now=NOW) == 0.0
    assert one(6, NOW - 1, now=NOW) == 1.0 and one(6, NOW - 3599, now=NOW) == 1.0
    assert one(6, NOW - 3600, now=NOW) == 72.0 / 73.0
    assert one(6, NOW - 72 * 3600, now=NOW) == 0.5
    last = 3 * 365 * 24 - 1
    assert one(6, NOW - last * 3600, now=NOW) == 72.0 / (last + 72.0) and one(6, NOW - (last + 1) * 3600, now=NOW) == 0.0
    assert one(6, NOW - 100, now=None) == 0.0


def test_region_points():
    counts, total = [30, None, 10], 40
    kw = dict(region_counts=counts, region_total=total)
    assert one(8, 0, **kw) == 0.75 and one(8, 1, **kw) == 0.0 and one(8, 2, **kw) == 0.25 and one(8, 7, **kw) == 0.0
    assert one(8, 2, selected=2, **kw) == 50.25 and one(8, 0, selected=2, **kw) == 0.75
    assert one(8, 2, selected=2) == 0.0                      # no RegionCount: the signal is 0, boost included
