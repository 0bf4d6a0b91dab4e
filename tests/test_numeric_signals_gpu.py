"""The numeric CoreSignals' value -> score transforms built by the library from RAW fast-field columns
(sb200_signals_create_raw; core/src/ranking/signals/core/non_text.rs): every score bit-equal to the oracle restatement,
including the edge values of each transform, and the table used end to end in the multi-field signal program."""
import numpy as np
import pytest

import oracle
from stract_b200.bm25 import NUMERIC_SIGNALS, RawSignalTable

pytestmark = pytest.mark.gpu

NOW = 1_700_000_000
WHICH = {"HostCentrality": 0, "PageCentrality": 0, "HostCentralityRank": 1, "PageCentralityRank": 1, "IsHomepage": 2, "HasAds": 3,
         "TrackerScore": 4, "UrlDigits": 4, "UrlSlashes": 4, "FetchTimeMs": 5, "UpdateTimestamp": 6, "LinkDensity": 7, "Region": 8}


def raw_columns(n, seed=3):
    rng = np.random.default_rng(seed)
    ts = NOW - rng.integers(-5000, 4 * 365 * 24 * 3600, n)            # future, fresh, beyond the 3-year cache
    ts[:6] = [NOW, NOW - 1, NOW - 3599, NOW - 3600, NOW - 3 * 365 * 24 * 3600, NOW - 3 * 365 * 24 * 3600 + 1]
    rank = rng.integers(0, 2 ** 40, n).astype(np.uint64)
    rank[:5] = [0, 1, 7, 8 ** 10 - 2, 8 ** 10]                            # score_rank hits 10, ..., exactly 0, clamped
    fetch = rng.integers(0, 3000, n).astype(np.uint64)
    fetch[:3] = [0, 999, 1000]
    dens = rng.random(n)
    dens[:3] = [0.5, 0.5000000001, 0.0]
    return {"HostCentrality": rng.random(n) ** 8, "HostCentralityRank": rank, "PageCentrality": rng.random(n) ** 3,
            "PageCentralityRank": rng.integers(0, 10 ** 9, n).astype(np.uint64), "IsHomepage": rng.integers(0, 2, n).astype(np.uint8),
            "FetchTimeMs": fetch, "UpdateTimestamp": np.maximum(ts, 0).astype(np.uint64),
            "TrackerScore": rng.integers(0, 40, n).astype(np.uint64), "Region": rng.integers(0, 9, n).astype(np.uint64),
            "UrlDigits": rng.integers(0, 30, n).astype(np.uint64), "UrlSlashes": rng.integers(0, 12, n).astype(np.uint64),
            "LinkDensity": dens, "HasAds": rng.integers(0, 2, n).astype(np.uint8)}


def want_column(name, raw, now, region_count, selected):
    counts, total = region_count if region_count is not None else (None, 0)
    return oracle.numeric_scores(WHICH[name], raw, now=now, region_counts=counts, region_total=total, selected=selected)


@pytest.mark.parametrize("now,region_count,selected", [(NOW, ([120, None, 30, 0, 77, 1, 5], 233), 2), (None, None, None), (NOW, ([5, 5, 5], 15), None)])
def test_every_numeric_signal_bit_exact(now, region_count, selected):
    n = 20_000
    cols = raw_columns(n)
    tab = RawSignalTable(cols, current_timestamp=now, region_count=region_count, selected_region=selected)
    try:
        got = tab.read()
        assert [name for name, _, _ in tab.numeric] == [s[0] for s in NUMERIC_SIGNALS]      # CoreSignalEnum order
        assert [c for _, c, _ in tab.numeric] == list(range(13))
        for name, col, coef in tab.numeric:
            want = want_column(name, cols[name], now, region_count, selected)
            assert got[:, col].tobytes() == want.tobytes(), (name, np.flatnonzero(got[:, col] != want)[:5])
        # a window of the table
        assert np.array_equal(tab.read(100, 50), got[100:150])
    finally:
        tab.close()


def test_subset_of_signals_in_enum_order():
    n = 5_000
    cols = raw_columns(n, seed=9)
    sub = {k: cols[k] for k in ("FetchTimeMs", "HostCentrality", "HasAds")}
    tab = RawSignalTable(sub)
    try:
        assert [name for name, _, _ in tab.numeric] == ["HostCentrality", "FetchTimeMs", "HasAds"]
        got = tab.read()
        for name, col, _ in tab.numeric:
            assert got[:, col].tobytes() == want_column(name, cols[name], None, None, None).tobytes()
    finally:
        tab.close()
    with pytest.raises(KeyError):
        RawSignalTable({"Bm25Title": cols["HostCentrality"]})
