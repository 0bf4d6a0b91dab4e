"""Seeded synthetic inputs of the shapes BASELINE.json names.  Host-side numpy
versions; bench.py uses the device-side generators of the library for the 1B-edge graph."""
import numpy as np

MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)
NOFOLLOW = 1 << 8  # RelFlags::NOFOLLOW, crates/core/src/webpage/html/links.rs:124


def splitmix64(seed, i):
    """Counter-based splitmix64: the i-th output of the stream seeded with `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (np.asarray(i, np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def node_ids(idx, id_seed=7):
    """128-bit id of node index j: hi = splitmix64(7, 2j), lo = splitmix64(7, 2j+1)."""
    idx = np.asarray(idx, np.uint64)
    return splitmix64(id_seed, 2 * idx + 1), splitmix64(id_seed, 2 * idx)  # (lo, hi)


def rel_flags(n_edges, flag_seed=9):
    i = np.arange(n_edges, dtype=np.uint64)
    return np.where(splitmix64(flag_seed, i) % np.uint64(10) == 0, np.uint64(NOFOLLOW), np.uint64(0))


def edges_from_indices(fi, ti, flag_seed=9, id_seed=7):
    flo, fhi = node_ids(fi, id_seed)
    tlo, thi = node_ids(ti, id_seed)
    return dict(from_lo=flo, from_hi=fhi, to_lo=tlo, to_hi=thi, rel_flags=rel_flags(len(fi), flag_seed))


def uniform_graph(n_nodes=100_000, n_edges=1_000_000, seed=42):
    """Config C1: edge i = (splitmix64(seed,2i) mod N, splitmix64(seed,2i+1) mod N); 10 % NOFOLLOW.
    Self-loops and duplicates are left in the stream (the library and the oracle drop them)."""
    i = np.arange(n_edges, dtype=np.uint64)
    fi = splitmix64(seed, 2 * i) % np.uint64(n_nodes)
    ti = splitmix64(seed, 2 * i + 1) % np.uint64(n_nodes)
    return edges_from_indices(fi, ti)


def rmat_indices(n_nodes, n_edges, seed=42, scale=26, first=0):
    """R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05), `scale` levels, folded mod N.  Level l of edge i uses
    16 bits of splitmix64(seed, 7*i + l//4): quadrant thresholds 37356 / 49807 / 62259 of 65536."""
    i = np.arange(first, first + n_edges, dtype=np.uint64)
    f = np.zeros(n_edges, np.uint64)
    t = np.zeros(n_edges, np.uint64)
    for w in range((scale + 3) // 4):
        r = splitmix64(seed, np.uint64(7) * i + np.uint64(w))
        for k in range(4):
            lvl = w * 4 + k
            if lvl >= scale:
                break
            x = (r >> np.uint64(16 * k)) & np.uint64(0xFFFF)
            fb = (x >= np.uint64(49807)).astype(np.uint64)
            tb = (((x >= np.uint64(37356)) & (x < np.uint64(49807))) | (x >= np.uint64(62259))).astype(np.uint64)
            f = (f << np.uint64(1)) | fb
            t = (t << np.uint64(1)) | tb
    return f % np.uint64(n_nodes), t % np.uint64(n_nodes)


def rmat_graph(n_nodes, n_edges, seed=42, scale=None):
    """Config C2/C3 shape at any size (scale defaults to ceil(log2 N))."""
    if scale is None:
        scale = max(1, int(np.ceil(np.log2(max(n_nodes, 2)))))
    fi, ti = rmat_indices(n_nodes, n_edges, seed, scale)
    return edges_from_indices(fi, ti)
