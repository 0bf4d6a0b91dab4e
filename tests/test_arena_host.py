"""The opt-in device-memory arena (stract_b200/csrc/arena.h) is pure host logic over an abstract slab backend;
sb200_arena_selftest drives it over malloc'ed slabs, so this runs without a GPU: randomised alloc/free with
content checks (no two live blocks overlap), exact tiling of every slab, merging of free neighbours, per-stream
tags, steady state under a replayed allocation sequence, trim."""
import ctypes as C

import pytest

from stract_b200._lib import lib


@pytest.mark.parametrize("seed", [1, 2, 3, 12345])
def test_arena_selftest(seed):
    L = lib()
    L.sb200_arena_selftest.restype = C.c_int
    L.sb200_arena_selftest.argtypes = [C.c_uint64, C.c_uint32]
    assert L.sb200_arena_selftest(seed, 20000) == 0


def test_arena_stats_without_arena():
    L = lib()
    L.sb200_arena_stats.restype = C.c_int
    L.sb200_arena_stats.argtypes = [C.c_int] + [C.POINTER(C.c_uint64)] * 4
    r, u, p, s = (C.c_uint64(7) for _ in range(4))
    assert L.sb200_arena_stats(0, C.byref(r), C.byref(u), C.byref(p), C.byref(s)) == 0
    assert (r.value, u.value, p.value, s.value) == (0, 0, 0, 0)
