"""Several ranks of the sharded HyperBall path inside ONE emulated process: every rank is its own handle, the register
arrays and bitmaps are caller-owned numpy buffers handed over with sb200_hyperball_bind_state, and each rank's publish
targets are the other ranks' buffers by address (sb200_hyperball_set_publish_targets) -- the same calls the NVSwitch
multicast / symmetric-memory set-up makes, with unicast targets.  After every iteration all replicas must be identical
and equal to the oracle's registers.  Started by tests/test_hyperball_emulated.py."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np  # noqa: E402
from stract_b200 import _lib, synth  # noqa: E402
L = _lib.declare(C.CDLL(os.path.join(HERE, "libsb200_emu.so")))
_lib._LIB = L
from stract_b200.webgraph import DeviceGraph, Webgraph
from stract_b200._lib import check
from oracle import DenseHyperBall, hyperball_faithful
def aligned(nbytes, dtype):
    raw = np.zeros(nbytes + 64, np.uint8); off = (-raw.ctypes.data) % 64
    return raw[off:off + nbytes].view(dtype), raw
for world in (2, 3):
    d = synth.rmat_graph(3000, 40000, seed=7)
    g = Webgraph.from_arrays(d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    hs = [DeviceGraph(g, device=0, rank=r, world_size=world) for r in range(world)]
    rb, bb = C.c_uint64(), C.c_uint64()
    check(L.sb200_hyperball_state_bytes(hs[0]._h, C.byref(rb), C.byref(bb)))
    bufs = []
    for r in range(world):
        b = [aligned(rb.value, np.uint8), aligned(rb.value, np.uint8), aligned(bb.value, np.uint32), aligned(bb.value, np.uint32)]
        bufs.append(b)
        check(L.sb200_hyperball_bind_state(hs[r]._h, *(x[0].ctypes.data for x in b)))
    for r in range(world):
        peers = [p for p in range(world) if p != r]
        cols = [(C.c_uint64 * len(peers))(*(bufs[p][i][0].ctypes.data for p in peers)) for i in range(4)]
        check(L.sb200_hyperball_set_publish_targets(hs[r]._h, len(peers), *cols))
    ref = DenseHyperBall(d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    t = 0
    while True:
        sts = [h.step() for h in hs]
        total = sum(s["n_changed"] for s in sts)
        for h in hs: h.exchange_done(total)
        ch = ref.step(); t += 1
        regs = [h.registers() for h in hs]
        assert all(np.array_equal(regs[0], x) for x in regs[1:]), ("replicas differ", t)
        assert np.array_equal(regs[0], ref.registers()), ("registers differ from the oracle", t)
        assert (total == 0) == (not ch), (t, total, ch)
        if total == 0: break
    lo = np.concatenate([h.result()[0] for h in hs]); hi = np.concatenate([h.result()[1] for h in hs]); c = np.concatenate([h.result()[2] for h in hs])
    f = hyperball_faithful(d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    key = hi.astype(object) * (1 << 64) + lo.astype(object); o = np.argsort(key)
    assert np.array_equal(lo[o], f["ids_lo"]) and np.array_equal(c[o], f["centrality"]) and t == f["iters"]
    for h in hs: h.close()
    print("world", world, "fused exchange by address: ok,", t, "iterations")
print("sharded emulated parity ok")
