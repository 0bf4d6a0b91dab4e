"""Worker of tests/test_bench_flow.py: runs bench.main() for one rank of an N > 1 launch WITHOUT GPUs, to exercise the control
flow of the multi-GPU bench paths (exchange selection and fallback, the end-to-end leg before / after the kernel loop, parity
merge, the JSON line).  Test infrastructure only.

What is real: bench.py itself, stract_b200.webgraph (gather_edge_shards, ShardedHarmonicCentrality, run_sharded_loop), the C ABI
(on the CPU SIMT emulator, tests/emu) and torch.distributed (gloo).  What is faked: CUDA device objects / events / pinned memory,
the synthetic graph generator (a small graph), nvidia-smi sampling, and the exchange between ranks -- every rank computes the
whole small graph on a single-rank emulated handle and reports an interleaved share of it as "owned"."""
import ctypes as C
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def install_fakes(symm_mode):
    import numpy as np
    import torch as real_torch
    import torch.distributed as real_dist
    from stract_b200 import _lib, synth
    import stract_b200.webgraph as wg

    L = _lib.declare(C.CDLL(os.path.join(HERE, "emu", "libsb200_emu.so")))
    assert b"emulation" in L.sb200_version()
    _lib._LIB = L

    # ---- torch proxy: CPU devices, fake events, no pinned memory, gloo instead of nccl
    class FakeEvent:
        def __init__(self, enable_timing=False): self.t = None
        def record(self, stream=None): self.t = time.perf_counter()
        def elapsed_time(self, other): return (other.t - self.t) * 1e3

    cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda *a: None,
                                 empty_cache=lambda: None, Event=FakeEvent)

    class DistProxy(types.ModuleType):
        def __getattr__(self, k): return getattr(real_dist, k)
    dist = DistProxy("torch.distributed")
    dist.init_process_group = lambda backend=None, device_id=None, **kw: real_dist.init_process_group("gloo", **kw)

    class TorchProxy(types.ModuleType):
        def __getattr__(self, k): return getattr(real_torch, k)
    tp = TorchProxy("torch")
    tp.cuda = cuda
    tp.distributed = dist
    tp.device = lambda *a, **k: real_torch.device("cpu")

    def empty(*a, **k):
        k.pop("pin_memory", None)
        return real_torch.empty(*a, **k)
    tp.empty = empty
    sys.modules["torch"] = tp
    sys.modules["torch.distributed"] = dist

    # ---- the handle: one emulated single-rank graph per process, an interleaved share reported as owned
    Real = wg.DeviceGraph

    class DryDeviceGraph(Real):
        def __init__(self, graph, device=0, rank=0, world_size=1, skipped_rel=wg.SKIPPED_REL):
            g2 = wg.Webgraph.from_arrays(*[np.ascontiguousarray(np.asarray(a).view(np.uint64)) for a in
                                           (graph.from_lo, graph.from_hi, graph.to_lo, graph.to_hi, graph.rel)])
            super().__init__(g2, device=0, rank=0, world_size=1, skipped_rel=skipped_rel)
            self.world_size, self.rank, self.device = world_size, rank, device

        def info(self):
            i = super().info()
            i["n_edges_local"] = i["n_edges_kept"] // self.world_size
            return i

        def enable_p2p(self, group=None):
            real_dist.barrier(group=group)
            self.p2p = True

        def enable_symmetric(self, group=None, multicast=True):
            if symm_mode == "raise":
                raise RuntimeError("dry run: no symmetric memory")
            self.p2p = True
            return "unicast" if symm_mode == "unicast" else ("multicast" if multicast else "unicast")

        def run_sharded(self, max_iters=0, cap=256):
            return self.run(max_iters, cap)

        def exchange_done(self, total):
            pass

        def ownership(self):
            n = super().info()["n_nodes"]
            owned = (((np.arange(n) >> 5) % self.world_size) == self.rank).astype(np.uint8)
            return owned, np.zeros(n, np.uint32)

        def result(self):
            lo, hi, c = super().result()
            m = (np.arange(len(c)) % self.world_size) == self.rank
            return lo[m], hi[m], c[m]
    wg.DeviceGraph = DryDeviceGraph

    import bench
    d = synth.rmat_graph(2500, 30000, seed=13)
    cols = [real_torch.from_numpy(np.ascontiguousarray(d[k]).view(np.int64).copy()) for k in ("from_lo", "from_hi", "to_lo", "to_hi", "rel_flags")]
    bench.gen_device_graph = lambda torch, L_, dev_index, nodes, edges, scale, kind=1: [c.clone() for c in cols]

    class FakeSampler:
        def __init__(self, gpu_index=0): pass
        def start(self): pass
        def stop(self): return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    bench.ClockSampler = FakeSampler
    return bench, len(cols[0])


def fingerprint():
    """The fingerprint file of the dry-run graph (N = 1 on the emulator), in the format of tests/golden/path1_c2.json."""
    bench, n_edges = install_fakes("ok")
    import numpy as np
    import stract_b200.webgraph as wg
    import torch
    cols = bench.gen_device_graph(torch, None, 0, 0, 0, 0)
    dg = wg.DeviceGraph(wg.Webgraph.from_arrays(*cols))
    t, _ = dg.run()
    lo, hi, c = dg.result()
    regs = dg.registers()
    out = {"registers_checksum": bench.registers_checksum(regs), "result_checksum": bench.result_checksum(lo, hi, c), "n_positive": int(len(c)),
           "n_nodes": int(dg.info()["n_nodes"]), "iterations": int(t)}
    dg.close()
    return out, n_edges


if __name__ == "__main__":
    symm_mode = os.environ.get("DRY_SYMM", "ok")
    if sys.argv[1] == "--fingerprint":
        import json
        fp, n_edges = fingerprint()
        fp["n_edges"] = n_edges
        print(json.dumps(fp))
        sys.exit(0)
    bench, n_edges = install_fakes(symm_mode)
    sys.argv = ["bench.py"] + sys.argv[1:] + ["--edges", str(n_edges), "--nodes", "2500"]
    sys.exit(bench.main())
