#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for stract_b200.

Headline metric (BASELINE.json): webgraph edges/sec per centrality iteration, on configs[1]
(50M-node / 1B-edge R-MAT host graph, harmonic centrality to convergence on 1xB200; with --gpus N the
same graph is destination-row partitioned over N GPUs = configs[2]).  A "step" is one complete
HarmonicCentrality computation (reset + all HyperBall iterations to convergence) on the graph
resident in HBM; value = kept_edges x iterations / device time.  `e2e` is the same metric through the
C-ABI call sequence a Rust shim makes (sb200_graph_create from HOST buffers -> run -> result to host),
host<->device copies and the on-device CSR staging inside the timed region.  BM25 postings/sec is
reported in the same line under "bm25".

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                          the reference's CPU path (oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "webgraph_edges_per_sec_per_centrality_iter"
UNIT = "edges/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx = float(c[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def _unique_kept_edges(d):
    import numpy as np
    from stract_b200.webgraph import SKIPPED_REL
    key = np.stack([d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"]], 1)
    _, first = np.unique(key, axis=0, return_index=True)
    keep = (d["rel_flags"][first] & np.uint64(SKIPPED_REL)) == 0
    return int(keep.sum())


# ------------------------------------------------------------------------------------------------
def cpu_baseline_dense(nodes, edges, threads):
    """Oracle 'dense' port (flat arrays, all host threads) on a bounded R-MAT sample of the workload."""
    import oracle
    from stract_b200 import synth
    d = synth.rmat_graph(nodes, edges, seed=42)
    a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    o = oracle.DenseHyperBall(*a, threads=threads)
    t0 = time.perf_counter()
    iters = o.run()
    dt = time.perf_counter() - t0
    kept = o.n_edges
    o.close()
    return {"value": kept * iters / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"R-MAT {nodes} nodes / {edges} edges (same generator, seed 42), {iters} iterations to convergence, "
                      f"oracle dense port on {threads} threads, iteration loop only (graph already staged in RAM)"}


def run_reference(args):
    """--impl reference: the reference's own CPU path = oracle 'faithful' port (the reference is Rust; no
    toolchain here).  It mirrors harmonic.rs structure by structure and is single-threaded because the
    reference's loop is (harmonic.rs:129-154).  Each step = one HarmonicCentrality::calculate on a bounded
    sample, graph scan/dedup/maps included, exactly what the reference call does."""
    import oracle
    from stract_b200 import synth
    nodes, edges = args.ref_nodes, args.ref_edges
    d = synth.rmat_graph(nodes, edges, seed=42)
    a = (d["from_lo"], d["from_hi"], d["to_lo"], d["to_hi"], d["rel_flags"])
    kept = _unique_kept_edges(d)
    for _ in range(min(args.warmup, 1)):
        oracle.hyperball_faithful(*a)
    t0 = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        r = oracle.hyperball_faithful(*a)
        iters = r["iters"]
    dt = (time.perf_counter() - t0) / args.steps
    value = kept * iters / dt
    sample = (f"R-MAT {nodes} nodes / {edges} edges (bounded sample of the 50M/1B workload, same generator), "
              f"{iters} iterations, one full calculate() per step")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "webgraph harmonic centrality (HyperBall), R-MAT 50M nodes / 1B edges -- bounded CPU sample",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def gen_device_graph(torch, L, dev_index, nodes, edges, scale):
    from stract_b200._lib import check
    t = [torch.empty(edges, dtype=torch.int64, device=f"cuda:{dev_index}") for _ in range(5)]
    CH = 1 << 27
    for first in range(0, edges, CH):
        cnt = min(CH, edges - first)
        check(L.sb200_synth_edges(1, nodes, first, cnt, 42, scale, dev_index, *(x.data_ptr() + first * 8 for x in t)))
    return t


def main():
    t_bench0 = time.perf_counter()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=50_000_000)
    ap.add_argument("--edges", type=int, default=1_000_000_000)
    ap.add_argument("--scale", type=int, default=26)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bm25", action="store_true")
    ap.add_argument("--no-experimental", action="store_true", help="skip the subprocess that times the opt-in BM25 kernels")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL byte-max all-reduce exchange instead of the fused peer-memory stores")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "symm", "multicast"],
                    help="multi-GPU fused exchange transport: CUDA IPC peer mappings (default), torch symmetric memory "
                         "unicast, or NVSwitch multicast stores")
    ap.add_argument("--cpu-nodes", type=int, default=1_000_000)
    ap.add_argument("--cpu-edges", type=int, default=20_000_000)
    ap.add_argument("--ref-nodes", type=int, default=50_000)
    ap.add_argument("--ref-edges", type=int, default=500_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from stract_b200 import kernel_launch_count, lib
    from stract_b200.webgraph import DeviceGraph, ShardedHarmonicCentrality, Webgraph

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    peaks, peak_src = _peaks()

    nodes, edges = args.nodes, args.edges
    tg = time.perf_counter()
    cols = gen_device_graph(torch, L, local_rank, nodes, edges, args.scale)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - tg
    graph = Webgraph.from_arrays(*cols)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    launches0 = kernel_launch_count()
    result = {}
    if world == 1:
        dg = DeviceGraph(graph, device=local_rank)
        info = dg.info()
        E = info["n_edges_kept"]

        def one_step():
            dg.reset()
            iters, stats = dg.run()
            return iters, stats
        for _ in range(max(args.warmup, 3)):
            one_step()
        dg.set_profiling(True)
        sampler = ClockSampler(local_rank); sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches_t0 = kernel_launch_count()
        ev0.record()
        tot_iters = 0
        stats_last = None
        for _ in range(args.steps):
            iters, stats_last = one_step()
            tot_iters += iters
        ev1.record()
        barrier()
        clocks = sampler.stop()
        launches_timed = kernel_launch_count() - launches_t0
        ms_total = ev0.elapsed_time(ev1)
        prof = dg.profile()
        dg.set_profiling(False)
        value = E * tot_iters / (ms_total * 1e-3)
        iters = tot_iters // args.steps
        dom = max((p for p in prof if "dense" in p["name"] and p["launches"]), key=lambda p: p["ms"], default=None)
        roofline = None
        if dom:
            ach = dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9
            # DRAM bytes per launch come from the committed `ncu --set full` capture of this kernel and only
            # apply when the workload is the one that was captured; otherwise null
            traffic, traffic_src = None, None
            try:
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as fh:
                    cap = json.load(fh)
                if cap["config"]["nodes"] == args.nodes and cap["config"]["edges"] == args.edges and "pull_warp" in dom["name"]:
                    traffic = cap["k_pull_warp<dense>"]["dram_bytes_per_launch"]
                    traffic_src = cap["source"]
            except (OSError, KeyError, ValueError):
                pass
            roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src,
                        "launches": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
                        "alg_bytes_per_launch": dom["alg_bytes"] / dom["launches"]}
        result.update(value=value, ms_per_step=ms_total / args.steps, iters=iters, E=E, info=info, clocks=clocks,
                      roofline=roofline, launches=launches_timed,
                      kernels=[{**p, "share_of_step": p["ms"] / ms_total} for p in prof if p["launches"]],
                      per_iter=[{"t": s["t"], "mode": s["mode"], "n_changed": s["n_changed"], "ms": round(s["ms"], 3)} for s in stats_last])
        dg.close()
    else:
        # configs[2]: the same graph, destination rows partitioned over `world` GPUs
        def one_run():
            return ShardedHarmonicCentrality.calculate(graph, local_rank, rank, world)
        # staging happens inside calculate(); time only the iteration loops via the per-iteration stats
        # (device ms of the step kernels) + exchange wall time => use wall clock around the loop
        from stract_b200.webgraph import DeviceGraph as _DG
        dg = _DG(graph, device=local_rank, rank=rank, world_size=world)
        exchange_kind = "nccl"
        if args.exchange in ("symm", "multicast"):
            exchange_kind = "symmetric-memory " + dg.enable_symmetric(multicast=(args.exchange == "multicast"))
        elif not args.no_p2p:
            dg.enable_p2p()
            exchange_kind = "p2p"
        info = dg.info()
        E = info["n_edges_kept"]
        ranges = dg.row_ranges()
        from stract_b200.webgraph import run_sharded_loop

        def one_step():
            dg.reset()
            t, _ = run_sharded_loop(dg, world)
            return t
        for _ in range(max(args.warmup, 3)):
            one_step()
        dg.set_profiling(True)
        sampler = ClockSampler(local_rank); sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches_t0 = kernel_launch_count()
        ev0.record()
        tot_iters = 0
        for _ in range(args.steps):
            tot_iters += one_step()
        ev1.record()
        barrier()
        clocks = sampler.stop()
        roofline_n = None
        try:   # rank 0's dominant kernel over its owned share of the rows (algorithmic bytes x owned fraction)
            prof = dg.profile()
            dom = max((p for p in prof if "dense" in p["name"] and p["launches"]), key=lambda p: p["ms"], default=None)
            if dom:
                ach = dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9
                roofline_n = {"bound": "hbm", "kernel": dom["name"] + " (rank 0, owned rows)", "achieved": ach, "peak": peaks["hbm_gbs"],
                              "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                              "launches": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
                              "alg_bytes_per_launch": dom["alg_bytes"] / dom["launches"],
                              "note": "includes the stores of the produced rows into the peers' replicas over NVLink"}
        except Exception:  # noqa: BLE001
            roofline_n = None
        dg.set_profiling(False)
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        nl = torch.tensor([kernel_launch_count() - launches_t0], device=dev, dtype=torch.int64)
        dist.all_reduce(nl)
        value = E * tot_iters / (ms_total * 1e-3)
        # one extra untimed run for the per-rank, per-iteration device times (load balance evidence)
        dg.reset()
        _, st_last = run_sharded_loop(dg, world)
        mine = {"rank": rank, "rows": [info["row_begin"], info["row_end"]], "edges_local": info["n_edges_local"],
                "iter_ms": [round(s["ms"], 3) for s in st_last], "modes": [s["mode"] for s in st_last]}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        result.update(exchange_kind=exchange_kind)
        result.update(value=value, ms_per_step=ms_total / args.steps, iters=tot_iters // args.steps, E=E, info=info,
                      clocks=clocks, roofline=roofline_n, launches=int(nl.item()), kernels=[], per_iter=allr)
        dg.close()

    # ---- e2e: the C-ABI call sequence from HOST buffers (rank 0 only drives it at N=1) ----------
    e2e = None
    if not args.no_e2e and world == 1:
        import psutil
        need = edges * 40
        avail = psutil.virtual_memory().available
        e_nodes, e_edges, e_scale, note = nodes, edges, args.scale, "full workload"
        if avail < need * 1.6:
            f = 1
            while (edges // f) * 40 * 1.6 > avail and f < 1024:
                f *= 2
            e_nodes, e_edges = max(nodes // f, 1000), edges // f
            e_scale = max(1, int(np.ceil(np.log2(e_nodes))))
            note = f"host RAM {avail / 2**30:.0f} GiB < 1.6x the {need / 2**30:.0f} GiB edge stream: scaled 1/{f}"
            del cols, graph
            torch.cuda.empty_cache()
            cols = gen_device_graph(torch, L, local_rank, e_nodes, e_edges, e_scale)
        host = []
        for c in cols:
            try:
                h = torch.empty(c.shape, dtype=c.dtype, pin_memory=True)
            except Exception:
                h = torch.empty(c.shape, dtype=c.dtype)
            h.copy_(c)
            host.append(h)
        pinned = all(h.is_pinned() for h in host)
        del cols
        torch.cuda.empty_cache()
        hgraph = Webgraph.from_arrays(*host)
        from stract_b200.webgraph import HarmonicCentrality

        def e2e_step():
            r = HarmonicCentrality.calculate(hgraph, device=local_rank)
            return r
        for _ in range(2):  # warm-up: device memory pool and the page-locked result blocks reach steady state
            r = e2e_step()
            del r
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        tot, d2h, checksum, walls = 0, 0, 0.0, []
        for _ in range(args.e2e_steps):
            r = e2e_step()
            tot += r.info["n_edges_kept"] * r.iterations
            d2h = len(r.values) * 24
            checksum += float(r.values[:1024].sum())  # the caller reads the result on the host, then drops it
            last = {"stage_ms": r.info["stage_ms"], "iterations": r.iterations}
            walls.append({k: round(v, 1) for k, v in (r.info.get("wall_ms") or {}).items()})
            del r
        ev1.record()
        barrier()
        ms_e = ev0.elapsed_time(ev1)
        e2e = {"value": tot / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e_edges * 40, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e / args.e2e_steps, "steps": args.e2e_steps, "pinned_host": pinned, "workload": note,
               "nodes": e_nodes, "edges": e_edges, "stage_ms": last["stage_ms"], "iterations": last["iterations"],
               "step_wall_ms": walls}

    if rank == 0:
        line = {"metric": METRIC, "value": result["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": result["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"webgraph harmonic centrality (HyperBall) to convergence, R-MAT(0.57,0.19,0.19,0.05) "
                                       f"{nodes} nodes / {edges} edges (BASELINE configs[{1 if world == 1 else 2}])",
                           "kept_edges": result["E"], "n_nodes": result["info"]["n_nodes"],
                           "iterations_per_step": result["iters"],
                           "l2_policy": "inputs >> L2: 2 x 3.2 GB register arrays + 4 GB CSR per iteration",
                           "parallelism": "1 GPU" if world == 1 else (f"destination-row partition x{world}, " + {"nccl": "NCCL byte-max all-reduce of the register replicas per iteration",
                                                                                                    "p2p": "fused exchange: pull kernels store produced rows into all peers' replicas over NVLink (CUDA IPC), NCCL all-reduce of the changed count as barrier"}.get(
                               result.get("exchange_kind", "p2p"), "fused exchange over " + str(result.get("exchange_kind")) + " stores, NCCL all-reduce of the changed count as barrier")),
                           "hbm_bytes": result["info"]["hbm_bytes"], "gen_s": round(gen_s, 2),
                           "stage_ms": result["info"]["stage_ms"]},
                "clocks": result["clocks"], "gpu_launches": result["launches"], "roofline": result["roofline"],
                "kernels": result["kernels"], "per_iter": result["per_iter"], "e2e": e2e}
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline_dense(args.cpu_nodes, args.cpu_edges, threads)
        if not args.no_bm25 and world == 1:
            try:
                from stract_b200 import bm25_bench
                line["bm25"] = bm25_bench.run(local_rank, peaks, peak_src)
            except ImportError:
                line["bm25"] = None
            if line.get("bm25") is not None and not args.no_experimental:
                # the opt-in BM25 kernels have passed their parity tests on the CPU emulator only: measured in a separate
                # process (own CUDA context, bounded time), reported beside -- never instead of -- the default kernels
                try:
                    r = subprocess.run([sys.executable, "-m", "stract_b200.bm25_bench", str(local_rank)], cwd=ROOT,
                                       capture_output=True, text=True, timeout=180)
                    line["bm25"]["experimental"] = (json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0
                                                    else {"error": (r.stderr or r.stdout)[-400:]})
                except Exception as ex:  # noqa: BLE001
                    line["bm25"]["experimental"] = {"error": repr(ex)[:400]}
        if world == 1 and not args.no_experimental and args.nodes == 50_000_000 and args.edges == 1_000_000_000:
            # path-1 switches that have only run on the CPU emulator, each in its own process and only
            # while the whole bench is still short; reported beside the main numbers, never instead of them
            exp = {}
            try:   # hand the staging pool of this process back first: the child needs the HBM
                from stract_b200._lib import lib as _sblib
                torch.cuda.empty_cache()
                _sblib().sb200_release_cached_memory(local_rank)
            except Exception:  # noqa: BLE001
                pass
            variants = [("l2_persist_64MB", {"SB200_L2_PERSIST_MB": "64"}, ["--no-e2e", "--steps", "3"]),
                        ("l2_evict_first_hints", {"SB200_L2_HINTS": "1"}, ["--no-e2e", "--steps", "3"]),
                        ("e2e_arena_rowperm", {"SB200_ARENA": "1", "SB200_STAGE_ROWPERM": "1"}, ["--steps", "1", "--e2e-steps", "3"])]
            for name, env, extra in variants:
                if time.perf_counter() - t_bench0 > 300:
                    exp[name] = {"skipped": "bench time budget"}
                    continue
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--no-bm25", "--no-cpu", "--no-experimental", *extra],
                                       cwd=ROOT, env={**os.environ, **env}, capture_output=True, text=True, timeout=150)
                    if r.returncode != 0:
                        exp[name] = {"error": (r.stderr or r.stdout)[-300:]}
                        continue
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                    exp[name] = {"env": env, "ms_per_step": d["ms_per_step"], "iterations": d["config"]["iterations_per_step"],
                                 "top_kernel_avg_ms": (d.get("roofline") or {}).get("avg_launch_ms"),
                                 "e2e_ms_per_step": (d.get("e2e") or {}).get("ms_per_step"),
                                 "e2e_step_wall_ms": (d.get("e2e") or {}).get("step_wall_ms")}
                except Exception as ex:  # noqa: BLE001
                    exp[name] = {"error": repr(ex)[:300]}
            line["experimental"] = exp
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
