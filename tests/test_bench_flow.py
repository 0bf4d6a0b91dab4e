"""Control flow of the multi-GPU paths of bench.py without GPUs (tests/dryrun_bench.py): every rank is a gloo process, the C ABI
runs on the CPU SIMT emulator, CUDA plumbing is faked.  Checks that each exchange selection prints ONE JSON line on rank 0 with
a green parity object (kernel-loop shares AND the end-to-end leg against the fingerprint) and an `e2e` object."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


def _launch(world, extra, symm="ok", tmp=None):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), DRY_SYMM=symm,
               SB200_BENCH_FINGERPRINT=tmp, OMP_NUM_THREADS="1")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dryrun_bench.py"), "--gpus", str(world), "--steps", "1", "--warmup", "0",
                                       "--e2e-steps", "2"] + extra, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), [(p.returncode, o[1][-1500:]) for p, o in zip(procs, outs)]
    lines = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    assert len(lines) == 1, outs[0][0][-2000:]
    assert all(not [l for l in o[0].splitlines() if l.startswith("{")] for o in outs[1:])     # only rank 0 prints
    return json.loads(lines[0])


@pytest.fixture(scope="module")
def fingerprint_file(tmp_path_factory):
    subprocess.check_call(["make", "-C", EMU], stdout=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, os.path.join(HERE, "dryrun_bench.py"), "--fingerprint"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    p = tmp_path_factory.mktemp("fp") / "fingerprint.json"
    p.write_text([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    return str(p)


def _check(line, world, kind_contains, e2e_expected=True):
    assert line["n_gpus"] == world and line["scaling"] == "strong" and line["value"] > 0
    assert kind_contains in line["run"]["parallelism"], line["run"]["parallelism"]
    c2 = line["parity"]["c2"]
    assert c2["green"] is True, c2
    if e2e_expected:
        assert c2["equal"]["e2e_result_checksum"] is True
        e = line["e2e"]
        assert e["value"] > 0 and e["steps"] == 2 and len(e["rank0_phase_ms"]) == 2 and "gather_shards" in e["rank0_phase_ms"][0], e
    assert len(line["per_iter"]) == world and line["kernels"]


def test_two_ranks_default_is_peer_stores_with_e2e_after(fingerprint_file):
    _check(_launch(2, [], tmp=fingerprint_file), 2, "CUDA IPC")


def test_explicit_multicast_runs_the_e2e_leg_first(fingerprint_file):
    _check(_launch(2, ["--exchange", "multicast"], tmp=fingerprint_file), 2, "symmetric-memory multicast")


def test_eight_ranks_auto_binds_multicast(fingerprint_file):
    _check(_launch(8, [], tmp=fingerprint_file), 8, "symmetric-memory multicast")


@pytest.mark.parametrize("symm", ["raise", "unicast"])
def test_eight_ranks_auto_falls_back_to_peer_stores(fingerprint_file, symm):
    _check(_launch(8, [], symm=symm, tmp=fingerprint_file), 8, "CUDA IPC")


def test_no_e2e_flag(fingerprint_file):
    line = _launch(2, ["--no-e2e"], tmp=fingerprint_file)
    _check(line, 2, "CUDA IPC", e2e_expected=False)
    assert line["e2e"] is None and "e2e_result_checksum" not in line["parity"]["c2"]["equal"]


def test_single_gpu_flow_with_full_oracle_parity():
    """N = 1: kernel loop, the host-buffer e2e leg and the full-graph oracle comparison (`parity.c2`, `cpu_baseline`) -- without
    the BM25 and C1 legs, which need their own GPU-sized inputs."""
    subprocess.check_call(["make", "-C", EMU], stdout=subprocess.DEVNULL)
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0", OMP_NUM_THREADS="2")
    env.pop("SB200_BENCH_FINGERPRINT", None)
    out = subprocess.run([sys.executable, os.path.join(HERE, "dryrun_bench.py"), "--gpus", "1", "--steps", "1", "--warmup", "0", "--no-bm25", "--no-c1",
                          "--e2e-steps", "2"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    c2 = d["parity"]["c2"]
    assert c2["green"] is True and all(c2["equal"].values()) and c2["equal"]["e2e_result"] is True
    assert d["n_gpus"] == 1 and d["e2e"]["steps"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["roofline"]["kernel"].startswith("k_pull")
    for k in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "clocks", "gpu_launches"):
        assert k in d
