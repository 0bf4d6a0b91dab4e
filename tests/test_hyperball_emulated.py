"""Path-1 (HyperBall) kernels and staging on the CPU SIMT emulator (tests/emu; see tests/test_bm25_emulated.py for what
the emulator is and is not).  The parity functions are those of tests/test_hyperball_gpu.py: registers bit-exact after
every iteration, KahanSum bit-exact, output ids and values bit-exact against the oracle, all three kernel families
forced, rows that span several work items.  Besides checking the kernels under the largest lane skew a GPU may show,
this validates two staging switches that were written without a GPU at hand: the row-permutation CSR relabel and the
slab arena."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


def _run(mode, **env):
    subprocess.check_call(["make", "-C", EMU], stdout=subprocess.DEVNULL)
    e = dict(os.environ)
    for k in ("SB200_ARENA", "SB200_STAGE_ROWPERM", "SB200_ARENA_SLAB_MB"):
        e.pop(k, None)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(EMU, "run_path1.py"), mode], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "path-1 emulated parity ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


def test_default_path_all_kernel_families():
    _run("full")


def test_row_permutation_relabel():
    _run("quick", SB200_STAGE_ROWPERM="1")


def test_slab_arena():
    out = _run("quick", SB200_ARENA="1", SB200_ARENA_SLAB_MB="4")
    assert "arena: reserved" in out


def test_fused_exchange_by_address_two_and_three_ranks():
    """sb200_hyperball_bind_state + sb200_hyperball_set_publish_targets with several ranks in one emulated process."""
    subprocess.check_call(["make", "-C", EMU], stdout=subprocess.DEVNULL)
    r = subprocess.run([sys.executable, os.path.join(EMU, "run_sharded.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "sharded emulated parity ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
