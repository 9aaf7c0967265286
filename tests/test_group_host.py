"""CPU tests of the multi-GPU plumbing behind the C ABI (volcanor_b200/csrc/group.hpp, compiled for the host by
tests/native/group_host.cpp): the target partition, the persistent worker pool, the two-barrier slot exchange that
vlc_wake_sweep's peer path uses, error reporting in member order, and the run-time binding of NCCL.  The device side of
the same code runs in tests/test_gpu_group.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent / "native"


@pytest.fixture(scope="module")
def glib():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    lib = C.CDLL(str(HERE / "libgroup_host.so"))
    lib.grp_shard.argtypes = [C.c_longlong, C.c_int, C.c_void_p]
    lib.grp_emulate_wake_sweeps.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.grp_worker_threads.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.grp_nccl_available.argtypes = [C.POINTER(C.c_int)]
    return lib


@pytest.mark.parametrize("M", [0, 1, 7, 8, 9, 555, 4320, 30200, 258176])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_every_target_once_with_equal_slots(glib, M, world):
    """Slices are contiguous, disjoint, cover [0, M), and all have the same slot size per with per*world >= M (what an
    in-place all-gather of equal counts needs); the same arithmetic as volcanor_b200/sharding.py (round 1, gloo tests)."""
    out = np.zeros(3 * world, dtype=np.int64)
    glib.grp_shard(M, world, out.ctypes.data)
    per, lo, hi = out[0::3], out[1::3], out[2::3]
    assert np.all(per == per[0]) and per[0] * world >= M and (per[0] - 1) * world < max(M, 1)
    assert lo[0] == 0 and hi[-1] == M and np.all(lo[1:] == hi[:-1]) and np.all(hi - lo <= per[0]) and np.all(hi >= lo)
    assert np.all(lo == np.minimum(np.arange(world) * per[0], M))
    from volcanor_b200.sharding import TargetShard
    for r in range(world):
        t = TargetShard(M, world, r)
        assert (t.per, t.lo, t.hi) == (max(per[0], 0) if M else t.per, lo[r], hi[r]) or M == 0


@pytest.mark.parametrize("n,M", [(1, 10), (2, 4320), (3, 555), (8, 30200), (8, 5)])
def test_slot_exchange_gives_every_member_the_whole_list(glib, n, M):
    err = C.c_int()
    assert glib.grp_emulate_wake_sweeps(n, M, 25, -1, 0, C.byref(err)) == 0
    assert err.value == 0


def test_first_failing_member_is_reported(glib):
    err = C.c_int()
    assert glib.grp_emulate_wake_sweeps(4, 1000, 3, 2, 7, C.byref(err)) == 0
    assert err.value == 7                      # member 2's status comes back from run(); later rounds are clean


def test_workers_are_persistent_one_thread_per_member(glib):
    on_caller = C.c_int()
    assert glib.grp_worker_threads(5, 200, C.byref(on_caller)) == 4
    assert on_caller.value == 1                # the leader's share runs on the caller's thread


def test_nccl_binds_at_run_time():
    """libnccl.so.2 is present in this image (system 2.27 / torch's 2.28): the dlopen binding finds every entry point it
    needs.  No communicator is made here (no GPU).  In a subprocess: binding the SYSTEM libnccl into the pytest process
    would make a later `import torch` pick it up by soname instead of its own newer copy (volcanor_b200/api.py preloads
    torch's copy for the same reason)."""
    import sys
    code = ("import ctypes as C; lib = C.CDLL(r'%s'); v = C.c_int(); ok = lib.grp_nccl_available(C.byref(v)); "
            "print(ok, v.value)" % (HERE / "libgroup_host.so"))
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == 1 and int(out[1]) >= 20000, out
