"""world_size-2 `gloo` test of the multi-GPU host logic (CPU): target sharding + the one all-gather per
convection stage give the same node positions as a single rank.  The per-slice velocities come from the CPU oracle
here (test infrastructure); on the GPUs the same code path runs with vlc_vind_dev and NCCL (bench.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")
    import torch
    import torch.distributed as dist
    from oracle import pyoracle
    from volcanor_b200 import synth
    from volcanor_b200.sharding import TargetShard, allgather_slices

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lats = synth.multirotor(1500, seed=3, n_rotor=1, nb=2, S=4, F=3)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    P = synth.targets_all(lats)
    m = P.shape[0]
    sh = TargetShard(m, world, rank)
    P_all = torch.zeros(sh.padded, 3, dtype=torch.float64)
    P_all[:m] = torch.from_numpy(P)
    dt = 1e-3
    for stage in range(2):                                # predictor + corrector stage of one step
        Pn = P_all[:m].numpy()
        if sh.count > 0:
            V = pyoracle.vind_flat(p1, p2, rvc, gam, flag, Pn[sh.lo:sh.hi])
            P_all[sh.lo:sh.hi] += torch.from_numpy(V) * dt   # convect my slice
        allgather_slices(P_all, sh)
    np.save(Path(out_dir) / f"P_rank{rank}.npy", P_all[:m].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_stage_equals_single_rank(tmp_path, oracle):
    import torch.multiprocessing as mp
    from volcanor_b200 import synth
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    a = np.load(tmp_path / "P_rank0.npy")
    b = np.load(tmp_path / "P_rank1.npy")
    assert np.array_equal(a, b)                              # every rank ends with the same node set
    # single-rank reference of the same two stages
    lats = synth.multirotor(1500, seed=3, n_rotor=1, nb=2, S=4, F=3)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    P = synth.targets_all(lats).copy()
    for stage in range(2):
        P += oracle.vind_flat(p1, p2, rvc, gam, flag, P) * 1e-3
    assert np.array_equal(a, P)


@pytest.mark.parametrize("m,world", [(0, 2), (1, 2), (7, 2), (8, 2), (9, 4), (258176, 8), (3, 8)])
def test_shard_arithmetic(m, world):
    from volcanor_b200.sharding import TargetShard
    covered = []
    for r in range(world):
        s = TargetShard(m, world, r)
        assert 0 <= s.lo <= s.hi <= m and s.count <= s.per and s.padded >= m
        covered += list(range(s.lo, s.hi))
    assert covered == list(range(m))                         # contiguous, disjoint, complete


# ---- gridgen: the cell list sharded across ranks (volcanor_b200/gridgen.py:sharded_velocities) ----

class _OracleGridCtx:
    """Stand-in for Context.gridgen / .gridgen_slice backed by the oracle's restatement of program gridgen (test
    infrastructure): the host logic under test is the slicing and the gather, not the velocities."""

    def __init__(self, oracle):
        self.o = oracle

    def gridgen(self, *a):
        return self.o.gridgen(*a)

    def gridgen_slice(self, *a):
        *args, first, count = a
        gc, vc = self.o.gridgen(*args)
        return gc.reshape(-1, 3)[first:first + count].copy(), vc.reshape(-1, 3)[first:first + count].copy()


def _grid_case():
    rng = np.random.default_rng(5)
    f = {"vrWing": np.zeros((0, 50)), "vrNwake": np.zeros((0, 50)), "vfNwakeTE": np.zeros((0, 12)), "gamNwakeTE": np.zeros(0),
         "vfFwake": np.concatenate([rng.uniform(-1, 1, (40, 6)), np.zeros((40, 3)), np.full((40, 1), 0.05), np.zeros((40, 2))], axis=1),
         "gamFwake": rng.uniform(-1, 1, 40)}
    cfg = {"nx": 6, "ny": 5, "nz": 4, "xyzMin": [-2.0, -1.5, -1.0], "xyzMax": [2.0, 1.5, 0.5], "vel": [3.0, 0.0, -1.0]}
    return cfg, f


def _grid_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")
    import torch.distributed as dist
    from oracle import pyoracle
    from volcanor_b200 import gridgen
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, f = _grid_case()
    vc = gridgen.sharded_velocities(_OracleGridCtx(pyoracle), cfg, f, world, rank)
    np.save(Path(out_dir) / f"vc_rank{rank}.npy", vc)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gridgen_cells_sharded_across_ranks(tmp_path, oracle, world):
    """60 cells over 2 and 3 ranks (3: the last slice is shorter than the others): every rank ends with the whole field,
    equal to the single-process call."""
    import torch.multiprocessing as mp
    from volcanor_b200 import gridgen
    mp.start_processes(_grid_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    cfg, f = _grid_case()
    ref = gridgen.sharded_velocities(_OracleGridCtx(oracle), cfg, f, 1, 0)
    assert ref.shape == (3, 4, 5, 3) and np.any(ref != np.array(cfg["vel"]))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"vc_rank{r}.npy"), ref), r


# ---- the reference cases with one process per GPU: the staged orchestration's sharded wake sweep on the CPU backend ----

def _case_worker(rank, world, port, out_dir, name, nsteps):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")
    import ctypes as C
    import json
    import torch
    import torch.distributed as dist
    from oracle import pyoracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    if name == "caradonna":
        fx["config"]["nt"] = 40
        fx["geom"][0].update(nNwake=12, wakeTruncateNt=18)
    lib = C.CDLL(str(ROOT / "tests" / "native" / "libcase_gpu_hooks.so"))
    lib.case_cpu_staged_hooks_install.restype = C.c_void_p
    lib.case_cpu_staged_hooks_install.argtypes = [C.c_void_p, C.c_int]
    lib.case_gpu_hooks_set_sharding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    lib.case_gpu_hooks_exchanges.restype = C.c_long
    lib.case_gpu_hooks_exchanges.argtypes = [C.c_void_p]
    c = pyoracle.Case(fx)
    c.init_rotors()
    h = lib.case_cpu_staged_hooks_install(c.h, c.nr)
    m_max = sum((c.rotor(ir).dims()["nNwake"] * (c.rotor(ir).ns + 1) + c.rotor(ir).nFwake) * c.rotor(ir).dims()["nbConvect"]
                for ir in range(c.nr))
    per_max = (m_max + world - 1) // world
    xbuf = np.full(3 * world * per_max, np.nan)          # host exchange buffer (the CPU backend's "device" memory)

    def exchange(_arg, per):
        n = 3 * per
        mine = torch.from_numpy(xbuf[rank * n:(rank + 1) * n].copy())
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        xbuf[:world * n] = torch.cat(parts).numpy()
        return 0

    cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_long)(exchange)
    lib.case_gpu_hooks_set_sharding(h, world, rank, xbuf.ctypes.data, world * per_max, C.cast(cb, C.c_void_p), None)
    c.init()
    hist = [c.force_nondim(0).copy()]
    for _ in range(nsteps):
        c.step()
        hist.append(c.force_nondim(0).copy())
    r = c.rotor(0)
    np.savez(Path(out_dir) / f"case_rank{rank}.npz", hist=np.array(hist), waN=r.waN(0), gam=r.vec(0),
             vel=r.vel(0, 0), exchanges=lib.case_gpu_hooks_exchanges(h))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,nsteps", [("katzNplotkin_AR04", 10), ("caradonna", 22)])
def test_case_path_sharded_wake_sweep_two_ranks(tmp_path, oracle, name, nsteps):
    """tests/native/case_gpu_hooks.c:staged_sweep with world = 2 on its CPU backend: each rank keeps only its slice of
    the swept velocities (the rest of its arrays is poisoned with NaN), the slices are all-gathered over gloo, every rank
    scatters the whole list.  Both ranks must end bit-identical to the single-process driver: forces, circulations, wake,
    velocity arrays; one exchange per wake sweep (fdScheme 3: 1 in the first step, 2 afterwards)."""
    import json
    import torch.multiprocessing as mp
    mp.start_processes(_case_worker, args=(2, _free_port(), str(tmp_path), name, nsteps), nprocs=2, join=True, start_method="spawn")
    fx = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    if name == "caradonna":
        fx["config"]["nt"] = 40
        fx["geom"][0].update(nNwake=12, wakeTruncateNt=18)
    a = oracle.Case(fx)
    a.init()
    hist = [a.force_nondim(0).copy()]
    for _ in range(nsteps):
        a.step()
        hist.append(a.force_nondim(0).copy())
    ra = a.rotor(0)
    for rank in range(2):
        z = np.load(tmp_path / f"case_rank{rank}.npz")
        assert np.array_equal(z["hist"], np.array(hist)), rank
        assert np.array_equal(z["gam"], ra.vec(0)) and np.array_equal(z["waN"], ra.waN(0)), rank
        assert np.array_equal(z["vel"], ra.vel(0, 0), equal_nan=False), rank
        assert int(z["exchanges"]) == 2 * nsteps - 1, (rank, int(z["exchanges"]))
