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
