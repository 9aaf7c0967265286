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
