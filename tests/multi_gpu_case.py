#!/usr/bin/env python
"""One process per GPU on a reference case (test infrastructure; launched by torchrun):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/multi_gpu_case.py --case katzNplotkin_AR04 --steps 160

Every rank runs the same driver (oracle/vlc_case.c) with the wake resident on ITS device (C ABI tier 2b through
tests/native/case_gpu_hooks.c); the two wake sweeps of a time step are sharded by target nodes: rank r sweeps targets
[r*per, (r+1)*per) against the full source set (vlc_wake_sweep_slice), the velocity slices are all-gathered -- NCCL over
NVLink when every rank has its own GPU, staged through the host with gloo when the ranks share one (so the logic is
testable on a single-GPU box) -- and every rank scatters the complete list (vlc_wake_sweep_scatter) and convects its own
copy of the wake.  One exchange per predictor and one per corrector stage (BASELINE.json north_star).

Rank 0 prints one JSON line: wall time, timesteps/s, exchanges, whether all ranks hold bitwise identical force
histories, and (cases with a golden file) the largest deviation from it in units of its 7th printed digit.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="katzNplotkin_AR04")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--short-caradonna", action="store_true", help="nt 40, nNwake 12 (roll-up inside a short window)")
    ap.add_argument("--cp", action="store_true",
                    help="also run the collocation-point stage (RHS, solve, map_gam, loads: C ABI tier 2c) on every rank's device")
    ap.add_argument("--lib-comm", action="store_true",
                    help="the LIBRARY owns the communicator (vlc_comm_init_rank): the driver calls plain vlc_wake_sweep, which "
                         "shards the targets and all-gathers the velocity slices itself (ncclAllGather); needs one GPU per rank")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import volcanor_b200 as vb
    from oracle import pyoracle
    from tests.test_gpu_resident import _resident_hooks

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    own_gpu = torch.cuda.device_count() >= world
    device = local if own_gpu else 0
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if own_gpu else "gloo",
                                **({"device_id": torch.device("cuda", device)} if own_gpu else {}))

    pyoracle.build()
    fx = json.loads((ROOT / "tests" / "golden" / f"{args.case}.json").read_text())
    if args.short_caradonna:
        fx["config"]["nt"] = 40
        fx["geom"][0]["nNwake"] = 12
    c = pyoracle.Case(fx)
    ctx = vb.Context(device)
    lib, h = _resident_hooks(c, ctx)                     # init_rotors + resident hook table
    lib.case_gpu_hooks_set_sharding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    lib.case_gpu_hooks_exchanges.restype = C.c_long
    lib.case_gpu_hooks_exchanges.argtypes = [C.c_void_p]
    if args.cp:                                          # O(nc*ns) stage, replicated on every rank like the wake mutators
        lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
        if lib.case_hooks_enable_cp(h) != 0:
            raise RuntimeError(f"case_hooks_enable_cp: {ctx.lib.vlc_last_error(ctx.h)}")

    m_max = 0
    for ir in range(c.nr):
        d = c.rotor(ir).dims()
        if d["nNwake"] > 0:
            m_max += (d["nNwake"] * (d["ns"] + 1) + d["nFwake"]) * d["nbConvect"]
    per_max = (m_max + world - 1) // world
    xbuf = torch.zeros(3 * world * max(per_max, 1), dtype=torch.float64, device=torch.device("cuda", device))
    errors = []

    def exchange(_arg, per):
        try:
            n = 3 * per
            t = xbuf[:world * n]
            mine = t[rank * n:(rank + 1) * n]
            if own_gpu:
                dist.all_gather_into_tensor(t, mine.clone())
            else:                                        # ranks share one GPU: stage through the host
                s = mine.cpu()
                parts = [torch.empty_like(s) for _ in range(world)]
                dist.all_gather(parts, s)
                t.copy_(torch.cat(parts))
            torch.cuda.synchronize()
            return 0
        except Exception as e:                           # noqa: BLE001 -- reported by rank 0 below
            errors.append(repr(e))
            return 9

    CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_long)
    cb = CB(exchange)
    if world > 1 and args.lib_comm:
        if not own_gpu:
            raise SystemExit("--lib-comm needs one GPU per rank (NCCL)")
        uid = [vb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(world, rank, uid[0])
        assert ctx.comm_info() == {"world": world, "rank": rank, "transport": "nccl"}
    elif world > 1:
        lib.case_gpu_hooks_set_sharding(h, world, rank, xbuf.data_ptr(), world * per_max, C.cast(cb, C.c_void_p), None)

    c.init()
    nt = c.config.nt if args.steps <= 0 else min(args.steps, c.config.nt)
    hist = [c.force_nondim(0).copy()]
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ok = True
    for it in range(nt):
        try:
            c.step()
        except RuntimeError as e:
            errors.append(f"step {it + 1}: {e}; {ctx.lib.vlc_last_error(ctx.h)}")
            ok = False
            break
        hist.append(c.force_nondim(0).copy())
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    hist = np.array(hist)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (hist.tobytes(), wall, errors))
    else:
        gathered = [(hist.tobytes(), wall, errors)]
    if rank == 0:
        identical = all(g[0] == gathered[0][0] for g in gathered)
        out = {"case": args.case, "world": world, "exchange": ("nccl inside the library" if args.lib_comm else "nccl" if own_gpu else "gloo via host") if world > 1 else "none",
               "steps": len(hist) - 1, "wall_s": max(g[1] for g in gathered), "ok": ok and not any(g[2] for g in gathered),
               "errors": [e for g in gathered for e in g[2]][:3], "ranks_identical": identical,
               "exchanges": int(lib.case_gpu_hooks_exchanges(h)) if not args.lib_comm else None, "final_CT_or_CL": float(hist[-1, 0]),
               "cp_stage_on_device": bool(args.cp)}
        out["timesteps_per_s"] = out["steps"] / out["wall_s"] if out["wall_s"] > 0 else 0.0
        if "ref_ForceNonDim" in fx and not args.short_caradonna:
            ref = np.array(fx["ref_ForceNonDim"]["rows"])[:len(hist)]
            ulp = 10.0 ** (np.floor(np.log10(np.abs(ref[:, 1]))) - 6)
            out["golden_max_dev_7th_digit"] = float(np.max(np.abs(hist[:len(ref), 0] - ref[:, 1]) / ulp))
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
