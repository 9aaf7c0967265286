#!/usr/bin/env python
"""bench.py -- Biot-Savart pair-interactions/s on the synthetic multirotor wake (BASELINE.json configs[4]).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --gpus N --single-process                # the same step, ONE process driving N GPUs (vlc_create_multi)
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference OpenMP path

Everything goes through the entry points the Fortran shim binds (include/volcanor_b200.h, tiers 2 / 2b): the wake is
handed over as the reference's own records (Nwake_class / Fwake_class, classdef.f90:181-220).

One "step" = one wake time step of the reference's fdScheme 3 (main.f90:1002-1115) with the wake resident on the
device(s): rotor%dissipate_wake, the wake sweep on the current wake (every convected wake node against every filament of
every rotor: libCommon.f90:114-211), Adams-Bashforth predictor + convectwake('P'), the sweep on the predicted wake,
Adams-Moulton corrector + convectwake('C').  With N > 1 the LIBRARY shards the targets of both sweeps and all-gathers
the velocity slices (ncclAllGather inside vlc_wake_sweep); work is fixed as N grows: strong scaling.
`value` = pair interactions of all GPUs / max-over-ranks device time.  `e2e` = the per-sweep hand-over the unmodified
call sites make: vlc_rotor_put_nwake / _put_fwake of every blade + vlc_vind_onNwake_byRotor / _onFwake_byRotor per
(target blade, source rotor) with pinned HOST buffers.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOPS_PER_PAIR = 76        # algorithmic flops of vf_vind as written + gam scale/accumulate
NU, VISC_COEFF = 1.8e-5, 5.0
DT_STEP = 1e-4             # s: wake nodes move by <= ~4e-3 rotor radii per step (induced velocities <= ~40 m/s)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filaments", type=int, default=int(os.environ.get("VLC_BENCH_FILAMENTS", "1000000")),
                    help="size of the synthetic wake (default 1e6; VLC_BENCH_FILAMENTS sets it for a driver-run 1e7 line)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--T", type=int, default=0, help="targets per thread of the flat kernel (0 = auto)")
    ap.add_argument("--nsplit", type=int, default=0, help="source splits (0 = auto)")
    ap.add_argument("--lat-w", type=int, default=0, help="strip width of the shared-node kernel, 1..4 (0 = default)")
    ap.add_argument("--lat-t", type=int, default=0, help="targets per thread of the shared-node kernel, 1..3 (0 = default)")
    ap.add_argument("--flat", action="store_true",
                    help="force the flat kernel on the reference's enumeration (default: shared-node lattice kernel)")
    ap.add_argument("--single-process", action="store_true",
                    help="ONE process, --gpus N devices behind one handle (vlc_create_multi): what a single-process Fortran "
                         "driver gets; the default for N > 1 is one process per GPU under torchrun (vlc_comm_init_rank)")
    ap.add_argument("--core-slope", type=float, default=0.0,
                    help="non-uniform streamwiseCoreVec: vf(1)/vf(3) core radii of column j scaled by (1 + slope*j), so that the two "
                         "copies of every interior streamwise edge differ (SURVEY C2) and the sweeps use the DUAL form of the lattice "
                         "kernel; 0 = the BASELINE workload (uniform)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(args):
    from volcanor_b200 import synth
    lats = synth.multirotor(args.filaments, seed=args.seed)
    n_src = sum(l.n_filaments() for l in lats)
    m = sum(l.targets().shape[0] for l in lats)
    name = (f"synthetic 4-rotor(2 blades)+wing wake, {n_src} filaments x {m} wake-node targets, "
            f"seed {args.seed} (BASELINE.json configs[4] at ~1e{int(round(np.log10(max(n_src, 1))))})")
    return lats, n_src, m, name


def config_of(args, name, n_src, m):
    """The `config` object: identical in our arm and in --impl reference (the driver compares them)."""
    return {"workload": name, "filaments": int(n_src), "targets": int(m), "seed": args.seed,
            "step": "one wake time step of fdScheme 3 (main.f90:1002-1115), wake resident on the device(s): dissipate_wake, "
                    "wake sweep on the current wake (all wake-node targets x all filaments), AB2 predictor + convectwake('P'), "
                    "wake sweep on the predicted wake, AM2 corrector + convectwake('C'); 2 sweeps = 2 x targets x filaments "
                    f"pair interactions; dt = {DT_STEP:g} s (the wake moves; no shedding / roll-up: fixed size)",
            "l2": "flushed every step by a 256 MiB memset inside the timed region (vlc_l2_flush)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(args, p1, p2, rvc, gam, flag, P, seconds: float):
    """C restatement of the reference OpenMP path (oracle, 'port') on the host cores, bounded sample.  Also returns
    the velocities of the sampled targets: bench.py's in-run parity check compares the GPU's against them."""
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must use the host's cores whatever launched it
    # (round 1: the reference arm ran on one core at N >= 2).  Set before libgomp is loaded by the oracle library.
    if os.environ.get("OMP_NUM_THREADS") == "1" and ("WORLD_SIZE" in os.environ or "TORCHELASTIC_RUN_ID" in os.environ):
        os.environ["OMP_NUM_THREADS"] = str(host_threads())
    from oracle import pyoracle
    pyoracle.build()
    lib = None
    tmp = tempfile.mkdtemp(prefix="vlc_native_")
    nat = pyoracle.build_native(Path(tmp))
    flags = "-O2 -march=native -fopenmp"
    if nat is not None:
        try:
            lib = pyoracle.load(path=nat)
        except OSError:
            lib = None
    if lib is None:
        lib = pyoracle.load("omp")
        flags = "-O2 -march=x86-64-v3 -fopenmp (prebuilt)"
    os.environ.setdefault("OMP_SCHEDULE", "dynamic,16")
    cores = lib.orc_num_threads()
    n = rvc.size
    probe = min(P.shape[0], max(cores * 4, 64))
    t0 = time.perf_counter()
    pyoracle.vind_flat(p1, p2, rvc, gam, flag, P[:probe], lib=lib)
    dt = time.perf_counter() - t0
    rate = probe * n / max(dt, 1e-9)
    m_s = int(min(P.shape[0], max(probe, rate * seconds / n)))
    idx = np.linspace(0, P.shape[0] - 1, m_s).astype(np.int64)
    Ps = np.ascontiguousarray(P[idx])
    t0 = time.perf_counter()
    Vs = pyoracle.vind_flat(p1, p2, rvc, gam, flag, Ps, lib=lib)
    dt = time.perf_counter() - t0
    return {"value": m_s * n / dt, "unit": "pair-interactions/s", "cores": int(cores), "kind": "port",
            "sample": f"{m_s} evenly spaced targets x all {n} filaments of the same workload, "
                      f"{dt:.1f} s, gcc {flags}, OMP_SCHEDULE={os.environ.get('OMP_SCHEDULE')}",
            "lib": lib, "m_sample": m_s, "idx": idx, "V": Vs}


def run_reference(args):
    """--impl reference: the reference's own CPU path (C restatement; no Fortran compiler exists here), all host
    threads, bounded sample per step.  Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from volcanor_b200 import synth
    lats, n_src, m, name = workload(args)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    P = synth.targets_all(lats)
    base = cpu_baseline(args, p1, p2, rvc, gam, flag, P, seconds=3.0)
    lib, m_s = base.pop("lib"), base.pop("m_sample")
    base.pop("idx")
    base.pop("V")
    from oracle import pyoracle
    idx = np.linspace(0, P.shape[0] - 1, m_s).astype(np.int64)
    Ps = np.ascontiguousarray(P[idx])
    for _ in range(args.warmup):
        pyoracle.vind_flat(p1, p2, rvc, gam, flag, Ps, lib=lib)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pyoracle.vind_flat(p1, p2, rvc, gam, flag, Ps, lib=lib)
    dt = time.perf_counter() - t0
    val = args.steps * m_s * n_src / dt
    base["value"] = val
    base["sample"] = (f"each step = {m_s} evenly spaced targets x all {n_src} filaments (bounded sample of the "
                      f"workload), C restatement of the reference OpenMP loops (libCommon.f90:132-146), {base['cores']} threads")
    out = {"impl": "reference", "metric": "biot_savart_pair_interactions_per_s", "value": val,
           "unit": "pair-interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config_of(args, name, n_src, m),
           "cpu_baseline": base,
           "e2e": {"value": val, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm

def define_rotors(ctx, rotors):
    """vlc_rotor_define + the rotor_class members the wake mutators read.  surfaceType = 2: the synthetic workload has
    wake filaments only (a non-lifting surface contributes no wing sources, classdef.f90:4437-4441)."""
    for ir, r in enumerate(rotors):
        ctx.rotor_define(ir, r["nb"], 1, r["ns"], r["nNwake"], r["nFwake"], 2)
        ctx.rotor_set_wake_params(ir, r["nb"], 0, 0, 0, 1, r["ns"], 1.0, VISC_COEFF, 0.0, 0.0)
        ctx.rotor_set_rows(ir, 1, 1)


def upload_wake(ctx, rotors, predicted=False):
    for ir, r in enumerate(rotors):
        for ib in range(r["nb"]):
            ctx.rotor_put_nwake(ir, ib, r["waN"][ib], predicted)
            if r["nFwake"]:
                ctx.rotor_put_fwake(ir, ib, r["waF"][ib], predicted)


def resident_step(ctx, nr, first):
    """main.f90:466-506 (dissipation) + :1002-1115 (fdScheme 3) with the library's device twins; tests/native/
    case_gpu_hooks.c:g_convect is the same sequence inside the reference driver."""
    C = type(ctx)
    ctx.l2_flush()
    for ir in range(nr):
        ctx.rotor_dissipate_wake(ir, DT_STEP, NU)
    ctx.wake_sweep(False)                                   # stage 1: sources 'C', targets = every convected wake node
    if first:                                               # iter == 1 (main.f90:1003-1020)
        for ir in range(nr):
            ctx.rotor_convectwake(ir, DT_STEP, "C")
            ctx.rotor_wakevel_op(ir, C.VEL_FIRST_STEP)
        ctx.wake_sweep(False)                               # keep two sweeps per step in every step of the bench
        return
    for ir in range(nr):
        ctx.rotor_wake_to_predicted(ir)
        ctx.rotor_wakevel_op(ir, C.VEL_AB2)                 # velStep = vel; vel = 0.5*(3 vel - vel1)
        ctx.rotor_convectwake(ir, DT_STEP, "P")
    ctx.wake_sweep(True)                                    # stage 2 on the predicted wake
    for ir in range(nr):
        ctx.rotor_wakevel_op(ir, C.VEL_AM2)                 # vel = (velPredicted + velStep)*0.5
        ctx.rotor_convectwake(ir, DT_STEP, "C")
        ctx.rotor_wakevel_op(ir, C.VEL_SHIFT_HISTORY)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries ONE JSON line.  Native libraries write there too (NCCL prints "NCCL version ..." on rank 0 whatever
    # NCCL_DEBUG / NCCL_DEBUG_FILE say on this image), so file descriptor 1 points at stderr until the line is printed.
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import volcanor_b200 as vb
    from volcanor_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (volcanor_b200 has no CPU fallback)")
    single = bool(args.single_process and world == 1 and args.gpus > 1)
    n_gpus = args.gpus if single else world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lats, n_src, m, name = workload(args)
    rotors = synth.rotors_from_lattices(lats)
    nr = len(rotors)
    if args.core_slope:
        for r in rotors:
            for w in r["waN"]:
                f = 1.0 + args.core_slope * np.arange(r["ns"])[:, None]
                for k in (8, 9, 24 + 8, 24 + 9):
                    w[:, :, k] *= f
    if single:
        ctx = vb.Context(devices=list(range(args.gpus)))   # one handle, N GPUs: the library owns threads + communicators
    else:
        ctx = vb.Context(local)
        if world > 1:                                      # one process per GPU: the library still owns the communicator
            uid = [vb.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init_rank(world, rank, uid[0])
    comm = ctx.comm_info()
    ctx.set_tuning(args.T, args.nsplit)
    ctx.set_shared_nodes(not args.flat)
    ctx.set_lattice_tuning(args.lat_w, args.lat_t)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- resident wake: uploaded once in the reference's record layout -----------------------------------------
    define_rotors(ctx, rotors)
    upload_wake(ctx, rotors)
    assert ctx.wake_sweep_count() == m, (ctx.wake_sweep_count(), m)
    info = [ctx.rotor_info(ir) for ir in range(nr)]
    assert sum(i["filaments"] for i in info) == n_src, (info, n_src)
    shared = all(i["shared_active"] in (1, 2) for i in info)
    dual = any(i["shared_active"] == 2 for i in info)
    fp64_peak, _ = ctx.measure_fp64_peak(20000)
    fp64_rate3, _ = ctx.measure_fp64_rate(1, 20000)    # DFMA rate with three changing register operands (informational)

    first = True
    for _ in range(max(args.warmup, 3)):
        resident_step(ctx, nr, first)
        first = False
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ctx.sweep_stats(reset=1)
    barrier()
    ctx.event_record(0)
    for _ in range(args.steps):
        resident_step(ctx, nr, False)
    ctx.event_record(1)
    barrier()
    elapsed_ms = reduce_max(ctx.event_elapsed_ms(0, 1))
    stats = ctx.sweep_stats(reset=-1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    pairs_step = 2.0 * float(m) * float(n_src)             # two sweeps per time step
    value = pairs_step * args.steps / (elapsed_ms * 1e-3)

    # every rank / member must hold the same wake after the last all-gather, and it must be finite: blade 0 of rotor 0
    w0 = ctx.rotor_get_nwake(0, 0, rotors[0]["nNwake"], rotors[0]["ns"])
    wake_finite = bool(np.isfinite(w0).all())
    moved = float(np.max(np.abs(w0[:, :, 0:3] - rotors[0]["waN"][0][:, :, 0:3])))
    ranks_consistent = True
    if world > 1:
        chk = torch.tensor([float(w0[:, :, 12:15].sum())], dtype=torch.float64, device=dev)
        hi_, lo_ = chk.clone(), chk.clone()
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        ranks_consistent = bool(torch.equal(hi_, lo_))

    # ---- same-work figure: ONE sweep of the flat kernel on the reference's enumeration, after the timed region ----
    same_work = None
    if shared:
        ctx.set_shared_nodes(False)
        ctx.wake_sweep(False)                              # packs the flat enumeration, warm
        ctx.sweep_stats(reset=1)
        ctx.wake_sweep(False)
        sf = ctx.sweep_stats(reset=-1)["bs_sweep_kernel"]
        ctx.set_shared_nodes(True)
        if sf["ms"] > 0:
            same_work = {"kernel": "bs_sweep_kernel", "launches": sf["launches"], "ms": sf["ms"],
                         "achieved": sf["pairs"] * FLOPS_PER_PAIR / (sf["ms"] * 1e-3) / 1e12,
                         "frac": sf["pairs"] * FLOPS_PER_PAIR / (sf["ms"] * 1e-3) / fp64_peak,
                         "pipe_frac": sf["fp64_instr"] * 2 / (sf["ms"] * 1e-3) / fp64_peak}

    # ---- e2e: the per-sweep hand-over of the unmodified call sites, HOST buffers (H2D + D2H inside the timed region) ----
    e2e, e2e_V, e2e_off = None, None, None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hrot = [{**r, "waN": [pin(w) for w in r["waN"]], "waF": [pin(w) for w in r["waF"]]} for r in rotors]
        ectx = ctx
        # a second set of rotors (indices nr..2nr-1) so that the resident wake above stays untouched
        for k, r in enumerate(hrot):
            ectx.rotor_define(nr + k, r["nb"], 1, r["ns"], r["nNwake"], r["nFwake"], 2)
            ectx.rotor_set_rows(nr + k, 1, 1)
        h2d = sum(8 * sum(w.size for w in r["waN"]) + 8 * sum(w.size for w in r["waF"]) for r in hrot)
        tgt_bytes = 0
        out_near = [[np.zeros((r["ns"] + 1, r["nNwake"], 3)) for _ in range(r["nb"])] for r in hrot]
        out_far = [[np.zeros((r["nFwake"], 3)) for _ in range(r["nb"])] for r in hrot]
        # the function result of one call (libCommon.f90:124, :186): one array per shape, reused like the caller's temporary
        tmp_near = {(r["ns"], r["nNwake"]): np.zeros((r["ns"] + 1, r["nNwake"], 3)) for r in hrot}
        tmp_far = {r["nFwake"]: np.zeros((r["nFwake"], 3)) for r in hrot}

        e2e_host = {"put_s": 0.0, "vind_s": 0.0}

        def e2e_sweep():
            """main.f90:814-841 as the shim executes it: refresh the device copies of every blade's records, then per
            target blade and source rotor one vind_onNwake_byRotor and one vind_onFwake_byRotor, summed on the host."""
            nonlocal tgt_bytes
            tgt_bytes = 0
            tp = time.perf_counter()
            for k, r in enumerate(hrot):
                for ib in range(r["nb"]):
                    ectx.rotor_put_nwake(nr + k, ib, r["waN"][ib])
                    if r["nFwake"]:
                        ectx.rotor_put_fwake(nr + k, ib, r["waF"][ib])
            tv = time.perf_counter()
            e2e_host["put_s"] += tv - tp
            for k, r in enumerate(hrot):
                for ib in range(r["nb"]):
                    vn, vf = out_near[k][ib], out_far[k][ib]
                    vn[...] = 0.0
                    vf[...] = 0.0
                    for j in range(nr):
                        vn += ectx.vind_onNwake_byRotor(nr + j, r["waN"][ib], r["nNwake"], r["ns"], r["nNwake"],
                                                        out=tmp_near[(r["ns"], r["nNwake"])])
                        tgt_bytes += 24 * r["nNwake"] * (r["ns"] + 1)
                        if r["nFwake"]:
                            vf += ectx.vind_onFwake_byRotor(nr + j, r["waF"][ib], r["nFwake"], out=tmp_far[r["nFwake"]])
                            tgt_bytes += 24 * r["nFwake"]
            e2e_host["vind_s"] += time.perf_counter() - tv

        def e2e_step():
            e2e_sweep()        # current wake
            e2e_sweep()        # predicted wake (the same records stand in for it: same sizes, same work)

        e2e_step()
        barrier()
        e2e_host["put_s"] = e2e_host["vind_s"] = 0.0
        ectx.sweep_stats(1)
        w0_ = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        w = reduce_max(time.perf_counter() - w0_)
        est = ectx.sweep_stats(-1)
        e2e = {"value": pairs_step * args.steps / w, "unit": "pair-interactions/s",
               "h2d_bytes_per_step": int(2 * (h2d + tgt_bytes)), "d2h_bytes_per_step": int(2 * tgt_bytes),
               "call": "per wake sweep (2 per time step): vlc_rotor_put_nwake + vlc_rotor_put_fwake of every blade (the reference's "
                       "400-byte Nwake_class / 104-byte Fwake_class records, pinned host memory), then per (target blade, source "
                       "rotor) vlc_vind_onNwake_byRotor + vlc_vind_onFwake_byRotor (libCommon.f90:114-211) with host target records "
                       "in and host velocity arrays out, summed over source rotors on the host like main.f90:817-826",
               "calls_per_step": int(2 * sum(r["nb"] for r in hrot) * nr * 2), "ms_per_step": 1e3 * w / args.steps,
               "breakdown_ms_per_step": {"host_in_put_calls": 1e3 * e2e_host["put_s"] / args.steps,
                                         "host_in_vind_calls": 1e3 * e2e_host["vind_s"] / args.steps,
                                         "device_in_sweeps": sum(v["sweep_ms"] for v in est.values()) / args.steps,
                                         "sweeps": int(sum(v["launches"] for v in est.values()) / args.steps)}}
        # velocities of all targets in the order of synth.targets_all (per blade: near nodes, then far nodes)
        e2e_V = np.concatenate([np.concatenate([out_near[k][ib].reshape(-1, 3), out_far[k][ib]])
                                for k, r in enumerate(hrot) for ib in range(r["nb"])])

    if rank != 0:
        if world > 1:
            ctx.close()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: one CUDA-event pair per launch on the launching stream, over the timed region ----
    peak = fp64_peak / 1e12
    kernel = "bs_lattice_kernel" if shared else "bs_sweep_kernel"
    st = dict(stats[kernel])
    if shared and dual:     # the library counts the merged form's instructions (the host never reads the flag): 58 W instead of 50 W
        W_ = int(info[0]["strip_width"])
        extra = st["fp64_instr"] * ((11 * (W_ + 1) + 58 * W_) / (11 * (W_ + 1) + 50 * W_) - 1.0)
        st["fp64_instr"] += extra
        st["sweep_fp64_instr"] += extra
    kern_ms_total, n_launch = st["ms"], max(st["launches"], 1)
    achieved = st["pairs"] * FLOPS_PER_PAIR / (kern_ms_total * 1e-3) / 1e12 if kern_ms_total > 0 else 0.0
    note = ("shared-node lattice kernel: every lattice node evaluated once per target and every interior edge once with the "
            "merged strength of its two rings -- the reference's ring-by-ring sum regrouped, so `frac` (the reference's 76 flop x "
            "4 filaments per ring) exceeds what the kernel executes ((11(W+1)+50W)/W = 63.75 FP64 instructions per (target, ring) "
            "at W = 4) and may exceed 1; `pipe_frac` = issued FP64 instructions x 2 / peak is the hardware utilisation; "
            "`same_work` = the flat kernel on the reference's own enumeration, one sweep after the timed region"
            if shared else "flat kernel on the reference's enumeration: frac = algorithmic 76 flop/pair, pipe_frac = 43 issued FP64 instr/pair")
    # DRAM traffic of the dominant kernel: not measurable outside a profiler; quoted from the committed ncu --set full capture
    # of THIS command (profiles/r03z_bs_sweep_full.md), and only for the workload and launch shape it was taken on
    prof = ROOT / "profiles" / "r03z_bs_sweep_full.md"
    traffic, traffic_source = None, None
    if shared and not dual and n_gpus == 1 and n_src == 1000192 and args.lat_w in (0, 4) and args.lat_t in (0, 2) and prof.exists():
        import re
        txt = prof.read_text()
        rd = re.search(r"dram__bytes_read\.sum \| ([0-9.]+) \| Mbyte", txt)
        wr = re.search(r"dram__bytes_write\.sum \| ([0-9.]+) \| Mbyte", txt)
        if rd and wr:
            traffic = (float(rd.group(1)) + float(wr.group(1))) * 1e6
            traffic_source = ("profiles/r03z_bs_sweep_full.md: ncu --set full of this command (same workload, same launch shape), "
                              "dram__bytes_read.sum + dram__bytes_write.sum of one bs_lattice_kernel launch; not re-measured in this "
                              "run. Algorithmic bytes per launch: 36.0 MB of strip records (62 536 x 576 B) + 6.2 MB of targets "
                              "read + 74 MB of source-split partial sums written (12 x 6.2 MB), most of which stay in L2")
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_source,
                "kernel": kernel, "launches": st["launches"], "kernel_ms": kern_ms_total / n_launch,
                "kernel_ms_per_step": kern_ms_total / args.steps,
                "kernel_share_of_step": kern_ms_total / elapsed_ms if elapsed_ms > 0 else None,
                "pairs_per_launch": st["pairs"] / n_launch, "flops_per_pair": FLOPS_PER_PAIR,
                "pipe_frac": st["fp64_instr"] * 2 / (kern_ms_total * 1e-3) / fp64_peak if kern_ms_total > 0 else 0.0,
                "sweep": {"ms": st["sweep_ms"] / n_launch, "pairs": st["sweep_pairs"] / n_launch,
                          "pipe_frac": st["sweep_fp64_instr"] * 2 / (st["sweep_ms"] * 1e-3) / fp64_peak if st["sweep_ms"] > 0 else 0.0,
                          "note": "whole sweep = dominant kernel + flat remainder (last columns, horseshoe corrections, far wakes; "
                                  "low-priority side stream: its CTAs share SMs with the dominant kernel inside that kernel's window) "
                                  "+ fixed-order reduce of the source-split partial sums"},
                "same_work": same_work, "frac_same_work": same_work["frac"] if same_work else None,
                "dfma_3reg_tflops": fp64_rate3 / 1e12,
                "peak_source": "measured live: vlc_measure_fp64_peak (register-resident DFMA chains, all SMs); MEASURED_PEAKS.json has no "
                               "FP64 entry; nominal 148*64*2*1.965 GHz = 37.2; dfma_3reg_tflops = the same measurement with three "
                               "distinct changing register operands per DFMA, the practical ceiling of register-fed FP64 code",
                "timing": "one CUDA event pair per launch on the launching stream (vlc_sweep_stats), summed over the timed region; "
                          "per GPU (rank 0's slice of the targets)",
                "note": "compute-bound pairwise N-body on the FP64 pipe (no tensor cores by construction); " + note}
    cpu, checks = None, {"ranks_hold_identical_wake": ranks_consistent, "wake_finite": wake_finite,
                         "wake_moved_max_abs": moved}
    if n_gpus == 1 and not args.no_cpu_baseline:
        p1, p2, rvc, gam, flag = synth.flatten_all(lats)
        cpu = cpu_baseline(args, p1, p2, rvc, gam, flag, synth.targets_all(lats), args.cpu_seconds)
        cpu.pop("lib")
        cpu.pop("m_sample")
        idx, Vs = cpu.pop("idx"), cpu.pop("V")
        if e2e_V is not None:
            # in-run parity: the host-path velocities of the sampled targets against the CPU restatement, per target,
            # relative to the target's own velocity magnitude budget max|V| of the sample (no long-double pass here)
            err = np.abs(e2e_V[idx] - Vs).max(axis=1)
            scale = np.maximum(np.abs(Vs).max(axis=1), 1e-300)
            checks["sampled_oracle"] = {"targets": int(idx.size), "max_abs_err": float(err.max()),
                                        "max_err_over_own_velocity": float((err / scale).max()),
                                        "max_err_over_sample_velocity_scale": float(err.max() / np.abs(Vs).max()),
                                        "note": "e2e (host-path) velocities of the CPU baseline's sampled targets vs the C restatement "
                                                "of the reference; tests/ measure against sum|terms| per target (1e-12)"}

    out = {"metric": "biot_savart_pair_interactions_per_s", "value": value, "unit": "pair-interactions/s",
           "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_of(args, name, n_src, m),     # the same object in both arms (the driver compares them)
           "run": {"parallelism": (f"target-sharded x{n_gpus} behind the C ABI, sources replicated, 1 all-gather of velocity slices per "
                                      f"wake sweep (2 per step) inside vlc_wake_sweep; transport {comm['transport']}; "
                                      + ("ONE process, one worker thread per GPU (vlc_create_multi)" if single else
                                         "one process per GPU, library-owned communicator (vlc_comm_init_rank)" if world > 1 else
                                         "single GPU")),
                      "launch": "eager: every kernel launched per step on the library's stream(s)",
                      "tuning": {"T": args.T, "nsplit": args.nsplit},
                      "sources": ({"form": "shared-node lattice" + (", DUAL form (two core radii per streamwise edge)" if dual else ""),
                                   "core_slope": args.core_slope, "strip_width": int(info[0]["strip_width"]),
                                   "strip_records": int(sum(i["lattice_records"] for i in info)),
                                   "remainder_filaments": int(sum(i["remainder_filaments"] for i in info)),
                                   "source_rotors": nr} if shared else {"form": "flat reference enumeration", "source_rotors": nr})},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "timesteps_per_s": args.steps / (elapsed_ms * 1e-3), "sweeps_per_step": 2, "checks": checks,
           "fp64_peak_measured_tflops": peak}
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)          # NCCL teardown messages, if any
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
