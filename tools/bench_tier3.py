#!/usr/bin/env python
"""bench.py -- Biot-Savart pair-interactions/s on the synthetic multirotor wake (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference OpenMP path

One "step" = one wake-convection stage of the hot path at the named size: re-pack the filament set from
the device-resident wake lattices, sweep every convected wake node (targets) against every filament
(sources) with the sm_100a kernel, convect the nodes, and (N > 1) all-gather the updated node slices
over NCCL and scatter them back into the lattices.  Work is fixed as N grows (targets are sharded):
strong scaling.  `value` = total pair interactions of all ranks / max-over-ranks device time.
Prints ONE JSON line on rank 0.
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

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

FLOPS_PER_PAIR = 76        # algorithmic flops of vf_vind as written + gam scale/accumulate
PIPE_INSTR_PER_PAIR = {0: 43, 1: 41}   # FP64-pipe instructions per pair in bs_sweep_kernel (SASS count; full / fast)
def pipe_instr_per_record(W):           # ... per (target, strip record of width W) in bs_lattice_kernel: W+1 nodes, 2W edges
    return 11 * (W + 1) + 50 * W


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filaments", type=int, default=1_000_000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--T", type=int, default=0, help="targets per thread (0 = auto)")
    ap.add_argument("--nsplit", type=int, default=0, help="source splits (0 = auto)")
    ap.add_argument("--precision", type=int, default=0, choices=[0, 1],
                    help="0 = full (third-order rsqrt, default), 1 = fast (second order, pair error <= 6.4e-13)")
    ap.add_argument("--lat-w", type=int, default=0, help="strip width of the shared-node kernel, 1..4 (0 = default)")
    ap.add_argument("--lat-t", type=int, default=0, help="targets per thread of the shared-node kernel, 1..3 (0 = default)")
    ap.add_argument("--flat", action="store_true",
                    help="force the flat kernel on the reference's enumeration (default: shared-node lattice kernel)")
    ap.add_argument("--graph", action="store_true",
                    help="1 GPU only: capture the time step in a CUDA graph and replay it (launch-bound small wakes: ~215 "
                         "launches per step); the per-kernel roofline is then taken from one eager step after the timed region")
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


def cpu_baseline(args, p1, p2, rvc, gam, flag, P, seconds: float):
    """C restatement of the reference OpenMP path (oracle, 'port') on the host cores, bounded sample."""
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
    pyoracle.vind_flat(p1, p2, rvc, gam, flag, Ps, lib=lib)
    dt = time.perf_counter() - t0
    return {"value": m_s * n / dt, "unit": "pair-interactions/s", "cores": int(cores), "kind": "port",
            "sample": f"{m_s} evenly spaced targets x all {n} filaments of the same workload, "
                      f"{dt:.1f} s, gcc {flags}, OMP_SCHEDULE={os.environ.get('OMP_SCHEDULE')}",
            "lib": lib, "m_sample": m_s}


def run_reference(args):
    """--impl reference: the reference's own CPU path (C restatement; no Fortran compiler exists here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from volcanor_b200 import synth
    lats, n_src, m, name = workload(args)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    P = synth.targets_all(lats)
    base = cpu_baseline(args, p1, p2, rvc, gam, flag, P, seconds=3.0)
    lib, m_s = base.pop("lib"), base.pop("m_sample")
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
                      f"workload), C restatement of the reference OpenMP loops (libCommon.f90:132-146)")
    out = {"impl": "reference", "metric": "biot_savart_pair_interactions_per_s", "value": val,
           "unit": "pair-interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": name, "filaments": n_src, "targets": m},
           "cpu_baseline": base,
           "e2e": {"value": val, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lats, n_src, m, name = workload(args)
    ctx = vb.Context(local)
    ctx.set_tuning(args.T, args.nsplit)
    ctx.set_precision(args.precision)
    ctx.set_shared_nodes(not args.flat)
    ctx.set_lattice_tuning(args.lat_w, args.lat_t)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- device-resident wake state (node-indexed SoA per lattice): current ('C') and predicted ('P') node sets ----
    d = []
    for l in lats:
        nodes = torch.from_numpy(l.nodes).to(dev)
        far = torch.from_numpy(l.far_nodes).to(dev) if l.F > 0 else None
        d.append({"R": l.R, "S": l.S, "F": l.F, "nodes": nodes, "nodesP": nodes.clone(),
                  "gam": torch.from_numpy(l.gam).to(dev), "rvc4": torch.from_numpy(l.rvc4).to(dev),
                  "far": far, "farP": far.clone() if far is not None else None,
                  "gamF": torch.from_numpy(l.gamF).to(dev) if l.F > 0 else None,
                  "rvcF": torch.from_numpy(l.rvcF).to(dev) if l.F > 0 else None})
    # target list = convected nodes of every lattice (+ far-chain nodes), padded to world * per
    from volcanor_b200.sharding import TargetShard, allgather_slices
    shard = TargetShard(m, world, rank)
    per, lo, hi, m_loc = shard.per, shard.lo, shard.hi, shard.count
    f64 = dict(dtype=torch.float64, device=dev)
    P_all, Pp_all = torch.zeros(shard.padded, 3, **f64), torch.zeros(shard.padded, 3, **f64)
    V, Vp, V1, Vw = (torch.zeros(max(per, 1), 3, **f64) for _ in range(4))   # this rank's slice: vel, predicted, previous, work
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dt_step, nu, visc_coeff = 1e-9, 1.8e-5, 5.0
    state = {"first": True}

    def gather_targets(key, far_key, dst):
        off = 0
        for L in d:
            k = L["R"] * (L["S"] + 1)
            ctx.lattice_targets_dev(L["R"], L["S"], L[key], dst[off:off + k])
            off += k
            if L["F"] > 0:
                dst[off:off + L["F"]].copy_(L[far_key][1:])
                off += L["F"]

    def scatter_targets(key, far_key, src):
        off = 0
        for L in d:
            k = L["R"] * (L["S"] + 1)
            ctx.lattice_scatter_dev(L["R"], L["S"], L[key], src[off:off + k])
            off += k
            if L["F"] > 0:
                L[far_key][1:].copy_(src[off:off + L["F"]])
                off += L["F"]

    def pack(set_=0, key="nodes", far_key="far"):
        for i, L in enumerate(d):
            ctx.pack_lattice_dev(set_, i > 0, L["R"], L["S"], L[key], L["gam"], L["rvc4"], L["F"], L[far_key],
                                 L["gamF"], L["rvcF"])

    ev_k0, ev_k1 = [], []

    def sweep(set_, P, out, record):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        if m_loc > 0:
            ctx.vind_dev(set_, m_loc, P[lo:hi], out)       # THE sweep: m_loc targets x n_src filaments
        if record:
            e1.record(stream)
            ev_k0.append(e0)
            ev_k1.append(e1)

    def step(record=False):
        """One wake time step of the reference's fdScheme 3 (main.f90:1002-1115) on device-resident state:
        core growth, sweep on the current wake, Adams-Bashforth predictor, sweep on the predicted wake,
        Adams-Moulton corrector; one all-gather of node positions per stage."""
        flush.zero_()                                      # L2 flush (inside the timed region, ~0.05 ms)
        for L in d:                                        # rotor_dissipate_wake (classdef.f90:4356-4408)
            ctx.dissipate_lattice_dev(L["R"], L["S"], L["rvc4"], L["gam"], visc_coeff, nu, 0.0, dt_step)
        pack(0, "nodes", "far")                            # sources 'C' from the current lattices
        gather_targets("nodes", "far", P_all)
        sweep(0, P_all, V, record)                         # stage 1
        if state["first"]:
            V1.copy_(V)                                    # iter == 1: plain convection (main.f90:1003-1020)
            state["first"] = False
        if m_loc > 0:
            ctx.ab2_dev(m_loc, V, V1, Vw)                  # vel = 0.5*(3 vel - vel1)      (main.f90:1032-1034)
            Pp_all[lo:hi].copy_(P_all[lo:hi])
            ctx.convect_dev(m_loc, Pp_all[lo:hi], Vw, dt_step)   # convectwake('P')
        allgather_slices(Pp_all, shard)                    # exchange 1: predicted node positions (NCCL)
        scatter_targets("nodesP", "farP", Pp_all)
        pack(1, "nodesP", "farP")                          # sources 'P'
        sweep(1, Pp_all, Vp, record)                       # stage 2 on the predicted wake
        if m_loc > 0:
            ctx.am2_dev(m_loc, Vp, V, Vw)                  # vel = (velPredicted + velStep)*0.5  (main.f90:1094-1096)
            ctx.convect_dev(m_loc, P_all[lo:hi], Vw, dt_step)    # convectwake('C')
        allgather_slices(P_all, shard)                     # exchange 2: corrected node positions
        scatter_targets("nodes", "far", P_all)
        V1.copy_(V)                                        # vel1 = velStep                  (main.f90:1105-1106)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pack()
    assert ctx.num_sources(0) == n_src, (ctx.num_sources(0), n_src)
    info = ctx.set_info(0)
    shared = info["shared_active"] == 1
    fp64_peak, _ = ctx.measure_fp64_peak(20000)
    fp64_rate3, _ = ctx.measure_fp64_rate(1, 20000)    # DFMA rate with three changing register operands (informational)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    use_graph = bool(args.graph and world == 1)
    graph = None
    if use_graph:
        # The whole step (library launches on the context's stream, its side-stream fork/join, torch copies) is captured
        # once and replayed: one cudaGraphLaunch per time step instead of ~215 kernel launches.
        gstream = torch.cuda.Stream()
        gstream.wait_stream(stream)
        ctx.set_stream(gstream.cuda_stream)
        launches_before = ctx.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=gstream):
            step()
        launches_per_step = ctx.launch_count - launches_before
        stream = gstream
        with torch.cuda.stream(gstream):
            graph.replay()                                   # warm replay
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    if use_graph:
        with torch.cuda.stream(stream):
            for _ in range(args.steps):
                graph.replay()
    else:
        for _ in range(args.steps):
            step(record=True)
    t1.record(stream)
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    launches = (launches_per_step * args.steps) if use_graph else (ctx.launch_count - launches0)
    clocks = sampler.stop() if rank == 0 else None
    if use_graph:                                            # per-kernel timings for the roofline: one eager step
        with torch.cuda.stream(stream):
            step(record=True)
        torch.cuda.synchronize()
    sweep_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)])) if ev_k0 else 0.0
    # dominant kernel alone (CUDA events recorded by the library on the launching stream around that launch, last step)
    main_ms, total_ms = ctx.last_sweep_ms() if m_loc > 0 else (0.0, 0.0)
    kern_ms = sweep_ms * (main_ms / total_ms) if total_ms > 0 else sweep_ms
    tt = torch.tensor([elapsed_ms, kern_ms, sweep_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms, kern_ms_max, sweep_ms_max = float(tt[0]), float(tt[1]), float(tt[2])
    # every rank must hold the same wake after the last all-gather (bitwise): max - min over ranks of a position checksum
    chk = P_all[:m].sum(dim=0)
    ranks_consistent = True
    if world > 1:
        hi_, lo_ = chk.clone(), chk.clone()
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        ranks_consistent = bool(torch.equal(hi_, lo_))
    wake_finite = bool(torch.isfinite(chk).all())
    pairs_step = 2.0 * float(m) * float(n_src)             # two sweeps per time step
    value = pairs_step * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        P_host = synth.targets_all(lats)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hl = [{"R": l.R, "S": l.S, "F": l.F, "nodes": pin(l.nodes), "gam": pin(l.gam), "rvc4": pin(l.rvc4),
               "far": pin(l.far_nodes) if l.F > 0 else None, "gamF": pin(l.gamF) if l.F > 0 else None,
               "rvcF": pin(l.rvcF) if l.F > 0 else None} for l in lats]
        h2d = sum(8 * (L["nodes"].numel() + L["gam"].numel() + L["rvc4"].numel()
                       + (L["far"].numel() + 2 * L["F"] if L["F"] > 0 else 0)) for L in hl)
        hP = pin(P_host[lo:hi]) if m_loc > 0 else None
        hV = torch.empty(max(m_loc, 1), 3, dtype=torch.float64).pin_memory()

        def e2e_step():
            # the caller-facing C-ABI calls with HOST buffers, twice per time step (current and predicted wake):
            # wake lattices in, velocities of this rank's targets out
            for stage in range(2):
                for i, L in enumerate(hl):
                    ctx.pack_lattice(2, i > 0, L["R"], L["S"], L["nodes"], L["gam"], L["rvc4"], L["F"], L["far"],
                                     L["gamF"], L["rvcF"])
                if m_loc > 0:
                    ctx.vind_into(2, m_loc, hP, hV)

        for _ in range(2):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        w = time.perf_counter() - w0
        tw = torch.tensor([w], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs_step * args.steps / float(tw[0]), "unit": "pair-interactions/s",
               "h2d_bytes_per_step": int(2 * (h2d + 24 * m_loc)), "d2h_bytes_per_step": int(2 * 24 * m_loc),
               "call": "per time step 2 x [vlc_pack_lattice (host wake lattices of all blades) + vlc_vind (host targets -> "
                       "host velocities)], pinned buffers, per rank: all sources, its target slice",
               "ms_per_step": 1e3 * float(tw[0]) / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    peak = fp64_peak / 1e12
    if shared:
        # bs_lattice_kernel covers the 4 ring filaments of every near-wake ring; the remainder kernel the rest
        kernel = "bs_lattice_kernel"
        pairs_launch = float(m_loc) * 4.0 * float(sum(l.R * l.S for l in lats))
        issued = float(m_loc) * float(info["lattice_records"]) * pipe_instr_per_record(info["strip_width"])
        note = ("shared-node lattice kernel: every lattice node evaluated once per target and every interior edge once "
                "with the merged strength of its two rings -- the reference's ring-by-ring sum regrouped, so frac counts "
                "the reference's 76 flop x 4 filaments per ring while the kernel issues (11(W+1)+50W)/W = "
                "72 / 66.5 / 64.7 / 63.75 FP64 instructions per (target, ring) for strip width W = 1..4: frac may exceed 1; pipe_frac = issued FP64 instructions vs the pipe's peak")
    else:
        kernel = "bs_sweep_kernel"
        pairs_launch = float(m_loc) * float(n_src)
        issued = pairs_launch * PIPE_INSTR_PER_PAIR[args.precision]
        note = ("flat kernel on the reference's enumeration: frac = algorithmic 76 flop/pair, pipe_frac = issued FP64 "
                "instr/pair (43 full, 41 fast)")
    achieved = pairs_launch * FLOPS_PER_PAIR / (kern_ms_max * 1e-3) / 1e12 if kern_ms_max > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch from the committed ncu --set full capture of this same command
    # (profiles/r01i_bs_sweep_full.md: dram__bytes_read.sum 37.57 MB + dram__bytes_write.sum 29.13 MB; the strip records
    # are 28.0 MB, targets 6.2 MB, the 12 source-split partial sums 74 MB, more than half of them absorbed by L2): no
    # wasted re-reads -- 0.003 % of the HBM peak.  Only quoted for the workload and launch shape the capture was taken on.
    traffic = 66.70e6 if (shared and world == 1 and args.filaments == 1_000_000 and args.lat_w in (0, 4)
                          and args.lat_t in (0, 2) and args.nsplit == 0) else None
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "kernel": kernel, "kernel_ms": kern_ms_max, "sweep_ms": sweep_ms_max,
                "pairs_per_launch": pairs_launch, "flops_per_pair": FLOPS_PER_PAIR,
                "pipe_frac": issued * 2 / (kern_ms_max * 1e-3) / fp64_peak if kern_ms_max > 0 else 0.0,
                "dfma_3reg_tflops": fp64_rate3 / 1e12,
                "pipe_frac_of_3reg_rate": issued * 2 / (kern_ms_max * 1e-3) / fp64_rate3 if kern_ms_max > 0 else 0.0,
                "peak_source": "measured live: vlc_measure_fp64_peak (register-resident DFMA chains, all SMs); "
                               "MEASURED_PEAKS.json has no FP64 entry; nominal 148*64*2*1.965 GHz = 37.2; dfma_3reg_tflops = the same "
                               "measurement with three distinct changing register operands per DFMA (vlc_measure_fp64_rate "
                               "pattern 1), the practical ceiling of register-fed FP64 code",
                "note": "compute-bound pairwise N-body on the FP64 pipe (no tensor cores by construction); " + note}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        p1, p2, rvc, gam, flag = synth.flatten_all(lats)
        cpu = cpu_baseline(args, p1, p2, rvc, gam, flag, synth.targets_all(lats), args.cpu_seconds)
        cpu.pop("lib")
        cpu.pop("m_sample")

    out = {"metric": "biot_savart_pair_interactions_per_s", "value": value, "unit": "pair-interactions/s",
           "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": name, "filaments": n_src, "targets": m, "targets_per_rank": per,
                      "step": "one wake time step of fdScheme 3 on device-resident state: core growth, pack, sweep on the "
                              "current wake (targets slice x all filaments), AB2 predictor + all-gather, pack, sweep on the "
                              "predicted wake, AM2 corrector + all-gather",
                      "l2": "flushed every step by a 256 MiB memset inside the timed region",
                      "parallelism": f"target-sharded x{world}, sources replicated, 1 NCCL all-gather per stage (2 per step)",
                      "launch": ("the step is captured once in a CUDA graph and replayed (one graph launch per time step); "
                                 "roofline kernel times from one eager step after the timed region") if use_graph
                                else "eager: one stream, every kernel launched per step",
                      "tuning": {"T": args.T, "nsplit": args.nsplit},
                      "sources": ({"form": "shared-node lattice", "strip_width": int(info["strip_width"]),
                                   "strip_records": int(info["lattice_records"]),
                                   "remainder_filaments": int(info["remainder_filaments"])} if shared
                                  else {"form": "flat reference enumeration"}),
                      "precision": ["full: third-order rsqrt refinement, pair error ~1e-16",
                                    "fast: second-order rsqrt refinement, pair error <= 6.4e-13"][args.precision]},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "timesteps_per_s": args.steps / (elapsed_ms * 1e-3), "sweeps_per_step": 2,
           "checks": {"ranks_hold_identical_wake": ranks_consistent, "wake_finite": wake_finite},
           "fp64_peak_measured_tflops": peak}
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)          # NCCL teardown messages, if any
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
