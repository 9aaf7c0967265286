"""Collocation-point stage of a LARGE case through a group handle: vlc_rotor_calc_RHS of a lifting rotor with `ncp`
collocation points against the synthetic 10^6-filament wake, on one GPU, on a group that replicates the sweeps
(VLC_RHS_SHARE_MIN_PAIRS=-1) and on a group that shares the source splits out (default threshold).
Usage (2+ GPUs): PYTHONPATH=. python tools/rhs_share_probe.py [filaments] [gpus]"""
import os
import sys
import time

import numpy as np

import volcanor_b200 as vb
from volcanor_b200 import synth

KREC, KCP, KNCAP = 104, 64, 67


def wing_records(nc, ns, rng):
    """nc x ns panels of a flat plate at z = 0.3 R below the hub: ring corners, CP, nCap = +z (the fields calc_RHS reads)."""
    w = np.zeros((ns, nc, KREC))
    x = np.linspace(0.2, 1.0, ns + 1)
    y = np.linspace(-0.05, 0.05, nc + 1)
    for j in range(ns):
        for i in range(nc):
            c = [(x[j], y[i]), (x[j], y[i + 1]), (x[j + 1], y[i + 1]), (x[j + 1], y[i])]
            for k in range(4):   # vf(k): fc(:,1) = corner k, fc(:,2) = corner k+1; then l0 lc rVc0 rVc age ageAzimuthal
                a, b = c[k], c[(k + 1) % 4]
                w[j, i, 12 * k:12 * k + 6] = [a[0], a[1], -0.3, b[0], b[1], -0.3]
                w[j, i, 12 * k + 8:12 * k + 10] = 0.01
            w[j, i, 48] = rng.uniform(0.5, 1.0)
            w[j, i, KCP:KCP + 3] = [0.5 * (x[j] + x[j + 1]), 0.5 * (y[i] + y[i + 1]), -0.3]
            w[j, i, KNCAP:KNCAP + 3] = [0.0, 0.0, 1.0]
    return w


def setup(ctx, rotors, wing, nc, ns):
    nr = len(rotors)
    for ir, r in enumerate(rotors):
        ctx.rotor_define(ir, r["nb"], 1, r["ns"], r["nNwake"], r["nFwake"], 2)
        ctx.rotor_set_wake_params(ir, r["nb"], 0, 0, 0, 1, r["ns"], 1.0, 1.0, 0.0, 0.0)
        ctx.rotor_set_rows(ir, 1, 1)
        for ib in range(r["nb"]):
            ctx.rotor_put_nwake(ir, ib, r["waN"][ib])
            if r["nFwake"]:
                ctx.rotor_put_fwake(ir, ib, r["waF"][ib])
    ctx.rotor_define(nr, 1, nc, ns, 0, 0, 0)      # the lifting surface: wing records only
    ctx.rotor_set_wake_params(nr, 1, 0, 0, 0, 1, ns, 1.0, 1.0, 0.0, 0.0)
    ctx.rotor_put_wing(nr, 0, wing)
    return nr


def timed(ctx, ir, m, reps=5):
    ctx.rotor_calc_RHS(ir, m, m)   # warm-up: packs, allocations
    ctx.sync()
    t = []
    for _ in range(reps):
        ctx.rotor_reset_velCP(ir)
        ctx.sync()
        t0 = time.perf_counter()
        v, rhs = ctx.rotor_calc_RHS(ir, m, m)
        t.append(time.perf_counter() - t0)
    return min(t), rhs


def main():
    nfil = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nc, ns = 16, 64
    m = nc * ns
    rotors = synth.rotors_from_lattices(synth.multirotor(nfil, seed=12345))
    wing = wing_records(nc, ns, np.random.default_rng(1))
    out = {}
    for name, devices, share in [("1 GPU", None, None), (f"{gpus} GPUs, sweeps replicated", list(range(gpus)), "-1"),
                                 (f"{gpus} GPUs, source splits shared", list(range(gpus)), None)]:
        if share is None:
            os.environ.pop("VLC_RHS_SHARE_MIN_PAIRS", None)
        else:
            os.environ["VLC_RHS_SHARE_MIN_PAIRS"] = share
        ctx = vb.Context(devices=devices) if devices else vb.Context(0)
        ir = setup(ctx, rotors, wing, nc, ns)
        dt, rhs = timed(ctx, ir, m)
        out[name] = rhs
        print(f"{name:36s}: vlc_rotor_calc_RHS of {m} collocation points x {sum(r['nb'] * r['ns'] * r['nNwake'] for r in rotors) * 4} "
              f"wake filaments: {dt * 1e3:8.3f} ms", flush=True)
        ctx.close()
    ref = out["1 GPU"]
    for k, v in out.items():
        print(f"{k:36s}: RHS bit-identical to one GPU: {np.array_equal(v, ref)}")


if __name__ == "__main__":
    main()
