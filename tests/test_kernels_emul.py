"""The product's O(N) record kernels run on the CPU: tests/native/kernels_emul.cpp compiles volcanor_b200/csrc/
wake_records.cuh (with pack.cuh, pfwake.cuh, vlc_device.cuh) with g++ against a stand-in for <cuda_runtime.h> and runs each
kernel body for every thread of the launch shape capi.cu uses, serially and in reverse thread order.  Held against the
oracle BIT FOR BIT.  Two purposes: kernels that have run on a B200 (age, dissipate, strain) show that the emulation tells
the truth; kernels written after the round's GPU minutes were spent (burst, skew, the prescribed far wake, the linear
combinations of fdScheme 4 / 5) get their index logic and arithmetic checked before their first GPU run."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
i32, f64, vp = C.c_int, C.c_double, C.c_void_p


def emul_lib():
    here = Path(__file__).resolve().parent / "native"
    so = here / "libkernels_emul.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(here)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    sig = {"emul_age_wake": (None, [i32] * 6 + [f64, f64, vp, vp]),
           "emul_dissipate_wake": (None, [i32] * 6 + [f64] * 4 + [vp, vp]),
           "emul_strain_wake": (None, [i32] * 3 + [vp]),
           "emul_burst_wake": (None, [i32] * 3 + [f64, f64, vp]),
           "emul_calc_skew": (None, [i32] * 6 + [vp]),
           "emul_updatePrescribedWake": (i32, [i32] * 6 + [f64] + [vp] * 6),
           "emul_lincomb": (None, [C.c_longlong, i32, vp, vp, vp, vp, vp, f64, vp]),
           "emul_assignshed": (None, [i32] * 6 + [vp, vp]),
           "emul_convectwake": (None, [i32] * 10 + [f64] + [vp] * 7),
           "emul_rollup": (None, [i32] * 9 + [vp, vp]),
           "emul_pack_wing_subset": (None, [i32] * 4 + [vp, vp]),
           "emul_pack_pfwake": (None, [i32, vp, vp]),
           "emul_vind_records": (None, [C.c_longlong, vp, C.c_longlong, vp, vp]),
           "emul_flat_sweep": (i32, [i32, i32, C.c_longlong] + [vp] * 5 + [C.c_longlong, vp, vp]),
           "emul_rsqrt": (None, [i32, C.c_longlong, vp, vp]),
           "emul_lattice_vind": (i32, [i32, i32, vp, i32, i32, i32, i32, C.c_longlong, vp, vp]),
           "emul_lattice_vind_plan": (i32, [i32, i32, i32, vp, i32, i32, i32, i32, C.c_longlong, vp, vp]),
           "emul_lattice_vind_split": (i32, [i32, i32, i32, i32, vp, i32, i32, i32, i32, C.c_longlong, vp, vp]),
           "emul_lattice_vind_dispatch": (i32, [i32, vp, i32, i32, i32, i32, C.c_longlong, vp, vp, vp])}
    for k, (res, args) in sig.items():
        getattr(lib, k).restype = res
        getattr(lib, k).argtypes = args
    return lib


def _case(oracle, nsteps, **geom):
    """A small three-blade hovering rotor (tutorials/caradonna.case with a coarse lattice): near wake of 6 rows, far wake of
    4 rows full after 10 steps, axisymmetric unless asked otherwise."""
    fx = json.loads((GOLDEN / "caradonna.json").read_text())
    fx["config"]["nt"] = 40
    g = fx["geom"][0]
    g.update(nb=3, nc=3, ns=6, nNwake=6, wakeTruncateNt=10, axisymmetrySwitch=1)
    g.update(geom)
    c = oracle.Case(fx)
    c.init()
    for _ in range(nsteps):
        c.step()
    return c, fx


def _stack(rot, what, *a):
    return np.ascontiguousarray(np.stack([getattr(rot, what)(ib, *a) for ib in range(rot.nb)]))


@pytest.mark.parametrize("nsteps", [4, 9])          # wake still growing (rowNear = 3, no far wake yet) / rolled up, far wake
def test_emulation_agrees_with_the_oracle_on_kernels_that_ran_on_a_b200(oracle, nsteps):
    case, fx = _case(oracle, nsteps)
    rot, lib, olib = case.rotor(0), emul_lib(), oracle.load()
    d, p = rot.dims(), rot.params()
    dt, nu = 0.0031, fx["config"]["kinematicVisc"]
    dims = (rot.nb, rot.ns, rot.nNwake, rot.nFwake, d["rowNear"], d["rowFar"])
    waN, waF = _stack(rot, "waN"), _stack(rot, "waF")
    lib.emul_age_wake(*dims, dt, p["omegaSlow"], waN.ctypes.data, waF.ctypes.data)
    olib.orc_rotor_age_wake(rot.h, dt)
    assert np.array_equal(waN, _stack(rot, "waN")) and np.array_equal(waF, _stack(rot, "waF"))
    lib.emul_dissipate_wake(*dims, p["apparentViscCoeff"], p["decayCoeff"], dt, nu, waN.ctypes.data, waF.ctypes.data)
    olib.orc_rotor_dissipate_wake(rot.h, dt, nu)
    assert np.array_equal(waN, _stack(rot, "waN")) and np.array_equal(waF, _stack(rot, "waF"))
    if rot.nFwake and d["rowFar"] <= rot.nFwake:
        before = waF.copy()
        lib.emul_strain_wake(rot.nb, rot.nFwake, d["rowFar"], waF.ctypes.data)
        olib.orc_rotor_strain_wake(rot.h)
        assert np.array_equal(waF, _stack(rot, "waF")) and not np.array_equal(waF, before)


@pytest.mark.parametrize("nsteps,axisym", [(4, 1), (7, 1), (12, 1), (12, 0)])
def test_emulated_convect_rollup_assignshed_agree_with_the_oracle(oracle, nsteps, axisym):
    """rec_convect / rec_continuity / rec_axisym, rec_shiftFwake / rec_rollup / rec_shiftwake, rec_assignshed (all green on a
    B200 through the C ABI) in the launch sequences of vlc_rotor_convectwake / _rollup / _assignshed: 'C' and 'P' (the
    predictor's loop quirk), growing wake, rolled-up wake with free far rows, full far wake (shiftFwake)."""
    case, fx = _case(oracle, nsteps, axisymmetrySwitch=axisym)
    rot, lib, olib = case.rotor(0), emul_lib(), oracle.load()
    d, p = rot.dims(), rot.params()
    T, rotate = _rotations(oracle, rot.nb, p["shaftAxis"])
    hub = np.ascontiguousarray(p["hubCoords"])
    dt = abs(case.dt) if hasattr(case, "dt") else 0.0007
    velN, velF = _stack(rot, "vel", 0), _stack(rot, "vel", 4)
    assert np.any(velN != 0)
    for pred in (True, False):
        waN, waF = _stack(rot, "waN", pred), _stack(rot, "waF", pred)
        lib.emul_convectwake(int(pred), rot.nb, p["nbConvect"], axisym, p["ductSwitch"], rot.ns, rot.nNwake, rot.nFwake,
                             d["rowNear"], d["rowFar"], dt, hub.ctypes.data, T.ctypes.data, rotate.ctypes.data, velN.ctypes.data,
                             velF.ctypes.data, waN.ctypes.data, waF.ctypes.data)
        before = _stack(rot, "waN", pred)
        olib.orc_rotor_convectwake(rot.h, 0, dt, b"P" if pred else b"C")
        assert np.array_equal(waN, _stack(rot, "waN", pred)) and np.array_equal(waF, _stack(rot, "waF", pred)), pred
        assert pred or not np.array_equal(waN, before)      # the predictor's near-wake loop only visits row 1 (quirk C1)
    wiP = _stack(rot, "wiP")
    for edge in (b"TE", b"LE"):
        waN = _stack(rot, "waN")
        lib.emul_assignshed(int(edge == b"TE"), rot.nb, rot.nc, rot.ns, rot.nNwake, d["rowNear"], wiP.ctypes.data, waN.ctypes.data)
        olib.orc_rotor_assignshed(rot.h, edge)
        assert np.array_equal(waN, _stack(rot, "waN")), edge
    if d["rowNear"] == 1:
        waN, waF = _stack(rot, "waN"), _stack(rot, "waF")
        sgn = int(np.copysign(1.0, p["Omega"] * p["theta0"]) > np.finfo(float).eps)
        lib.emul_rollup(rot.nb, rot.ns, rot.nNwake, rot.nFwake, d["rowFar"], p["rollupStart"], p["rollupEnd"], sgn,
                        p["suppressFwakeSwitch"], waN.ctypes.data, waF.ctypes.data)
        olib.orc_rotor_rollup(rot.h)
        assert np.array_equal(waN, _stack(rot, "waN")) and np.array_equal(waF, _stack(rot, "waF"))


@pytest.mark.parametrize("axisym", [1, 0])
def test_skew_kernel(oracle, axisym):
    case, _ = _case(oracle, 4, axisymmetrySwitch=axisym)
    rot, lib = case.rotor(0), emul_lib()
    d, p = rot.dims(), rot.params()
    assert d["rowNear"] == 3
    waN = _stack(rot, "waN")
    lib.emul_calc_skew(rot.nb, p["nbConvect"], axisym, rot.ns, rot.nNwake, d["rowNear"], waN.ctypes.data)
    oracle.load().orc_rotor_calc_skew(rot.h)
    ref = _stack(rot, "waN")
    assert np.array_equal(waN, ref)
    assert np.any(ref[:, :, 2:, 49] > 0) and np.all(ref[:, :, :2, 49] == 0)


def test_burst_kernel(oracle):
    case, _ = _case(oracle, 15, wakeTruncateNt=14)
    rot, lib, olib = case.rotor(0), emul_lib(), oracle.load()
    d = rot.dims()
    olib.orc_burst_pair.restype = C.c_int
    olib.orc_burst_pair.argtypes = [vp, vp, f64]
    waF = _stack(rot, "waF")
    row0 = d["rowFar"] - 1
    skews = []
    for ib in range(rot.nb):
        seg = waF[ib, row0:, 3:6] - waF[ib, row0:, 0:3]
        cosang = np.einsum("ij,ij->i", seg[:-1], -seg[1:]) / (np.linalg.norm(seg[:-1], axis=1) * np.linalg.norm(seg[1:], axis=1))
        skews += list(np.abs(np.arccos(np.clip(cosang, -1, 1)) - np.pi) / np.pi)
    sk = np.sort(skews)
    k = int(np.nonzero(np.diff(sk) > 1e-6 * sk[-1])[0][-1])       # the last clear gap: only the sharpest kinks burst
    limit = 0.5 * (sk[k] + sk[k + 1])
    ref = waF.copy()
    for ib in range(rot.nb):
        for i in range(row0, rot.nFwake - 1):
            if olib.orc_burst_pair(waF[ib, i].ctypes.data, waF[ib, i + 1].ctypes.data, limit):
                ref[ib, i, 9] = ref[ib, i + 1, 9] = 0.77
    got = waF.copy()
    lib.emul_burst_wake(rot.nb, rot.nFwake, d["rowFar"], limit, 0.77, got.ctypes.data)
    assert np.array_equal(got, ref)
    n_burst = int(np.sum(ref[:, :, 9] != waF[:, :, 9]))
    assert 2 <= n_burst < rot.nb * rot.nFwake
    # rowFar = nFwake (one active filament) and rowFar > nFwake (none): nothing to do, nothing touched
    for rowFar in (rot.nFwake, rot.nFwake + 1):
        got = waF.copy()
        lib.emul_burst_wake(rot.nb, rot.nFwake, rowFar, 0.0, 0.77, got.ctypes.data)
        assert np.array_equal(got, waF)


def _rotations(oracle, nb, axis):
    olib = oracle.load()
    olib.orc_getTransformAxis.argtypes = [f64, vp, vp]
    T, rotate = np.zeros((nb, 9)), np.zeros(nb, dtype=np.int32)
    two_pi = 2.0 * (np.arctan(1.0) * 4.0)
    for ib in range(1, nb):
        off = two_pi / nb * ib
        rotate[ib] = abs(off) > np.finfo(float).eps
        olib.orc_getTransformAxis(off, np.ascontiguousarray(axis).ctypes.data, T[ib].ctypes.data)
    return T, rotate


@pytest.mark.parametrize("gen,axisym", [(0, 1), (2, 1), (3, 0)])
def test_prescribed_wake_kernels(oracle, gen, axisym):
    """pf_fit_kernel + pf_helix_kernel against rotor%updatePrescribedWake of the oracle over three successive updates of
    both record sets: records of every blade and the fit parameters bit-identical (libm's cos / sin on both sides)."""
    case, _ = _case(oracle, 15, prescWakeAfterTruncNt=2, prescWakeGenNt=gen, axisymmetrySwitch=axisym)
    rot, lib, olib = case.rotor(0), emul_lib(), oracle.load()
    d, p = rot.dims(), rot.params()
    T, rotate = _rotations(oracle, rot.nb, p["shaftAxis"])
    hub = np.ascontiguousarray(p["hubCoords"])
    for pred in (False, True):
        helix = np.zeros((rot.nb, 2))
        for ib in range(rot.nb):
            olib.orc_rotor_get_pfHelix(rot.h, ib, int(pred), helix[ib].ctypes.data)
        wapF = _stack(rot, "wapF", pred)
        assert np.all(np.abs(wapF[:, :, 12]) > 0)
        for rep in range(3):
            dt = 0.0137 * (rep + 1)
            waF = _stack(rot, "waF", pred)
            rc = lib.emul_updatePrescribedWake(rot.nb, p["nbConvect"], axisym, rot.nFwake, d["rowFar"], gen, p["omegaSlow"] * dt,
                                               hub.ctypes.data, T.ctypes.data, rotate.ctypes.data, waF.ctypes.data,
                                               wapF.ctypes.data, helix.ctypes.data)
            assert rc == 0
            assert olib.orc_rotor_updatePrescribedWake(rot.h, dt, b"P" if pred else b"C") == 0
            assert np.array_equal(wapF, _stack(rot, "wapF", pred)), (pred, rep)
            for ib in range(rot.nb):
                ref = np.zeros(2)
                olib.orc_rotor_get_pfHelix(rot.h, ib, int(pred), ref.ctypes.data)
                assert np.array_equal(helix[ib], ref), (pred, rep, ib)


@pytest.mark.parametrize("nterms,coef,div", [(3, [23.0, -16.0, 5.0], 12.0), (4, [55.0, -59.0, 37.0, -9.0], 24.0),
                                             (3, [5.0, 8.0, -1.0], 12.0), (4, [9.0, 19.0, -5.0, 1.0], 24.0), (1, [1.0], 1.0)])
def test_lincomb_kernel(nterms, coef, div):
    """rec_lincomb_kernel = the multistep formulas of fdScheme 4 / 5 (main.f90:1160-1172, :1222-1231, :1309-1325, :1370-1381),
    left to right, unfused, dst allowed to be the first source; n not a multiple of the block size."""
    rng = np.random.default_rng(nterms)
    n = 3 * 7 * 17 * 2 + 1
    src = [rng.standard_normal(n) for _ in range(4)]
    want = coef[0] * src[0]
    for k in range(1, nterms):
        want = want + coef[k] * src[k]
    want = want / div
    dst = src[0]                                       # in place, like vel = (23*vel - 16*vel2 + 5*vel1)/12
    guard = src[1].copy()
    c = np.array(coef + [0.0] * (4 - nterms))
    emul_lib().emul_lincomb(n, nterms, src[0].ctypes.data, src[1].ctypes.data, src[2].ctypes.data, src[3].ctypes.data,
                            c.ctypes.data, div, dst.ctypes.data)
    assert np.array_equal(dst, want) and np.array_equal(src[1], guard)


def _vind(lib, rec, P):
    rec, P = np.ascontiguousarray(rec), np.ascontiguousarray(P)
    V = np.empty_like(P)
    lib.emul_vind_records(rec.size // 12, rec.ctypes.data, P.shape[0], P.ctypes.data, V.ctypes.data)
    return V


def test_pack_kernels_and_pair_arithmetic_bound_chordwise_prescribed(oracle):
    """pack_rings_kernel with the filament masks of pack_bound (vf2 + vf4 - TE; green on a B200) and pack_chord (vf1 + vf3 +
    TE; not yet run on a GPU), pack_fwake_kernel on the 240 prescribed filaments per blade, evaluated with the sweeps' pair
    arithmetic (pair_accumulate, exact seed): against the oracle's source loops (classdef.f90:1376-1418, :1471-1476) at the
    per-call bar 1e-12 of the velocity scale; bound + chordwise = the whole wing."""
    case, _ = _case(oracle, 16, prescWakeAfterTruncNt=2)
    rot, lib = case.rotor(0), emul_lib()
    wiP = _stack(rot, "wiP")
    rng = np.random.default_rng(2)
    span = float(np.max(np.abs(wiP[:, :, :, 0:3])))
    P = np.concatenate([rng.uniform(-1.2, 1.2, (60, 3)) * span, wiP[0, ::3, 0, 64:67] + [0.0, 0.0, 0.02],
                        wiP[2, ::5, 1, 0:3]])                   # random, just above collocation points, on ring corners
    per_blade = 2 * rot.nc * rot.ns + rot.ns
    got = {}
    for what in (3, 4):
        rec = np.zeros((rot.nb * per_blade, 12))
        lib.emul_pack_wing_subset(what, rot.nb, rot.nc, rot.ns, wiP.ctypes.data, rec.ctypes.data)
        assert np.all(np.any(rec[:, 0:6] != 0, axis=1))           # every record written
        got[what] = _vind(lib, rec, P)
        ref = rot.vind_points(what, P)
        scale = 50.0 * float(np.abs(ref).max())
        assert np.max(np.abs(got[what] - ref)) < 1e-12 * scale, (what, float(np.max(np.abs(got[what] - ref))) / scale)
    whole = rot.vind_points(0, P)
    assert np.max(np.abs(got[3] + got[4] - whole)) < 1e-12 * 50.0 * float(np.abs(whole).max())
    # the prescribed helix alone: vind_bywake with and without it differ by exactly its 240 x nb filaments
    wapF = _stack(rot, "wapF")
    assert np.all(np.abs(wapF[:, :, 12]) > 0)
    rec = np.zeros((rot.nb * 240, 12))
    lib.emul_pack_pfwake(rot.nb, wapF.ctypes.data, rec.ctypes.data)
    Pw = np.concatenate([rng.uniform(-1.5, 1.5, (60, 3)) * float(np.max(np.abs(wapF[:, :, 0:2]))), wapF[0, ::11, 0:3]])
    helix = _vind(lib, rec, Pw)
    with_helix = rot.vind_points(1, Pw)
    for ib in range(rot.nb):
        rot.wapF(ib)[:, 12] = 0.0                                 # abs(gam) > eps rule: the helix drops out
    without = rot.vind_points(1, Pw)
    scale = 50.0 * float(np.abs(with_helix).max())
    assert np.max(np.abs(helix - (with_helix - without))) < 1e-12 * scale
    assert np.max(np.abs(helix)) > 1e-4 * float(np.abs(with_helix).max())


@pytest.mark.parametrize("W,T,ns,nsteps", [(1, 1, 6, 4), (2, 2, 6, 4), (3, 1, 6, 8), (3, 2, 6, 4), (4, 1, 8, 4), (4, 2, 8, 8), (1, 3, 5, 4)])
def test_lattice_kernel_on_the_cpu(oracle, W, T, ns, nsteps):
    """The dominant kernel itself: pack_rings_shared_kernel<W> -> bs_lattice_kernel<W, T> (node-by-node, edge-by-edge
    regrouping with merged strengths, strip walk with the previous row in registers, tile staging, the c2 > eps^2 guard) +
    the flat remainder, per blade, on the near wake of a hovering rotor: against rotor%vind_bywake of the oracle at the
    per-call bar.  Targets: random points, the wake's own nodes (every adjacent edge must drop out exactly) and points on
    edges.  Wake still growing (rows 1, 2 inactive) and fully shed.  The reciprocal-square-root seed is exact here instead of
    MUFU.RSQ64H's 20 bits; its refinement and everything else is the product's source."""
    case, _ = _case(oracle, nsteps, ns=ns, wakeTruncateNt=0, nNwake=8)       # no far wake within 8 steps
    rot, lib = case.rotor(0), emul_lib()
    d = rot.dims()
    assert d["rowFar"] > rot.nFwake and d["rowNear"] == max(1, 9 - nsteps)
    nrows, i0 = rot.nNwake - d["rowNear"] + 1, d["rowNear"] - 1
    # end of a step with every row shed: the current records are mid-update (after shiftwake the newest row's vf(4)%rVc waits
    # for the next dissipate_wake, classdef.f90:4386-4392) and NOT mergeable -- pack_rings_shared_kernel must say so (the
    # product then sweeps the flat enumeration); the predicted records, which the last corrector sweep used, are a lattice
    pred = d["rowNear"] == 1
    if pred:
        V = np.empty((1, 3))
        assert lib.emul_lattice_vind(W, T, np.ascontiguousarray(_stack(rot, "waN")[0]).ctypes.data, rot.nNwake, rot.ns, i0, nrows,
                                     1, np.zeros((1, 3)).ctypes.data, V.ctypes.data) == 1
    waN = _stack(rot, "waN", pred)
    rng = np.random.default_rng(W * 10 + T)
    nodes = waN[0, :, i0:, 12:15].reshape(-1, 3)                                # corner 2 of every active ring of blade 1
    mid = 0.5 * (waN[1, :, i0:, 12:15] + waN[1, :, i0:, 24:27]).reshape(-1, 3)  # on the TE edges of blade 2
    P = np.ascontiguousarray(np.concatenate([rng.uniform(-1.3, 1.3, (300, 3)) * float(np.abs(nodes).max()), nodes, mid[::3]]))
    got = np.zeros_like(P)
    for ib in range(rot.nb):
        V = np.empty_like(P)
        rc = lib.emul_lattice_vind(W, T, np.ascontiguousarray(waN[ib]).ctypes.data, rot.nNwake, rot.ns, i0, nrows, P.shape[0],
                                   P.ctypes.data, V.ctypes.data)
        assert rc == 0
        got = got + V
    ref = rot.vind_points(1, P, pred)
    assert np.all(np.isfinite(got))
    scale = 50.0 * float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref)))
    print(f"lattice kernel W={W} T={T} on the CPU: {P.shape[0]} targets x {rot.nb * nrows * rot.ns} rings, max error "
          f"{err / scale * 50:.2e} of the velocity scale")
    assert err < 1e-12 * scale, err / scale


@pytest.mark.parametrize("W,T,tailW,ns", [(4, 2, 0, 6), (4, 2, 2, 6), (4, 2, 1, 5), (4, 1, 3, 7), (4, 2, 1, 9), (3, 2, 0, 5),
                                          (2, 2, 0, 5), (4, 2, 0, 3), (4, 2, 0, 2), (1, 3, 0, 2)])
def test_lattice_kernel_on_the_cpu_strip_plans(oracle, W, T, tailW, ns):
    """Column counts that are not a multiple of the strip width, covered the two ways capi.cu's plan_strips does it: a
    partial last strip of the same width (fill_strip_record clips at ns), or floor(ns / 4) strips of width 4 and one tail
    strip of width ns mod 4 packed with col_base = 4 floor(ns / 4) and swept by a second lattice launch (green on a B200:
    tests/test_gpu_lattice.py::test_tail_strips_cover_column_counts_that_are_not_multiples_of_four, ::test_degenerate_
    lattice_shapes).  Same targets, oracle and bar as test_lattice_kernel_on_the_cpu; growing wake (rows 1..4 inactive)."""
    case, _ = _case(oracle, 4, ns=ns, wakeTruncateNt=0, nNwake=8)
    rot, lib = case.rotor(0), emul_lib()
    d = rot.dims()
    nrows, i0 = rot.nNwake - d["rowNear"] + 1, d["rowNear"] - 1
    assert nrows == 4 and d["rowFar"] > rot.nFwake
    waN = _stack(rot, "waN")
    rng = np.random.default_rng(100 * W + 10 * tailW + ns)
    nodes = waN[0, :, i0:, 12:15].reshape(-1, 3)
    mid = 0.5 * (waN[1, :, i0:, 12:15] + waN[1, :, i0:, 24:27]).reshape(-1, 3)
    last = waN[2, ns - 1, i0:, 24:27].reshape(-1, 3)                              # corner 3 of the last column (the flat remainder's edge)
    P = np.ascontiguousarray(np.concatenate([rng.uniform(-1.3, 1.3, (200, 3)) * float(np.abs(nodes).max()), nodes, mid, last]))
    got = np.zeros_like(P)
    for ib in range(rot.nb):
        V = np.empty_like(P)
        rc = lib.emul_lattice_vind_plan(W, T, tailW, np.ascontiguousarray(waN[ib]).ctypes.data, rot.nNwake, rot.ns, i0, nrows,
                                        P.shape[0], P.ctypes.data, V.ctypes.data)
        assert rc == 0, rc
        got = got + V
    ref = rot.vind_points(1, P, False)
    assert np.all(np.isfinite(got)) and np.abs(ref).max() > 0
    scale = 50.0 * float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref)))
    assert err < 1e-12 * scale, err / scale


@pytest.mark.parametrize("W,tailW,ns,nsplit", [(4, 0, 8, 1), (4, 0, 8, 3), (4, 2, 6, 2), (4, 1, 5, 1), (2, 0, 6, 2), (1, 0, 5, 1),
                                               (3, 0, 6, 1)])
def test_dual_core_streamwise_edges_on_the_cpu(oracle, W, tailW, ns, nsplit):
    """A wake shed with a NON-UNIFORM streamwiseCoreVec (classdef.f90:3841: vf(1) of ring (i, j) gets streamwiseCoreVec(j),
    vf(3) gets streamwiseCoreVec(j+1); rotor_dissipate_wake then copies vf(1)%rVc into vf(3)%rVc, :4371-4372):
    the two copies of every interior streamwise edge carry different core radii.  check_rings_kernel classifies the set as
    DUAL (flag 2), pack_rings_shared_kernel writes the dual records, the dual form of bs_lattice_kernel (33-instruction
    edge: one cross product, two reciprocal square roots) runs and the merged form does not; the result is the reference's
    ring-by-ring sum at the per-call bar.  Round 1 sent such a set to the flat enumeration whole (2.6x slower)."""
    case, _ = _case(oracle, 6, ns=ns, wakeTruncateNt=0, nNwake=8)
    rot, lib = case.rotor(0), emul_lib()
    d = rot.dims()
    nrows, i0 = rot.nNwake - d["rowNear"] + 1, d["rowNear"] - 1
    assert nrows >= 4 and d["rowFar"] > rot.nFwake
    # what the reference's records look like after shedding with streamwiseCoreVec(j) = base*(1 + 0.3 j) and one dissipation
    # step: vf(1) and vf(3) of ring (i, j) hold the SAME radius (the left edge's), so ring j's vf(3) differs from ring j+1's vf(1)
    for ib in range(rot.nb):
        w = rot.waN(ib)
        for j in range(ns):
            w[j, :, 9] *= 1.0 + 0.3 * j
            w[j, :, 24 + 9] = w[j, :, 9]
    waN = _stack(rot, "waN")
    rng = np.random.default_rng(7 * W + ns)
    nodes = waN[0, :, i0:, 12:15].reshape(-1, 3)
    mid = 0.5 * (waN[1, :, i0:, 12:15] + waN[1, :, i0:, 0:3]).reshape(-1, 3)       # on the streamwise (dual) edges of blade 2
    P = np.ascontiguousarray(np.concatenate([rng.uniform(-1.3, 1.3, (250, 3)) * float(np.abs(nodes).max()), nodes, mid]))
    got = np.zeros_like(P)
    for ib in range(rot.nb):
        V = np.empty_like(P)
        rc = lib.emul_lattice_vind_split(W, {1: 1, 2: 2, 3: 1, 4: 1}[W], tailW, nsplit, np.ascontiguousarray(waN[ib]).ctypes.data,
                                         rot.nNwake, rot.ns, i0, nrows, P.shape[0], P.ctypes.data, V.ctypes.data)
        assert rc == -2, rc                                                         # done, by the dual form
        got = got + V
    ref = rot.vind_points(1, P, False)
    assert np.all(np.isfinite(got))
    scale = 50.0 * float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref)))
    print(f"dual form W={W} tail={tailW} on the CPU: max error {err / scale * 50:.2e} of the velocity scale")
    assert err < 1e-12 * scale, err / scale
    # and the merged form on the same records would be WRONG by far more than the bar (the test has teeth)
    for ib in range(rot.nb):
        w = rot.waN(ib)
        w[:, :, 24 + 9] = np.roll(w[:, :, 9], -1, axis=0)                           # make the copies agree: a different wake
    assert np.max(np.abs(rot.vind_points(1, P, False) - ref)) > 1e-8 * scale
    # device-side dispatch with every launch made: flag 2 -> the dual slots are the ones summed
    for ib in range(rot.nb):
        w = rot.waN(ib)
        w[:, :, 24 + 9] = w[:, :, 9]
    if W == 4 and tailW == 0:
        waN = _stack(rot, "waN")
        got = np.zeros_like(P)
        for ib in range(rot.nb):
            V, flag = np.empty_like(P), C.c_int(-1)
            assert lib.emul_lattice_vind_dispatch(2, np.ascontiguousarray(waN[ib]).ctypes.data, rot.nNwake, rot.ns, i0, nrows, P.shape[0],
                                                  P.ctypes.data, V.ctypes.data, C.byref(flag)) == 0
            assert flag.value == 2
            got = got + V
        assert np.all(np.isfinite(got)) and np.max(np.abs(got - ref)) < 1e-12 * scale


@pytest.mark.parametrize("W,T,tailW,nsplit,ns", [(1, 1, 0, 1, 8), (1, 3, 0, 2, 8), (1, 3, 0, 4, 8), (2, 2, 0, 3, 8), (4, 2, 0, 1, 8),
                                                 (4, 2, 0, 2, 8), (4, 1, 0, 3, 8), (3, 2, 0, 2, 6), (4, 2, 2, 2, 6),
                                                 (1, 3, 0, 3, 8), (2, 2, 0, 5, 8), (4, 2, 0, 7, 8)])
def test_lattice_kernel_on_the_cpu_source_splits_and_tile_ring(oracle, W, T, tailW, nsplit, ns):
    """Long near wake (30 active rows): the strip records fill several shared-memory tiles, so one CTA walks more tiles than
    the ring has stages (3) and the sweep is cut into source splits of whole granules (grid y, sweep_shared's chunks: the last
    three cases end a chunk in a partial tile -- 1.5 tiles, half a tile, half a tile); then
    check_rings_kernel and bs_reduce_select_kernel close the launch sequence of the mergeable path.  Oracle and bar as above;
    more than 128 T targets, so several CTAs in x as well."""
    case, _ = _case(oracle, 30, ns=ns, wakeTruncateNt=0, nNwake=32)
    rot, lib = case.rotor(0), emul_lib()
    d = rot.dims()
    nrows, i0 = rot.nNwake - d["rowNear"] + 1, d["rowNear"] - 1
    assert nrows == 30 and d["rowFar"] > rot.nFwake
    waN = _stack(rot, "waN")
    rng = np.random.default_rng(1000 * W + 100 * T + nsplit)
    nodes = waN[0, :, i0:, 12:15].reshape(-1, 3)
    P = np.ascontiguousarray(np.concatenate([rng.uniform(-1.3, 1.3, (190, 3)) * float(np.abs(nodes).max()), nodes]))
    assert P.shape[0] > 128 * T
    got = np.zeros_like(P)
    for ib in range(rot.nb):
        V = np.empty_like(P)
        rc = lib.emul_lattice_vind_split(W, T, tailW, nsplit, np.ascontiguousarray(waN[ib]).ctypes.data, rot.nNwake, rot.ns, i0,
                                         nrows, P.shape[0], P.ctypes.data, V.ctypes.data)
        assert rc == 0, rc
        got = got + V
    ref = rot.vind_points(1, P, False)
    assert np.all(np.isfinite(got)) and np.abs(ref).max() > 0
    scale = 50.0 * float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref)))
    assert err < 1e-12 * scale, err / scale


@pytest.mark.parametrize("predicted,nsplit_flat", [(False, 1), (False, 3), (True, 2)])
def test_device_side_dispatch_between_lattice_and_flat_enumeration(oracle, predicted, nsplit_flat):
    """sweep_shared never reads the mergeability flag on the host: the lattice launch and the flat remainder run when it is 0,
    the flat enumeration when it is 1, and bs_reduce_select_kernel sums the slots of the path that ran.  End of a step with
    every row shed: the CURRENT records are mid-update (newest row's vf(4)%rVc, classdef.f90:4386-4392) -> flag 1, flat path;
    the PREDICTED records are a lattice -> flag 0.  Both against rotor%vind_bywake of the oracle, partial buffer pre-filled
    with NaN.  (On a B200: tests/test_gpu_lattice.py::test_unmergeable_core_radii_fall_back_to_flat_enumeration.)"""
    case, _ = _case(oracle, 8, ns=6, wakeTruncateNt=0, nNwake=8)
    rot, lib = case.rotor(0), emul_lib()
    d = rot.dims()
    assert d["rowNear"] == 1 and d["rowFar"] > rot.nFwake
    waN = _stack(rot, "waN", predicted)
    rng = np.random.default_rng(7 + nsplit_flat)
    nodes = waN[0, :, :, 12:15].reshape(-1, 3)
    P = np.ascontiguousarray(np.concatenate([rng.uniform(-1.3, 1.3, (530, 3)) * float(np.abs(nodes).max()), nodes]))
    got = np.zeros_like(P)
    for ib in range(rot.nb):
        V, flag = np.empty_like(P), C.c_int(-1)
        rc = lib.emul_lattice_vind_dispatch(nsplit_flat, np.ascontiguousarray(waN[ib]).ctypes.data, rot.nNwake, rot.ns, 0,
                                            rot.nNwake, P.shape[0], P.ctypes.data, V.ctypes.data, C.byref(flag))
        assert rc == 0 and flag.value == (0 if predicted else 1), (rc, flag.value)
        got = got + V
    ref = rot.vind_points(1, P, predicted)
    assert np.all(np.isfinite(got))
    scale = 50.0 * float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref)))
    assert err < 1e-12 * scale, err / scale


def test_rsqrt_refinement_with_a_seed_as_coarse_as_the_devices():
    """rsqrt_fp64 (vlc_device.cuh): the emulation's seed keeps ~21 bits like MUFU.RSQ64H (measured 2^-20.06 on a B200,
    tests/test_gpu_parity.py::test_rsqrt_seed_accuracy); the third-order step must bring it to rounding level, the
    second-order step (opt-in precision mode) to <= 1.5 d^2, always low -- the bounds DESIGN.md section 4 states."""
    rng = np.random.default_rng(0)
    x = np.ascontiguousarray(np.exp(rng.uniform(np.log(1e-60), np.log(1e60), 200000)))
    lib = emul_lib()
    exact = 1.0 / np.sqrt(x.astype(np.longdouble))
    for fast, bound in ((0, 4e-16), (1, 1.5 * 2.0 ** -40 * 1.3)):
        y = np.empty_like(x)
        lib.emul_rsqrt(fast, x.size, x.ctypes.data, y.ctypes.data)
        rel = ((y.astype(np.longdouble) - exact) / exact).astype(np.float64)
        assert np.max(np.abs(rel)) < bound, (fast, float(np.max(np.abs(rel))))
        if fast:
            assert np.max(rel) < 1e-15                               # never high: what the centring factor relies on


@pytest.mark.parametrize("n,m,nsplit,fast", [(1, 1, 1, 0), (7, 3, 1, 0), (127, 129, 1, 0), (129, 513, 2, 0), (1000, 77, 3, 0),
                                             (2000, 520, 4, 0), (1000, 77, 2, 1), (234, 104, 8, 0), (600, 130, 7, 0)])
def test_flat_sweep_kernel_on_the_cpu(oracle, n, m, nsplit, fast):
    """pack_flat_kernel -> bs_sweep_kernel<4, 128, 128, 3> (tile ring, T = 4 targets per thread, source splits) ->
    bs_reduce_kernel on the random sets of tests/test_gpu_parity.py::test_flat_random_vs_oracle (targets on end points and
    on filaments, gam = 0 and |gam| <= eps with and without the wake rule): the same measure and the same bar as on the GPU,
    err = max|V - V_oracle| / max sum|terms| < 1e-12; second-order precision mode included.  Chunks are cut in quarter
    tiles like plan_flat's (the last two cases: 8 chunks of 32 sources; 7 chunks of 96 = three quarters of a tile)."""
    from tests.helpers import scaled_err
    from volcanor_b200 import synth
    p1, p2, rvc, gam, flag, P = synth.random_filaments(n, m, seed=n + m)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (p1, p2, rvc, gam)]
    flag = np.ascontiguousarray(flag, dtype=np.uint8)
    P = np.ascontiguousarray(P)
    V = np.empty_like(P)
    rc = emul_lib().emul_flat_sweep(fast, nsplit, n, *[a.ctypes.data for a in arrs], flag.ctypes.data, m, P.ctypes.data,
                                    V.ctypes.data)
    assert rc == 0 and np.all(np.isfinite(V))
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, P)
    _, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, P)
    e = scaled_err(V, Vo, Vabs)
    assert e < 1e-12, e
    if not fast:
        assert e < 1e-14, e                                       # full precision sits at rounding level
