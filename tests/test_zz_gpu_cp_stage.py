"""The collocation-point stage on the device copies of the wing records (tier 2c of the C ABI),
through the C ABI on the GPU:

 1. vlc_rotor_calc_force alone, given the oracle's records (same velCPTotal): loads, sectional arrays and the updated
    wing records BIT-IDENTICAL to the CPU restatement (secAlpha: atan2, 4 ulp) -- the arithmetic is the same source the
    CPU suite checks with g++ (tests/test_cp_stage_host.py);
 2. vlc_rotor_calc_RHS / _solve_map_gam / _calc_velCPTotal against the oracle's sweeps added in the driver's order
    (tolerance 1e-12 of the velocity scale, the per-call bar of the sweeps), two rotors included;
 3. whole cases with the stage on the device (tests/native/case_gpu_hooks.c: h_cp_rhs_solve / h_cp_forces), wake
    resident: both golden histories of the reference to 7 digits on every row and its golden sectional distributions
    (r01b01ForceDistNNNNN.csv.ref: secCL, secCLu, secLift, secVel, secAlpha ...) from the device's loads, CL/CT and
    circulations against the CPU driver within 1e-8.

Runs last (file name) and on its own library context: the stage sums over every rotor the context knows."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.test_cp_stage_host import NET, SEC1, SEC3, _mut, force_params, pack_sections

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
TOL = 1e-12
TOL_HISTORY = 1e-8


@pytest.fixture()
def cctx():
    import volcanor_b200 as vb
    c = vb.Context(0)
    yield c
    c.close()


def _developed(oracle, fx, nsteps):
    c = oracle.Case(fx)
    c.init()
    for _ in range(nsteps):
        c.step()
    return c


def _define(ctx, rot, ir, wake=True):
    d, p = rot.dims(), rot.params()
    ctx.rotor_define(ir, rot.nb, rot.nc, rot.ns, rot.nNwake, rot.nFwake, 1)
    if hasattr(ctx, "attach_sources"):
        ctx.attach_sources(ir, rot)
    ctx.rotor_set_wake_params(ir, p["nbConvect"], p["axisymmetrySwitch"], p["ductSwitch"], p["suppressFwakeSwitch"],
                              p["rollupStart"], p["rollupEnd"], p["Omega"] * p["theta0"], p["apparentViscCoeff"],
                              p["decayCoeff"], p["initWakeVel"])
    ctx.rotor_set_rows(ir, d["rowNear"], d["rowFar"])
    for ib in range(rot.nb):
        ctx.rotor_put_wing(ir, ib, rot.wiP(ib))
        if wake and rot.nNwake:
            ctx.rotor_put_nwake(ir, ib, rot.waN(ib))
            if rot.nFwake:
                ctx.rotor_put_fwake(ir, ib, rot.waF(ib))


# ------------------------------------------------------------------------------------------- 1. loads, bit-identical

@pytest.mark.parametrize("name,nsteps,mutate", [("katzNplotkin_AR04", 6, None), ("caradonna", 5, _mut(nNwake=8)),
                                                ("elevateTest", 5, _mut(nNwake=6)),
                                                ("simplewing", 4, _mut(spanwiseLiftSwitch=1))])
def test_calc_force_bit_identical_to_the_oracle(cctx, oracle, name, nsteps, mutate):
    check_calc_force(cctx, oracle, name, nsteps, mutate)


def check_calc_force(cctx, oracle, name, nsteps, mutate):
    """Body of the test above; also run on the CPU stand-in of tests/cp_stage_emulation.py (tests/test_cp_stage_host.py)."""
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    fx["config"]["rotorForcePlot"] = 1
    if mutate:
        mutate(fx)
    c = _developed(oracle, fx, nsteps)
    cfg, rot = c.config, c.rotor(0)
    fp = force_params(c, rot)
    nb, nc, ns, nbc = rot.nb, rot.nc, rot.ns, fp["nbConvect"]
    rng = np.random.default_rng(7)
    for ib in range(nbc):                       # unsteady terms that are not zero
        rot.wiP(ib)[:, :, 50] += 1e-3 * rng.standard_normal((ns, nc))
    _define(cctx, rot, 0, wake=False)
    for ib in range(nb):
        cctx.rotor_put_sections(0, ib, pack_sections(rot, ib))
    n0 = cctx.launch_count
    cctx.rotor_calc_force(0, cfg.density, cfg.dt, fp["Omega"], fp["spanwiseLiftSwitch"])
    assert cctx.launch_count > n0
    rot.lib.orc_rotor_calc_secAlpha(rot.h)
    rot.lib.orc_rotor_calc_force(rot.h, cfg.density, cfg.dt)
    for ib in range(nb):
        got = cctx.rotor_get_loads(0, ib, ns)
        for n in NET:
            assert np.array_equal(got[n], rot.sec(ib, n, 3)[0]), (name, ib, n, got[n], rot.sec(ib, n, 3)[0])
        for n in SEC3[(0 if ib < nbc else 1):]:          # the reference does not copy secChordwiseResVel to blades 2..nb
            assert np.array_equal(got[n], rot.sec(ib, n, 3)), (name, ib, n, float(np.max(np.abs(got[n] - rot.sec(ib, n, 3)))))
        for n in SEC1[1:]:
            assert np.array_equal(got[n], rot.sec(ib, n)), (name, ib, n)
        np.testing.assert_allclose(got["secAlpha"], rot.sec(ib, "secAlpha"), rtol=1e-15, atol=1e-18)
        w, wr = cctx.rotor_get_wing(0, ib, nc, ns), rot.wiP(ib)
        assert np.array_equal(w, wr), (name, ib, "wing records", np.argwhere(w != wr)[:6])
    assert np.any(cctx.rotor_get_loads(0, 0, ns)["secLiftUnsteady"] != 0.0)


# ------------------------------------------------------------ 2. velCP, RHS, solve + map_gam, velCPTotal vs the oracle

def _dot(a, b):  # unfused, in the reference's order
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def _two_body():
    from tests.test_oracle_case import two_body_case
    return two_body_case()


@pytest.mark.parametrize("case", ["caradonna", "elevateTest", "two_body"])
def test_rhs_solve_and_velcptotal_vs_oracle(cctx, oracle, case):
    check_rhs_solve_velcptotal(cctx, oracle, case)


def check_rhs_solve_velcptotal(cctx, oracle, case):
    if case == "two_body":
        fx, nsteps = _two_body(), 10
    else:
        fx, nsteps = json.loads((GOLDEN / f"{case}.json").read_text()), 8
        _mut(nNwake=6)(fx)
    c = _developed(oracle, fx, nsteps)
    rots = [c.rotor(ir) for ir in range(c.nr)]
    for ir, rot in enumerate(rots):
        for ib in range(rot.nb):                # the records carry the kinematic part in velCP (= velCPm, main.f90:546-547)
            rot.wiP(ib)[:, :, 76:79] = rot.wiP(ib)[:, :, 82:85]
        _define(cctx, rot, ir)
    for ir, rot in enumerate(rots):
        fp = force_params(c, rot)
        nbc, npb = fp["nbConvect"], rot.nc * rot.ns
        m = nbc * npb
        rec = np.concatenate([rot.wiP(ib).reshape(npb, 104) for ib in range(nbc)])
        P, ncap = rec[:, 64:67].copy(), rec[:, 67:70]
        v, scale = rec[:, 76:79].copy(), np.zeros(m)
        for jr, src in enumerate(rots):         # main.f90:551-560, added in the driver's order
            for what in ([1] if jr == ir else [1, 0]):
                dv = src.vind_points(what, P)
                v = v + dv
                scale = scale + np.abs(dv).max(axis=1)
        rhs_ref = np.zeros(rot.N)
        rhs_ref[:m] = _dot(v, ncap)
        if fp["axisym"]:
            for ib in range(1, rot.nb):
                rhs_ref[ib * npb:(ib + 1) * npb] = rhs_ref[:npb]
        rhs_ref = -1.0 * rhs_ref
        assert rot.calcAIC() == 0
        cctx.rotor_calcAIC(ir, rot.N, want_matrix=False)
        vg, rg = cctx.rotor_calc_RHS(ir, m, rot.N)
        s = 50.0 * max(float(scale.max()), float(np.abs(v).max()))    # ~ sum |terms| of the sweeps (tests/test_gpu_parity.py)
        assert np.max(np.abs(vg - v)) < TOL * s, (case, ir, float(np.max(np.abs(vg - v))), s)
        assert np.max(np.abs(rg - rhs_ref)) < TOL * s, (case, ir)
        assert np.array_equal(rg[:m], -1.0 * _dot(vg, ncap))            # the RHS kernel itself is exact
        g = cctx.rotor_solve_map_gam(ir, rot.N)
        gref = rot.AIC(inverse=True) @ rg
        assert np.max(np.abs(g - gref)) < 1e-8 * np.max(np.abs(gref)), (case, ir)      # LU solve vs explicit inverse
        rot.vec(0)[:] = g
        rot.lib.orc_rotor_map_gam(rot.h)
        for ib in range(rot.nb):
            w = cctx.rotor_get_wing(ir, ib, rot.nc, rot.ns)
            assert np.array_equal(w[:, :, 48], rot.wiP(ib)[:, :, 48]), (case, ir, ib, "map_gam")
            if ib < nbc:
                assert np.array_equal(w[:, :, 76:79].reshape(npb, 3), vg[ib * npb:(ib + 1) * npb])
                rot.wiP(ib)[:, :, 76:79] = w[:, :, 76:79]             # the oracle continues from the device's velCP
    import volcanor_b200 as vb
    with pytest.raises(vb.VlcError, match="before vlc_rotor_calc_RHS"):
        cctx.rotor_solve_map_gam(0, rots[0].N)                           # one solve per right-hand side
    # velCPTotal (main.f90:630-663) with the new circulations on both sides
    for ir, rot in enumerate(rots):
        fp = force_params(c, rot)
        nbc, npb = fp["nbConvect"], rot.nc * rot.ns
        rec = np.concatenate([rot.wiP(ib).reshape(npb, 104) for ib in range(nbc)])
        P, vt = rec[:, 64:67].copy(), rec[:, 76:79].copy()
        scale = np.abs(vt).max()
        for src in rots:
            dv = src.vind_points(3, P)
            vt = vt - dv
            scale = max(scale, float(np.abs(dv).max()))
        dv = rot.vind_points(0, P)
        vt = vt + dv
        scale = max(scale, float(np.abs(dv).max()))
        cctx.rotor_calc_velCPTotal(ir)
        for ib in range(rot.nb):
            w = cctx.rotor_get_wing(ir, ib, rot.nc, rot.ns)[:, :, 79:82].reshape(npb, 3)
            ref = vt[ib * npb:(ib + 1) * npb] if ib < nbc else (vt[:npb] if fp["axisym"] else None)
            if ref is not None:
                assert np.max(np.abs(w - ref)) < TOL * 50.0 * scale, (case, ir, ib, float(np.max(np.abs(w - ref))), scale)


def test_cp_stage_argument_and_state_errors(cctx):
    import volcanor_b200 as vb
    with pytest.raises(vb.VlcError, match="rotor not defined"):
        cctx.rotor_calc_RHS(3, 4, 4)
    cctx.rotor_define(0, 2, 2, 3, 4, 0, 1)
    with pytest.raises(vb.VlcError, match="before vlc_rotor_calcAIC"):
        cctx.rotor_solve_map_gam(0, 12)
    with pytest.raises(vb.VlcError, match="put_sections"):
        cctx.rotor_calc_force(0, 1.0, 0.1, 10.0)
    for ib in range(2):
        cctx.rotor_put_sections(0, ib, np.zeros(10 * 3 + 6))
    with pytest.raises(vb.VlcError, match="dt must be positive"):
        cctx.rotor_calc_force(0, 1.0, 0.0, 10.0)
    assert cctx.lib.vlc_rotor_put_sections(cctx.h, 0, 2, None) == 2      # VLC_ERR_ARG: blade index / null pointer
    assert cctx.lib.vlc_rotor_get_loads(cctx.h, 0, 0, None) == 2
    assert cctx.lib.vlc_rotor_get_wing(cctx.h, 0, 5, None) == 2


# ----------------------------------------------------------------------- 3. whole cases with the stage on the device

def _cp_hooks(case, ctx, resident=True):
    import ctypes as C
    if resident:
        from tests.test_gpu_resident import _resident_hooks
        lib, h = _resident_hooks(case, ctx)
    else:
        from tests.test_gpu_case import _native_hooks
        lib, h = _native_hooks(case, ctx)
    lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
    for f in ("case_hooks_cp_rhs_calls", "case_hooks_cp_force_calls"):
        getattr(lib, f).restype = C.c_long
        getattr(lib, f).argtypes = [C.c_void_p]
    assert lib.case_hooks_enable_cp(h) == 0, ctx.lib.vlc_last_error(ctx.h)
    return lib, h


def _step(c, lib, h, ctx, it):
    try:
        c.step()
    except RuntimeError as e:
        raise AssertionError(f"step {it}: {e}; rc={lib.case_gpu_hooks_last_rc(h)} {ctx.lib.vlc_last_error(ctx.h)}") from e


@pytest.mark.parametrize("name,nsteps", [("katzNplotkin_AR04", 160), ("elevateTest", 150)])
def test_cp_stage_run_reproduces_reference_golden_history(cctx, oracle, name, nsteps):
    """Wake resident AND the collocation-point stage on the device: per step only the moved wing goes up and gamVec, the
    wing records and the loads come back.  Every row of the reference's golden file to the 7 printed digits."""
    import time
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    c = oracle.Case(fx)
    lib, h = _cp_hooks(c, cctx)
    c.init()
    from tests.test_oracle_case import check_force_dist
    hist = [c.force_nondim(0)]
    t1 = time.perf_counter()
    dists = 0
    for it in range(nsteps):
        _step(c, lib, h, cctx, it + 1)
        hist.append(c.force_nondim(0))
        dists += check_force_dist(c, fx, it + 1) > 0      # the reference's r01b01ForceDistNNNNN.csv.ref of this step
    t2 = time.perf_counter()
    hist = np.array(hist)
    ref = np.array(fx["ref_ForceNonDim"]["rows"])
    ulp = 10.0 ** (np.floor(np.log10(np.abs(ref[:, 1]))) - 6)
    dev = np.abs(hist[:, 0] - ref[:len(hist), 1]) / ulp
    print(f"{name}: {nsteps} steps, wake + collocation-point stage on the device, in {t2 - t1:.2f} s "
          f"({nsteps / (t2 - t1):.1f} timesteps/s incl. the driver's host work); RHS stages {lib.case_hooks_cp_rhs_calls(h)}, "
          f"force stages {lib.case_hooks_cp_force_calls(h)}; max deviation from the golden file {dev.max():.2f} units of "
          f"the 7th digit")
    assert lib.case_hooks_cp_rhs_calls(h) == nsteps and lib.case_hooks_cp_force_calls(h) >= nsteps
    assert dev.max() <= 1.0, (dev.max(), int(dev.argmax()))
    # the sectional loads came from the device (vlc_rotor_calc_force): the golden distributions to 7 digits as well
    assert dists == len(fx["ref_ForceDists"])
    lib.case_gpu_hooks_free(h)


def _short_caradonna(fx):
    fx["config"]["nt"] = 40
    fx["geom"][0]["nNwake"] = 12


@pytest.mark.parametrize("name,nsteps,mutate,resident",
                         [("caradonna", 30, _short_caradonna, True), ("simplewing", 30, _mut(spanwiseLiftSwitch=1), True),
                          ("katzNplotkin_AR04", 25, None, False), ("two_body", 16, None, True)])
def test_cp_stage_vs_cpu_driver(cctx, oracle, name, nsteps, mutate, resident):
    """CL/CT, circulations and the sectional lift coefficients against the CPU driver over the window.  (The fdScheme 2 / 4 /
    5 cases of the same body are in tests/test_zzz_gpu_first_run.py.)"""
    check_cp_stage_vs_cpu_driver(cctx, oracle, name, nsteps, mutate, resident)


def check_cp_stage_vs_cpu_driver(cctx, oracle, name, nsteps, mutate, resident):
    fx = _two_body() if name == "two_body" else json.loads((GOLDEN / f"{name}.json").read_text())
    fx["config"]["rotorForcePlot"] = 1
    if mutate:
        mutate(fx)
    a, b = oracle.Case(fx), oracle.Case(fx)
    lib, h = _cp_hooks(b, cctx, resident)
    a.init()
    b.init()
    worst = [0.0, 0.0, 0.0]
    for it in range(nsteps):
        a.step()
        _step(b, lib, h, cctx, it + 1)
        for ir in range(a.nr):
            fa, fb = a.force_nondim(ir), b.force_nondim(ir)
            ga, gb = a.rotor(ir).vec(0), b.rotor(ir).vec(0)
            ca, cb = a.rotor(ir).sec(0, "secCL"), b.rotor(ir).sec(0, "secCL")
            worst[0] = max(worst[0], abs(fb[0] / fa[0] - 1.0))
            worst[1] = max(worst[1], float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
            worst[2] = max(worst[2], float(np.max(np.abs(cb - ca)) / np.max(np.abs(ca))))
    print(f"{name}: {nsteps} steps, collocation-point stage on the device ({'resident' if resident else 'per-sweep'} wake): "
          f"max rel err CL/CT {worst[0]:.3e}, gamVec {worst[1]:.3e}, secCL {worst[2]:.3e}")
    assert lib.case_hooks_cp_rhs_calls(h) == nsteps
    assert max(worst) < TOL_HISTORY, worst
    lib.case_gpu_hooks_free(h)
