"""GPU tests written after the round's GPU minutes were spent: they have not run on a B200 yet, so they come last (file
name) and a surprise here cannot hide the results of the suite before them.  Their logic is covered on the CPU -- same
orchestration on the CPU backend bit-identical to the driver's inline loop (tests/test_staged_hooks.py,
tests/test_prescribed_wake.py), host builds of the host + device sources against the oracle, test bodies run on an
emulated context -- what is left for the GPU is the CUDA side of each call.

 * Prescribed far wake (classdef.f90:4826-4828, :5170-5218, :998-1066): the 240 helix filaments per blade as sources of
   every sweep (the abs(gam) > eps rule of classdef.f90:1471-1476), uploaded (vlc_rotor_put_pfwake) or made on the device
   from its own far rows (vlc_rotor_updatePrescribedWake: pf_fit_kernel + pf_helix_kernel; tests/native/case_gpu_hooks.c:
   g_convect is the C twin of fortran/libGPU.f90: gpu_convect).  PARITY UNPINNED for this feature: no shipped case
   enables it.
 * rotor%burst_wake (classdef.f90:4911-4917): vlc_rotor_burst_wake per call and in a resident run.  Parity unpinned too.
 * rotor%calc_skew (classdef.f90:4919-4936): vlc_rotor_calc_skew per call.  Parity unpinned too.
 * fdScheme 2 / 4 / 5 with the collocation-point stage on the device (vlc_rotor_wakevel_copy / _lincomb, vel2 / vel3)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from tests.test_prescribed_wake import _with_prescribed_wake, rotor_fx
from tests.test_cp_stage_host import _mut
from tests.test_zz_gpu_cp_stage import (TOL_HISTORY, _cp_hooks, _short_caradonna, _step, cctx,  # noqa: F401  (fixture)
                                        check_cp_stage_vs_cpu_driver)

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("gen,mode", [(0, "resident+cp"), (2, "resident"), (0, "per-sweep")])
def test_prescribed_wake_case_vs_cpu_driver(cctx, oracle, gen, mode):  # noqa: F811
    """resident: the helix is made on the device (vlc_rotor_updatePrescribedWake after each vlc_rotor_convectwake);
    per-sweep: the driver makes it (rotor%updatePrescribedWake) and the shim uploads it with the wake (vlc_rotor_put_pfwake)."""
    fx = rotor_fx(small=(mode != "resident+cp"))                         # elevateTest (5 blades) once, else the small rotor
    _with_prescribed_wake(gen)(fx)                                       # prescWakeNt = 12
    fx["config"]["rotorForcePlot"] = 1
    a, b = oracle.Case(fx), oracle.Case(fx)
    cp = mode == "resident+cp"
    if cp:
        lib, h = _cp_hooks(b, cctx, True)
    elif mode == "resident":
        from tests.test_gpu_resident import _resident_hooks
        lib, h = _resident_hooks(b, cctx)
    else:
        from tests.test_gpu_case import _native_hooks
        lib, h = _native_hooks(b, cctx)
    a.init()
    b.init()
    worst = [0.0, 0.0]
    for it in range(20):
        a.step()
        _step(b, lib, h, cctx, it + 1)
        fa, fb = a.force_nondim(0), b.force_nondim(0)
        ga, gb = a.rotor(0).vec(0), b.rotor(0).vec(0)
        worst[0] = max(worst[0], abs(fb[0] / fa[0] - 1.0))
        worst[1] = max(worst[1], float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
    if mode != "per-sweep":
        assert lib.case_gpu_hooks_download_wake(h) == 0                  # the helix made on the device, back in b's records
    ra, rb = a.rotor(0), b.rotor(0)
    for ib in range(ra.nb):
        wa, wb = ra.wapF(ib), rb.wapF(ib)
        assert np.all(np.abs(wa[:, 12]) > 0)
        assert np.max(np.abs(wb[:, :6] - wa[:, :6])) < 1e-9 * np.max(np.abs(wa[:, :6]))
        assert np.max(np.abs(wb[:, 12] - wa[:, 12])) < 1e-9 * np.max(np.abs(wa[:, 12]))
    print(f"{fx['name']} ({ra.nb} blades) + prescribed far wake (prescWakeGenNt = {gen}), 20 steps, {mode}: max rel err CT {worst[0]:.3e}, "
          f"gamVec {worst[1]:.3e}")
    assert max(worst) < TOL_HISTORY, worst
    lib.case_gpu_hooks_free(h)


@pytest.mark.parametrize("predicted", [False, True])
def test_prescribed_filaments_are_sources_of_vind_bywake(cctx, oracle, predicted):  # noqa: F811
    """Per call: rotor%vind_bywake(P[, 'P']) with a live helix (classdef.f90:1471-1476, :1501-1507) against the oracle's
    loop at the per-call bar (1e-12 of the velocity scale); without vlc_rotor_put_pfwake the difference is the helix's own
    contribution, so the comparison is not vacuous."""
    from tests.test_zz_gpu_cp_stage import TOL, _define, _developed
    fx = rotor_fx()
    _with_prescribed_wake(0)(fx)
    case = _developed(oracle, fx, 16)
    rot = case.rotor(0)
    assert np.all(np.abs(rot.wapF(0, predicted)[:, 12]) > 0)
    _define(cctx, rot, 0)
    if predicted:
        for ib in range(rot.nb):
            cctx.rotor_put_nwake(0, ib, rot.waN(ib, True), predicted=True)
            cctx.rotor_put_fwake(0, ib, rot.waF(ib, True), predicted=True)
    rng = np.random.default_rng(11)
    P = np.concatenate([rng.uniform(-1.5, 1.5, (200, 3)) * np.max(np.abs(rot.wapF(0, predicted)[:, 0:2])),
                        rot.wapF(0, predicted)[::7, 0:3],                 # on the helix's own nodes: the c2 <= eps^2 rule
                        0.5 * (rot.wapF(2, predicted)[::9, 0:3] + rot.wapF(2, predicted)[::9, 3:6])])
    ref = rot.vind_points(1, P, predicted)
    without = cctx.rotor_vind_bywake(0, P, predicted)
    for ib in range(rot.nb):
        cctx.rotor_put_pfwake(0, ib, rot.wapF(ib, predicted), predicted=predicted)
    got = cctx.rotor_vind_bywake(0, P, predicted)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(got - ref)) < TOL * scale, np.max(np.abs(got - ref)) / scale
    assert np.max(np.abs(without - ref)) > 1e-6 * scale


@pytest.mark.parametrize("gen,axisym", [(0, 1), (2, 0)])
def test_update_prescribed_wake_on_the_device_vs_oracle(cctx, oracle, gen, axisym):  # noqa: F811
    """vlc_rotor_updatePrescribedWake alone, on the oracle's far wake: the fit parameters (sums, products, sqrt: no
    transcendental) BIT-IDENTICAL, the end points within 1e-13 of the helix radius (cos / sin / atan2 of CUDA vs libm), gam
    and rVc bit-identical; two successive updates from a zero fit (the relaxation carries state), both record sets.
    The body also runs on the CPU against the host build of the product's source (tests/test_prescribed_wake.py)."""
    from tests.test_prescribed_wake import check_update_prescribed_wake
    check_update_prescribed_wake(cctx, oracle, gen, axisym)


def test_update_prescribed_wake_error_behaviour(cctx):  # noqa: F811
    """error stop "Prescribed far wake only implemented for shaft along Z-axis" (classdef.f90:1010-1012) -> VLC_ERR_ARG with
    the message; no far wake / no row to fit -> VLC_ERR_STATE."""
    cctx.rotor_define(0, 2, 2, 3, 4, 0, 1)
    with pytest.raises(Exception, match="far wake"):
        cctx.rotor_updatePrescribedWake(0, 0.1, 0, "C")
    cctx.rotor_define(0, 2, 2, 3, 4, 5, 1)
    cctx.rotor_set_frame(0, [0.0, 1e-3, 1.0], [0.0, 0.0, 0.0])
    with pytest.raises(Exception, match="shaft along Z-axis"):
        cctx.rotor_updatePrescribedWake(0, 0.1, 0, "C")
    cctx.rotor_set_frame(0, [0.0, 0.0, 1.0], [0.0, 0.0, 0.0])
    with pytest.raises(Exception, match="no far-wake row"):       # rowFar = nFwake + 1 after vlc_rotor_define
        cctx.rotor_updatePrescribedWake(0, 0.1, 0, "C")
    with pytest.raises(Exception, match="no far-wake row"):
        cctx.rotor_updatePrescribedWake(0, 0.1, 5, "C")           # rowStart = 0
    cctx.rotor_updatePrescribedWake(0, 0.1, 2, "C")               # rows 3..5 of an all-zero far wake: a degenerate helix, no error
    # restart resume: records + fit parameters round trip, per blade and record set
    rng = np.random.default_rng(3)
    w = rng.standard_normal((240, 13))
    cctx.rotor_put_pfwake(0, 1, w, predicted=True)
    cctx.rotor_put_pfwake_helix(0, 1, [1.5, 2.5], predicted=True)
    got, hx = cctx.rotor_get_pfwake(0, 1, predicted=True)
    assert np.array_equal(got, w) and np.array_equal(hx, [1.5, 2.5])
    assert np.array_equal(cctx.rotor_get_pfwake(0, 1, predicted=False)[1], [0.0, 0.0])


def _burst_case():
    fx = rotor_fx()
    g = fx["geom"][0]
    g["nNwake"], g["wakeTruncateNt"], g["skewLimit"] = 6, 14, 0.004
    fx["config"]["wakeBurst"] = 2
    return fx


def test_burst_wake_on_the_device_vs_oracle(cctx, oracle):  # noqa: F811
    """vlc_rotor_burst_wake against rotor%burst_wake() of the oracle (classdef.f90:4911-4917, :2306-2339) on a developed far
    wake, with a limit placed in the widest gap of the skew values so that the last bits of acos cannot decide: far-wake
    records bit-identical (only rVc of the burst pairs changes).  Body shared with the CPU run on the host build."""
    from tests.test_prescribed_wake import check_burst_wake
    check_burst_wake(cctx, oracle, _burst_case())


def test_wake_burst_resident_vs_cpu_driver(cctx, oracle):  # noqa: F811
    fx = _burst_case()
    fx["config"]["rotorForcePlot"] = 1
    a, b = oracle.Case(fx), oracle.Case(fx)
    lib, h = _cp_hooks(b, cctx, True)
    a.init()
    b.init()
    worst = 0.0
    for it in range(18):
        a.step()
        _step(b, lib, h, cctx, it + 1)
        worst = max(worst, abs(b.force_nondim(0)[0] / a.force_nondim(0)[0] - 1.0))
    assert lib.case_gpu_hooks_download_wake(h) == 0
    chord = fx["geom"][0]["chord"]
    for ib in range(a.rotor(0).nb):
        assert np.array_equal(a.rotor(0).waF(ib)[:, 9] == chord, b.rotor(0).waF(ib)[:, 9] == chord)   # the same filaments burst
    assert sum(int(np.sum(a.rotor(0).waF(ib)[:, 9] == chord)) for ib in range(a.rotor(0).nb)) >= 2
    print(f"small rotor + wakeBurst = 2 (skewLimit 0.004), 18 steps, resident: max rel err CT {worst:.3e}")
    assert worst < TOL_HISTORY, worst
    lib.case_gpu_hooks_free(h)


@pytest.mark.parametrize("name,nsteps,mutate",
                         [  # fdScheme 2 (explicit Adams-Bashforth): AB2 + FIRST_STEP + COPY_TO_STEP on the device arrays
                          ("katzNplotkin_AR04", 25, _mut(fdScheme=2)),
                          # fdScheme 4 / 5 (third / fourth order multistep): vlc_rotor_wakevel_copy / _lincomb, vel2 / vel3
                          ("katzNplotkin_AR04", 25, _mut(fdScheme=4)),
                          ("caradonna", 30, lambda fx: (_short_caradonna(fx), fx["config"].update(fdScheme=5)))])
def test_cp_stage_vs_cpu_driver_other_fd_schemes(cctx, oracle, name, nsteps, mutate):  # noqa: F811
    check_cp_stage_vs_cpu_driver(cctx, oracle, name, nsteps, mutate, True)


@pytest.mark.parametrize("axisym", [1, 0])
def test_calc_skew_on_the_device_vs_oracle(cctx, oracle, axisym):  # noqa: F811
    """vlc_rotor_calc_skew (rec_skew_kernel) against rotor%calc_skew() of the oracle: records bit-identical.  Body shared with
    the CPU run on the host build."""
    from tests.test_prescribed_wake import check_calc_skew
    check_calc_skew(cctx, oracle, axisym)
