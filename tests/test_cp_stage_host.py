"""CPU check of the collocation-point stage (tier 2c of the C ABI): the host + device routines of
volcanor_b200/csrc/cp_stage.cuh -- per-section loads (blade_calc_secChordwiseResVel, secAlpha, dirLiftDrag, blade_calc_force,
sectional coefficients), blade sums, RHS entries, map_gam -- compiled by g++ (tests/native/cp_stage_host.cpp) and driven
with the loops the CUDA kernels run, against the oracle's restatement of classdef.f90:1704-1896, :2197-2380, :4181-4196
and main.f90:563-603 on developed reference cases: BIT-IDENTICAL (secAlpha to 2 ulp: atan2).  The same routines run on
the GPU in tests/test_zz_gpu_cp_stage.py through the C ABI."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
WP = 104
SEC3 = ("secChordwiseResVel", "secDragDir", "secLiftDir", "secForceInertial", "secLift", "secDrag", "secLiftUnsteady")
SEC1 = ("secAlpha", "secCL", "secCD", "secCLu")
NET = ("forceInertial", "lift", "drag", "liftUnsteady")


def host_lib():
    here = Path(__file__).resolve().parent / "native"
    so = here / "libcp_stage_host.so"
    src = [here / "cp_stage_host.cpp", here.parent.parent / "volcanor_b200" / "csrc" / "cp_stage.cuh"]
    if not so.exists() or any(so.stat().st_mtime < s.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(here), str(so)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    vp, i32, d = C.c_void_p, C.c_int, C.c_double
    lib.cp_host_loads.argtypes = [i32, i32, i32, d, d, d, i32, vp, vp, vp]
    lib.cp_host_rhs.argtypes = [i32, i32, i32, i32, vp, vp]
    lib.cp_host_map_gam.argtypes = [i32, i32, i32, i32, vp, vp]
    return lib


def pack_sections(rot, ib):
    """The 10*ns + 6 block of vlc_rotor_put_sections from the oracle's blade arrays."""
    return np.concatenate([rot.sec(ib, "secTauCapChord", 3).ravel(), rot.sec(ib, "secNormalVec", 3).ravel(),
                           rot.sec(ib, "secCP", 3).ravel(), rot.sec(ib, "secArea").ravel(),
                           rot.sec(ib, "yAxisAziFlap", 3)[0], rot.sec(ib, "zAxisAziFlap", 3)[0]]).copy()


def unpack_loads(a, ns):
    out = {n: a[3 * k:3 * k + 3] for k, n in enumerate(NET)}
    for k, n in enumerate(SEC3):
        out[n] = a[12 + 3 * ns * k:12 + 3 * ns * (k + 1)].reshape(ns, 3)
    for k, n in enumerate(SEC1):
        out[n] = a[12 + 21 * ns + ns * k:12 + 21 * ns + ns * (k + 1)]
    return out


def force_params(oracle_case, rot):
    o = np.zeros(4)
    rot.lib.orc_rotor_get_force_params(rot.h, o.ctypes.data)
    return dict(Omega=o[0], spanwiseLiftSwitch=int(o[1]), axisym=int(o[2]), nbConvect=int(o[3]))


def _mut(**kw):
    def m(fx):
        for k, v in kw.items():
            if k in fx["config"]:
                fx["config"][k] = v
            else:
                for g in fx["geom"]:
                    g[k] = v
    return m


CASES = [("katzNplotkin_AR04", 6, None), ("caradonna", 5, _mut(nNwake=8)), ("elevateTest", 5, _mut(nNwake=6)),
         ("simplewing", 4, _mut(spanwiseLiftSwitch=1)), ("tr1208", 3, None)]


@pytest.mark.parametrize("name,nsteps,mutate", CASES)
def test_loads_bit_identical_to_the_oracle(oracle, name, nsteps, mutate):
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    fx["config"]["rotorForcePlot"] = 1
    if mutate:
        mutate(fx)
    lib = host_lib()
    c = oracle.Case(fx)
    c.init()
    for _ in range(nsteps):
        c.step()
    cfg = c.config
    rng = np.random.default_rng(7)
    for ir in range(c.nr):
        rot = c.rotor(ir)
        fp = force_params(c, rot)
        nb, nc, ns, nbc = rot.nb, rot.nc, rot.ns, fp["nbConvect"]
        # a state worth checking: velCPTotal as the last force evaluation left it, gamPrev perturbed so that the
        # unsteady terms are not zero
        for ib in range(nbc):
            rot.wiP(ib)[:, :, 50] += 1e-3 * rng.standard_normal((ns, nc))
        wiP = np.concatenate([rot.wiP(ib).reshape(-1) for ib in range(nbc)]).copy()
        sec = np.concatenate([pack_sections(rot, ib) for ib in range(nbc)])
        loads = np.zeros(nbc * (12 + 25 * ns))
        lib.cp_host_loads(nbc, nc, ns, cfg.density, cfg.dt, fp["Omega"], fp["spanwiseLiftSwitch"], wiP.ctypes.data,
                          sec.ctypes.data, loads.ctypes.data)
        rot.lib.orc_rotor_calc_secAlpha(rot.h)
        rot.lib.orc_rotor_calc_force(rot.h, cfg.density, cfg.dt)
        for ib in range(nbc):
            got = unpack_loads(loads[ib * (12 + 25 * ns):(ib + 1) * (12 + 25 * ns)], ns)
            for n in NET:
                assert np.array_equal(got[n], rot.sec(ib, n, 3)[0]), (name, ir, ib, n)
            for n in SEC3:
                assert np.array_equal(got[n], rot.sec(ib, n, 3)), (name, ir, ib, n)
            for n in SEC1[1:]:
                assert np.array_equal(got[n], rot.sec(ib, n)), (name, ir, ib, n)
            np.testing.assert_allclose(got["secAlpha"], rot.sec(ib, "secAlpha"), rtol=5e-16, atol=1e-18)
            assert np.any(got["secLift"] != 0.0) and np.any(got["secLiftUnsteady"] != 0.0)
            w = wiP[ib * ns * nc * WP:(ib + 1) * ns * nc * WP].reshape(ns, nc, WP)
            assert np.array_equal(w, rot.wiP(ib)), (name, ir, ib, "wing records")   # gamPrev, gamTrapz, delP, forces ...


@pytest.mark.parametrize("name,nsteps,mutate", [("caradonna", 3, _mut(nNwake=6)), ("elevateTest", 3, _mut(nNwake=6))])
def test_rhs_and_map_gam_follow_the_oracle(oracle, name, nsteps, mutate):
    """RHS entries (with the axisymmetric replication and the -1 factor) and map_gam from the same records."""
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    mutate(fx)
    lib = host_lib()
    c = oracle.Case(fx)
    c.init()
    for _ in range(nsteps):
        c.step()
    rot = c.rotor(0)
    fp = force_params(c, rot)
    nb, npb = rot.nb, rot.nc * rot.ns
    wiP = np.concatenate([rot.wiP(ib).reshape(-1) for ib in range(nb)]).copy()
    # the records hold the velCP the last RHS was formed from (main.f90:563-603); rot.vec(1) is that RHS
    rhs = np.zeros(rot.N)
    lib.cp_host_rhs(rot.N, npb, fp["nbConvect"], fp["axisym"], wiP.ctypes.data, rhs.ctypes.data)
    assert np.array_equal(rhs, rot.vec(1))
    assert np.any(rhs != 0.0)
    if fp["axisym"]:
        assert np.array_equal(rhs[npb:2 * npb], rhs[:npb])
    # map_gam: a fresh gamVec goes into the records exactly where rotor_map_gam puts it
    g = np.random.default_rng(3).standard_normal(rot.N)
    lib.cp_host_map_gam(nb, npb, fp["nbConvect"], fp["axisym"], g.ctypes.data, wiP.ctypes.data)
    rot.vec(0)[:] = g
    rot.lib.orc_rotor_map_gam(rot.h)
    ref = np.concatenate([rot.wiP(ib).reshape(-1) for ib in range(nb)])
    assert np.array_equal(wiP, ref)


# ---- the bodies of the GPU tests (tests/test_zz_gpu_cp_stage.py) on the CPU stand-in for the library's tier 2c

@pytest.mark.parametrize("name,nsteps,mutate", [("katzNplotkin_AR04", 6, None), ("elevateTest", 5, _mut(nNwake=6))])
def test_gpu_test_body_calc_force_on_the_emulation(oracle, name, nsteps, mutate):
    from tests.cp_stage_emulation import EmulatedCpContext
    from tests.test_zz_gpu_cp_stage import check_calc_force
    check_calc_force(EmulatedCpContext(host_lib()), oracle, name, nsteps, mutate)


@pytest.mark.parametrize("case", ["elevateTest", "two_body"])
def test_gpu_test_body_rhs_solve_velcptotal_on_the_emulation(oracle, case):
    from tests.cp_stage_emulation import EmulatedCpContext
    from tests.test_zz_gpu_cp_stage import check_rhs_solve_velcptotal
    check_rhs_solve_velcptotal(EmulatedCpContext(host_lib()), oracle, case)


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_reference_wing1x3_force_kat_through_the_product_source(oracle, sign):
    """The reference's own known-answer test of the loads (tests/wing1x3_test.f90:150-194, 15 digits; NegPitch variant
    :150-194, 6 decimals) evaluated by the PRODUCT's source (cp_stage.cuh, host build): geometry and circulations from
    the oracle's restatement of rotor%init / the solve, the loads from section_loads + blade_sum_loads."""
    from tests.test_oracle_case import _base_geom, O_NCAP, O_VELCP, O_VELCPM, O_VELCPTOT
    fx = {"config": dict(nt=1, dt=0.00625, density=1.2, fdScheme=3),
          "geom": [_base_geom(spanSpacing=2, ns=3, chord=0.3, span=2.0, Omega=0.0, shaftAxis=[0, 0, 0], velBody=[-6, 0, 0],
                              theta0=7.0 * sign, symmetricTau=1, pivotLE=0.25, apparentViscCoeff=5.0)]}
    fx0 = json.loads(json.dumps(fx))
    fx0["geom"][0]["theta0"] = 0.0
    c0 = oracle.Case(fx0)
    c0.init_rotors()
    r0 = c0.rotor(0)
    assert r0.calcAIC() == 0
    c = oracle.Case(fx)
    c.init_rotors()
    r = c.rotor(0)
    w = r.wiP(0)
    for j in range(3):
        for off in (O_VELCP, O_VELCPM, O_VELCPTOT):
            w[j, 0, off:off + 3] = [6.0, 0.0, 0.0]
    lib = host_lib()
    rec = np.ascontiguousarray(w.reshape(-1))
    rhs = np.zeros(3)
    lib.cp_host_rhs(3, 3, 1, 0, rec.ctypes.data, rhs.ctypes.data)                 # RHS = -velCP.nCap
    g = np.ascontiguousarray(r0.AIC(inverse=True) @ rhs)
    assert np.max(np.abs(g - sign * np.array([-0.240131, -0.249833, -0.240131]))) < 1e-6
    lib.cp_host_map_gam(1, 3, 1, 0, g.ctypes.data, rec.ctypes.data)
    loads = np.zeros(12 + 25 * 3)
    sec = pack_sections(r, 0)
    lib.cp_host_loads(1, 1, 3, 1.2, 0.00625, 0.0, 0, rec.ctypes.data, sec.ctypes.data, loads.ctypes.data)
    got = unpack_loads(loads, 3)
    rec = rec.reshape(3, 1, WP)
    tol = 1e-11 if sign > 0 else 1e-6
    assert np.max(np.abs(rec[:, 0, 95] - sign * np.array([28.7727659410054, 29.9353746086400, 28.7727659410054]))) < tol
    nf = np.array([[0.525977713977048, 0.0, sign * 4.28374471602321], [1.09446133444262, 0.0, sign * 8.91367225972409],
                   [0.525977713977048, 0.0, sign * 4.28374471602321]])
    assert np.max(np.abs(rec[:, 0, 85:88] - nf)) < tol
    assert np.max(np.abs(got["secForceInertial"] - nf)) < tol and np.max(np.abs(got["secLift"] - nf * [0, 0, 1])) < tol
    assert np.max(np.abs(got["secLiftDir"] - [0, 0, 1])) < 1e-12 and np.max(np.abs(got["secDragDir"] - [1, 0, 0])) < 1e-12
    assert np.max(np.abs(got["secCL"] - sign * np.array([1.32214343087136, 1.37556670674755, 1.32214343087136]))) < tol
    assert np.max(np.abs(got["forceInertial"] - [2.14641676239672, 0.0, sign * 17.4811616917705])) < tol
    assert np.max(np.abs(got["lift"] - [0.0, 0.0, sign * 17.4811616917705])) < tol
