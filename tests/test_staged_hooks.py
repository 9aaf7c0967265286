"""CPU check of the host logic of device-resident stepping: the staged orchestration of tests/native/case_gpu_hooks.c
(h_wake_prestep / h_wake_convect: which wake stage runs when, per fdScheme -- the C twin of fortran/libGPU.f90:
gpu_wake_prestep / gpu_wake_convect) driven with its CPU backend, i.e. the oracle's own mutators called stage by stage
through the driver's two optional hooks.  It must reproduce the driver's inline statement of main.f90:466-506, :800-1440
BIT FOR BIT: force histories, circulations, every wake record and velocity array.  On the GPU the same orchestration
runs with the C-ABI backend (tests/test_gpu_resident.py)."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


def _lib():
    here = Path(__file__).resolve().parent / "native"
    so = here / "libcase_gpu_hooks.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(here)], check=True, capture_output=True)
    lib = C.CDLL(str(so))          # links libvolcanor_b200.so: loads without a GPU, nothing in it is called here
    lib.case_cpu_staged_hooks_install.restype = C.c_void_p
    lib.case_cpu_staged_hooks_install.argtypes = [C.c_void_p, C.c_int]
    lib.case_gpu_hooks_free.argtypes = [C.c_void_p]
    lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
    lib.case_hooks_cp_rhs_calls.argtypes = [C.c_void_p]
    lib.case_hooks_cp_rhs_calls.restype = C.c_long
    lib.case_hooks_cp_force_calls.argtypes = [C.c_void_p]
    lib.case_hooks_cp_force_calls.restype = C.c_long
    return lib


def _short_caradonna(fx):
    fx["config"]["nt"] = 40
    fx["geom"][0]["nNwake"] = 12
    fx["geom"][0]["wakeTruncateNt"] = 18          # far wake full after 18 steps: rollup through shiftFwake


def _elevate_short(fd):
    def m(fx):
        fx["config"]["fdScheme"] = fd
        fx["geom"][0]["nNwake"] = 6               # roll-up, far wake and truncation inside a short window
        fx["geom"][0]["wakeTruncateNt"] = 10
    return m


CASES = [("katzNplotkin_AR04", 12, None), ("katzNplotkin_AR04", 8, lambda fx: fx["config"].update(fdScheme=0)),
         ("katzNplotkin_AR04", 10, lambda fx: fx["config"].update(fdScheme=2)), ("elevateTest", 14, _elevate_short(2)),
         ("katzNplotkin_AR04", 8, lambda fx: fx["config"].update(fdScheme=1, wakeDissipation=1)),
         ("caradonna", 22, _short_caradonna), ("elevateTest", 14, _elevate_short(3)), ("elevateTest", 14, _elevate_short(1)),
         ("simplewing", 10, lambda fx: fx["config"].update(wakeStrain=1)),
         # fdScheme 4 / 5: third / fourth order Adams-Bashforth / Adams-Moulton with the histories vel2, vel3
         ("katzNplotkin_AR04", 9, lambda fx: fx["config"].update(fdScheme=4)), ("elevateTest", 14, _elevate_short(4)),
         ("katzNplotkin_AR04", 9, lambda fx: fx["config"].update(fdScheme=5)), ("elevateTest", 14, _elevate_short(5))]


@pytest.mark.parametrize("name,nsteps,mutate", CASES)
def test_staged_orchestration_equals_inline_time_loop(oracle, name, nsteps, mutate):
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    if mutate:
        mutate(fx)
    lib = _lib()
    a, b = oracle.Case(fx), oracle.Case(fx)
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    assert h
    a.init()
    b.init()
    for it in range(nsteps):
        a.step()
        b.step()
        assert np.array_equal(a.force_nondim(0), b.force_nondim(0)), (name, it + 1)
        assert a.pairs_last_step == b.pairs_last_step, (name, it + 1, a.pairs_last_step, b.pairs_last_step)
    ra, rb = a.rotor(0), b.rotor(0)
    assert ra.dims() == rb.dims()
    assert np.array_equal(ra.vec(0), rb.vec(0))
    for ib in range(ra.nb):
        for pred in (False, True):
            assert np.array_equal(ra.waN(ib, pred), rb.waN(ib, pred)), (name, "waN", ib, pred)
            if ra.nFwake:
                assert np.array_equal(ra.waF(ib, pred), rb.waF(ib, pred)), (name, "waF", ib, pred)
        for w in (list(range(8)) + [8, 9, 10, 11]) if ra.nFwake else [0, 1, 2, 3, 8, 9]:
            assert np.array_equal(ra.vel(ib, w), rb.vel(ib, w)), (name, "vel", ib, w)
    lib.case_gpu_hooks_free(h)


def test_staged_orchestration_two_rotors(oracle):
    from tests.test_oracle_case import two_body_case
    fx = two_body_case()
    lib = _lib()
    a, b = oracle.Case(fx), oracle.Case(fx)
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    a.init()
    b.init()
    for it in range(8):
        a.step()
        b.step()
    for ir in range(2):
        assert np.array_equal(a.force_nondim(ir), b.force_nondim(ir))
        for ib in range(a.rotor(ir).nb):
            assert np.array_equal(a.rotor(ir).waN(ib), b.rotor(ir).waN(ib))
    lib.case_gpu_hooks_free(h)


SEC3 = ("secChordwiseResVel", "secDragDir", "secLiftDir", "secForceInertial", "secLift", "secDrag", "secLiftUnsteady")
SEC1 = ("secAlpha", "secCL", "secCD", "secCLu")


@pytest.mark.parametrize("name,nsteps,mutate", [CASES[0], CASES[5], CASES[6], CASES[8]])
def test_cp_stage_orchestration_equals_inline_time_loop(oracle, name, nsteps, mutate):
    """The collocation-point stage (tier 2c) through the C twin's orchestration -- h_cp_rhs_solve / h_cp_forces with
    their write-backs into the driver's records -- against a CPU emulation of the library (own record copies, the g++
    build of cp_stage.cuh, the oracle's sweeps): forces, circulations, wing records and every sectional array of the
    driver BIT-IDENTICAL to its inline statement of main.f90:522-670."""
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    fx["config"]["rotorForcePlot"] = 1
    if mutate:
        mutate(fx)
    lib = _lib()
    a, b = oracle.Case(fx), oracle.Case(fx)
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    assert h and lib.case_hooks_enable_cp(h) == 0
    a.init()
    b.init()
    for it in range(nsteps):
        a.step()
        b.step()
        assert np.array_equal(a.force_nondim(0), b.force_nondim(0)), (name, it + 1)
        assert a.pairs_last_step == b.pairs_last_step, (name, it + 1, a.pairs_last_step, b.pairs_last_step)
    assert lib.case_hooks_cp_rhs_calls(h) == nsteps and lib.case_hooks_cp_force_calls(h) == nsteps + 1
    ra, rb = a.rotor(0), b.rotor(0)
    for w in range(3):
        assert np.array_equal(ra.vec(w), rb.vec(w)), (name, "vec", w)      # gamVec, RHS, gamVecPrev
    for ib in range(ra.nb):
        assert np.array_equal(ra.wiP(ib), rb.wiP(ib)), (name, "wiP", ib)
        assert np.array_equal(ra.waN(ib), rb.waN(ib)), (name, "waN", ib)
        for n in SEC3:
            assert np.array_equal(ra.sec(ib, n, 3), rb.sec(ib, n, 3)), (name, ib, n)
        for n in SEC1:
            assert np.array_equal(ra.sec(ib, n), rb.sec(ib, n)), (name, ib, n)
    lib.case_gpu_hooks_free(h)


def test_cp_stage_orchestration_two_rotors(oracle):
    from tests.test_oracle_case import two_body_case
    fx = two_body_case()
    fx["config"]["rotorForcePlot"] = 1
    lib = _lib()
    a, b = oracle.Case(fx), oracle.Case(fx)
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    assert h and lib.case_hooks_enable_cp(h) == 0
    a.init()
    b.init()
    for it in range(6):
        a.step()
        b.step()
    for ir in range(2):
        assert np.array_equal(a.force_nondim(ir), b.force_nondim(ir))
        assert np.array_equal(a.rotor(ir).vec(0), b.rotor(ir).vec(0))
        for ib in range(a.rotor(ir).nb):
            assert np.array_equal(a.rotor(ir).wiP(ib), b.rotor(ir).wiP(ib))
    lib.case_gpu_hooks_free(h)
