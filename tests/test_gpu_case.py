"""GPU parity at the level the north-star states it: the reference's driver loop (oracle/vlc_case.c, a restatement of
src/main.f90) run twice on the same case -- once with the CPU oracle at its five hot-path call sites, once with
every one of them forwarded to the C ABI (tests/case_hooks.py: what the iso_c_binding shim does) -- must give
CT/CL and circulation histories within 1e-8 relative over the first 50 steps."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.case_hooks import gpu_hooks

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
TOL_HISTORY = 1e-8


def _run_pair(oracle, ctx, name, nsteps, mutate=None):
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    if mutate:
        mutate(fx)
    a, b = oracle.Case(fx), oracle.Case(fx)
    h = gpu_hooks(b, ctx)
    b.set_hooks(h)
    a.init()
    try:
        b.init()
    except RuntimeError as e:
        raise AssertionError(f"init: {e}; hook errors: {h.errors!r}") from e
    assert not h.errors, h.errors
    out = []
    for it in range(nsteps):
        a.step()
        try:
            b.step()
        except RuntimeError as e:
            raise AssertionError(f"step {it + 1}: {e}; hook errors: {h.errors!r}") from e
        assert not h.errors, h.errors
        fa, fb = a.force_nondim(0), b.force_nondim(0)
        ga, gb = a.rotor(0).vec(0), b.rotor(0).vec(0)
        out.append((abs(fb[0] / fa[0] - 1.0), np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
    return a, b, np.array(out), h


def test_katzNplotkin_50_steps_CL_and_circulation(ctx, oracle):
    """BASELINE.json configs[1]: tests/katzNplotkin-AR04.case, fdScheme 3."""
    a, b, err, h = _run_pair(oracle, ctx, "katzNplotkin_AR04", 50)
    print(f"K&P AR-4, 50 steps: max rel CL err {err[:, 0].max():.3e}, max rel gamVec err {err[:, 1].max():.3e}, "
          f"{h.stats['calls']} uploads")
    assert err[:, 0].max() < TOL_HISTORY and err[:, 1].max() < TOL_HISTORY
    # the reference's own wake records describe a lattice: the shared-node kernel did the work (no silent fallback)
    info = ctx.rotor_info(0)
    # ns = 26 columns, a small wake (4160 rings at most): one strip width without padding, 13 strips of width 2, 50 rows
    # (+1 record per strip); from 2e4 rings on the cover would be 6 strips of width 4 + one tail strip of width 2
    assert info["shared_active"] == 1 and info["strip_width"] == 2 and info["tail_strip_width"] == 0, info
    assert info["lattice_records"] == 13 * 51, info
    # and the GPU-driven run still reproduces the reference's golden file to its 7 printed digits
    fx = json.loads((GOLDEN / "katzNplotkin_AR04.json").read_text())
    ref50 = fx["ref_ForceNonDim"]["rows"][50][1]
    assert abs(b.force_nondim(0)[0] - ref50) < 5e-8
    # wake node positions after 50 steps of predictor-corrector convection
    wa, wb = a.rotor(0).waN(0), b.rotor(0).waN(0)
    assert np.max(np.abs(wa - wb)) < 1e-9


def test_elevate_rotor_50_steps_CT_and_circulation(ctx, oracle):
    """5-blade axisymmetric rotor from a PLOT3D grid with dissipation, roll-up into the far wake (after 30 rows)."""
    a, b, err, h = _run_pair(oracle, ctx, "elevateTest", 50)
    print(f"elevateTest, 50 steps: max rel CT err {err[:, 0].max():.3e}, max rel gamVec err {err[:, 1].max():.3e}")
    assert err[:, 0].max() < TOL_HISTORY and err[:, 1].max() < TOL_HISTORY
    for pred in (False, True):       # 5 blades (4 of them rotated copies), far wake present: still a lattice per blade
        info = ctx.rotor_info(0, pred)
        assert info["shared_active"] == 1 and info["strip_width"] == 4, info
    fx = json.loads((GOLDEN / "elevateTest.json").read_text())
    assert abs(b.force_nondim(0)[0] - fx["ref_ForceNonDim"]["rows"][50][1]) < 5e-9
    assert np.max(np.abs(a.rotor(0).waF(0) - b.rotor(0).waF(0))) < 1e-9


def test_caradonna_two_blade_hover_20_steps(ctx, oracle):
    """BASELINE.json configs[2] (tutorials/caradonna.case), shortened: 2 blades convected independently."""
    def short(fx):
        fx["config"]["nt"] = 40
        fx["geom"][0]["nNwake"] = 12      # roll-up into the far wake starts inside the window
    a, b, err, h = _run_pair(oracle, ctx, "caradonna", 20, short)
    print(f"caradonna (short), 20 steps: max rel CT err {err[:, 0].max():.3e}, gamVec {err[:, 1].max():.3e}")
    assert err[:, 0].max() < TOL_HISTORY and err[:, 1].max() < TOL_HISTORY


# ------------------------------------------------------------------ native: no Python between the driver and the C ABI

def _native_hooks(case, ctx):
    """tests/native/case_gpu_hooks.c (the C twin of fortran/libGPU.f90) installed into the oracle's driver."""
    import ctypes as C
    import subprocess
    here = Path(__file__).resolve().parent / "native"
    so = here / "libcase_gpu_hooks.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(here)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.case_gpu_hooks_install.restype = C.c_void_p
    lib.case_gpu_hooks_install.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.case_gpu_hooks_uploads.restype = C.c_long
    lib.case_gpu_hooks_uploads.argtypes = [C.c_void_p]
    lib.case_gpu_hooks_last_rc.argtypes = [C.c_void_p]
    lib.case_gpu_hooks_free.argtypes = [C.c_void_p]
    case.init_rotors()
    h = lib.case_gpu_hooks_install(case.h, ctx.h, case.nr)
    assert h, "vlc_rotor_define failed"
    return lib, h


@pytest.mark.parametrize("name,nsteps", [("katzNplotkin_AR04", 160), ("elevateTest", 150)])
def test_native_gpu_run_reproduces_reference_golden_history(ctx, name, nsteps, oracle):
    """The reference's driver loop with its hot path on the GPU, natively through the C ABI, over the WHOLE history of
    the reference's two golden cases: every row of r01ForceNonDim.csv.ref to the 7 printed digits."""
    import time
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    c = oracle.Case(fx)
    lib, h = _native_hooks(c, ctx)
    t0 = time.perf_counter()
    try:
        c.init()
    except RuntimeError as e:
        raise AssertionError(f"init: {e}; library: rc={lib.case_gpu_hooks_last_rc(h)} {ctx.lib.vlc_last_error(ctx.h)}") from e
    hist = [c.force_nondim(0)]
    t1 = time.perf_counter()
    pairs = 0.0
    for it in range(nsteps):
        try:
            c.step()
        except RuntimeError as e:
            raise AssertionError(f"step {it + 1}: {e}; rc={lib.case_gpu_hooks_last_rc(h)} {ctx.lib.vlc_last_error(ctx.h)}") from e
        pairs += c.pairs_last_step
        hist.append(c.force_nondim(0))
    t2 = time.perf_counter()
    hist = np.array(hist)
    ref = np.array(fx["ref_ForceNonDim"]["rows"])
    ulp = 10.0 ** (np.floor(np.log10(np.abs(ref[:, 1]))) - 6)
    d = np.abs(hist[:, 0] - ref[:len(hist), 1]) / ulp
    print(f"{name}: {nsteps} steps natively through the C ABI in {t2 - t1:.2f} s ({(t2 - t1) / nsteps * 1e3:.1f} ms/step, "
          f"{nsteps / (t2 - t1):.1f} timesteps/s incl. the driver's host work; init {t1 - t0:.2f} s; "
          f"{pairs:.3e} pair interactions, {lib.case_gpu_hooks_uploads(h)} uploads); "
          f"max deviation from the golden file {d.max():.2f} units of the 7th digit")
    assert d.max() <= 1.0, (d.max(), int(d.argmax()))
    lib.case_gpu_hooks_free(h)


@pytest.mark.parametrize("name,nsteps", [("simplewing", 40), ("tr1208", 30)])
def test_native_gpu_vs_cpu_tutorial_cases(ctx, oracle, name, nsteps):
    """BASELINE.json configs[0] (tutorials/simplewing.case, whole run) and configs[3] (tutorials/tr1208.case: swept wing
    from a PLOT3D file, dissipation on): CL and circulation histories, GPU through the native shim vs CPU oracle."""
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    a, b = oracle.Case(fx), oracle.Case(fx)
    lib, h = _native_hooks(b, ctx)
    a.init()
    b.init()
    worst = [0.0, 0.0]
    for it in range(nsteps):
        a.step()
        b.step()
        fa, fb = a.force_nondim(0), b.force_nondim(0)
        ga, gb = a.rotor(0).vec(0), b.rotor(0).vec(0)
        worst[0] = max(worst[0], abs(fb[0] / fa[0] - 1.0))
        worst[1] = max(worst[1], float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
    print(f"{name}: {nsteps} steps, max rel CL err {worst[0]:.3e}, max rel gamVec err {worst[1]:.3e}, "
          f"shared-node form active: {ctx.rotor_info(0)['shared_active']}")
    assert worst[0] < TOL_HISTORY and worst[1] < TOL_HISTORY
    assert ctx.rotor_info(0)["shared_active"] == 1
    assert np.max(np.abs(a.rotor(0).waN(0) - b.rotor(0).waN(0))) < 1e-9
    lib.case_gpu_hooks_free(h)


def test_native_gpu_vs_cpu_two_rotor_case(ctx, oracle):
    """nr = 2 (wing + 2-blade rotor with far wake): the cross-rotor terms of the RHS and of both wake sweeps
    (main.f90:147-153, :549-561, :814-827) on the GPU vs the CPU oracle."""
    from tests.test_oracle_case import two_body_case
    fx = two_body_case()
    a, b = oracle.Case(fx), oracle.Case(fx)
    lib, h = _native_hooks(b, ctx)
    a.init()
    b.init()
    worst = 0.0
    for it in range(16):
        a.step()
        b.step()
        for ir in range(2):
            fa, fb = a.force_nondim(ir), b.force_nondim(ir)
            ga, gb = a.rotor(ir).vec(0), b.rotor(ir).vec(0)
            worst = max(worst, abs(fb[0] / fa[0] - 1.0), float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
    print(f"wing + rotor, 16 steps: max rel CL/CT/gamVec err {worst:.3e}; shared-node form: "
          f"wing wake {ctx.rotor_info(0)['shared_active']}, rotor wake {ctx.rotor_info(1)['shared_active']}")
    assert worst < TOL_HISTORY
    assert ctx.rotor_info(0)["shared_active"] == 1 and ctx.rotor_info(1)["shared_active"] == 1
    lib.case_gpu_hooks_free(h)
