"""The PRODUCT-side driver of an unmodified reference case (volcanor_b200/csrc/case_driver.cpp + volcanor_b200/run_case.py,
SURVEY 8f rank 4): namelists / PLOT3D in, the reference's force history out, every stage of the hot path in the CUDA library.

CPU part: the driver builds, takes every key of the shipped case files, and REFUSES to run without the library context (no CPU
path).  GPU part: both golden histories of the reference (tests/katzNplotkin-AR04.case, tests/elevateTest.case -- fixtures
under tests/golden/ made from the reference's own referenceResults) every row to the 7 printed digits with the oracle NOT in
the loop; against the oracle's CPU driver (the checker) on the other BASELINE configs, the other time-marching schemes, two
rotors, and sub-iterations (ntSub > 0: vlc_rotor_reset_velCP); the same through a multi-GPU handle.
"""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


def _fixture(name):
    return json.loads((GOLDEN / f"{name}.json").read_text())


def test_driver_builds_and_takes_every_key_of_the_shipped_cases():
    from volcanor_b200 import api
    from volcanor_b200.run_case import CaseDriver
    api.build_case_driver()
    for name in ("katzNplotkin_AR04", "elevateTest", "caradonna", "simplewing", "tr1208"):
        fx = _fixture(name)
        with pytest.raises(api.VlcError, match="no CPU path"):        # no context -> no run
            CaseDriver(fx, None)


def test_product_driver_does_not_use_the_oracle():
    src = (Path(__file__).resolve().parent.parent / "volcanor_b200" / "csrc" / "case_driver.cpp").read_text()
    body = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("//"))
    assert "orc_" not in body and "vlc_oracle" not in body and "vlc_case.h" not in body
    import subprocess

    from volcanor_b200 import api
    api.build_case_driver()
    out = subprocess.run(["ldd", str(api.case_lib_path())], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libvolcanor_b200" in out


def _run(ctx, fx, nsteps):
    from volcanor_b200.run_case import CaseDriver
    d = CaseDriver(fx, ctx)
    d.init()
    hist, gam = [d.force_nondim(0).copy()], []
    for _ in range(nsteps):
        d.step()
        hist.append(d.force_nondim(0).copy())
        gam.append(d.gamvec(0).copy())
    return d, np.array(hist), gam


@pytest.mark.gpu
@pytest.mark.parametrize("name,nsteps", [("katzNplotkin_AR04", 160), ("elevateTest", 150)])
def test_product_driver_reproduces_reference_golden_history(ctx, name, nsteps):
    """The reference's own r01ForceNonDim.csv.ref, every row to the 7 printed digits, from the product driver alone."""
    import time
    fx = _fixture(name)
    t0 = time.perf_counter()
    d, hist, _ = _run(ctx, fx, nsteps)
    dt = time.perf_counter() - t0
    ref = np.array(fx["ref_ForceNonDim"]["rows"])
    for col, refcol in ((0, 1), (6, 7), (8, 9)):                      # CL/CT, CFx, CFz
        r = ref[:len(hist), refcol]
        ulp = 10.0 ** (np.floor(np.log10(np.maximum(np.abs(r), 1e-300))) - 6)
        dev = np.abs(hist[:, col] - r) / ulp
        ok = np.abs(r) > 1e-12
        assert dev[ok].max() <= 1.0, (name, col, float(dev[ok].max()), int(dev[ok].argmax()))
    info = d.info()
    print(f"{name}: {nsteps} steps by the product driver in {dt:.2f} s; wing uploads {info['wing_uploads']} (one per step and rotor)")
    assert info["wing_uploads"] == (nsteps + 1) * d.nr


def _vs_oracle(ctx, oracle, fx, nsteps, tol=1e-8):
    a = oracle.Case(fx)
    a.init()
    d, hist, gam = _run(ctx, fx, nsteps)
    worst = [abs(hist[0, 0] / a.force_nondim(0)[0] - 1.0) if a.force_nondim(0)[0] != 0 else 0.0, 0.0]
    for it in range(nsteps):
        a.step()
        fa, ga = a.force_nondim(0), a.rotor(0).vec(0)
        worst[0] = max(worst[0], abs(hist[it + 1, 0] / fa[0] - 1.0))
        worst[1] = max(worst[1], float(np.max(np.abs(gam[it] - ga)) / np.max(np.abs(ga))))
    assert max(worst) < tol, worst
    return worst


def _short_caradonna(fx):
    fx["config"]["nt"] = 40
    fx["geom"][0]["nNwake"] = 12


@pytest.mark.gpu
@pytest.mark.parametrize("name,nsteps,mutate", [
    ("simplewing", 40, None), ("tr1208", 30, None), ("caradonna", 30, _short_caradonna),
    ("katzNplotkin_AR04", 30, lambda fx: fx["config"].update(fdScheme=0)),
    ("elevateTest", 30, lambda fx: fx["config"].update(fdScheme=1)),
    ("caradonna", 20, lambda fx: (_short_caradonna(fx), fx["config"].update(fdScheme=2))),
    ("caradonna", 20, lambda fx: (_short_caradonna(fx), fx["config"].update(fdScheme=4))),
    ("caradonna", 20, lambda fx: (_short_caradonna(fx), fx["config"].update(fdScheme=5))),
    ("caradonna", 20, lambda fx: (_short_caradonna(fx), fx["config"].update(initWakeVelNt=8, wakeStrain=1, fdScheme=1),
                                  fx["geom"][0].update(initWakeVel=-3.0)))])
def test_product_driver_vs_cpu_driver(ctx, oracle, name, nsteps, mutate):
    """CL/CT and circulation histories against the oracle's CPU restatement of the reference driver (the checker): the
    remaining BASELINE configs and every time-marching scheme, within the north-star's 1e-8."""
    fx = _fixture(name)
    if mutate:
        mutate(fx)
    w = _vs_oracle(ctx, oracle, fx, nsteps)
    print(f"{name} fdScheme {fx['config'].get('fdScheme')}: {nsteps} steps, max rel err CL/CT {w[0]:.2e}, gamVec {w[1]:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("ntSub", [1, 3])
def test_sub_iterations_on_the_device(ctx, oracle, ntSub):
    """switches%ntSub > 0 (main.f90:522-615): every pass of the sub-iteration loop restarts velCP from its kinematic part
    (vlc_rotor_reset_velCP), re-evaluates the right-hand side of EVERY rotor with the other rotors' circulations of the
    previous pass, solves, and stops when gamVec no longer changes.  Wing + rotor, where the passes do change the solution."""
    from tests.test_oracle_case import two_body_case
    fx = two_body_case()
    fx["config"].update(ntSub=ntSub, ntSubInit=ntSub)
    a = oracle.Case(fx)
    a.init()
    from volcanor_b200.run_case import CaseDriver
    d = CaseDriver(fx, ctx)
    d.init()
    worst = 0.0
    for it in range(12):
        a.step()
        d.step()
        for ir in range(2):
            fa, fb = a.force_nondim(ir), d.force_nondim(ir)
            ga, gb = a.rotor(ir).vec(0), d.gamvec(ir)
            worst = max(worst, abs(fb[0] / fa[0] - 1.0), float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
    print(f"ntSub = {ntSub}: wing + rotor, 12 steps, max rel err {worst:.2e}")
    assert worst < 1e-8
    # and the sub-iterations matter: without them the histories differ by far more than the tolerance
    fx0 = two_body_case()
    b = oracle.Case(fx0)
    b.init()
    for it in range(12):
        b.step()
    assert abs(b.force_nondim(1)[0] / a.force_nondim(1)[0] - 1.0) > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("two_rotors", [False, True])
def test_nonuniform_streamwiseCoreVec_case_runs_on_the_dual_form(ctx, oracle, two_rotors):
    """A case file with a NON-UNIFORM streamwiseCoreVec (classdef.f90:2722-2726, :3841) and wake dissipation on: after the
    first rotor_dissipate_wake (vf(3)%rVc <- vf(1)%rVc, :4371-4372) the two copies of every interior streamwise edge of the
    wake carry different core radii.  The whole time loop then runs on the dual form of the lattice kernel
    (rotor_info: shared_active == 2) and reproduces the CPU driver's histories within the north-star's 1e-8; round 1 ran such
    a case on the flat enumeration.  With a second rotor whose cores ARE uniform the combined wake sweep sees two forms and
    goes to the flat enumeration for that launch (or_flags_kernel) -- same answer."""
    from tests.test_oracle_case import two_body_case
    if two_rotors:
        fx = two_body_case()
        g = fx["geom"][1]
    else:
        fx = _fixture("caradonna")
        _short_caradonna(fx)
        g = fx["geom"][0]
    g["streamwiseCoreVec"] = [0.03 + 0.002 * j for j in range(int(g["ns"]) + 1)]
    fx["config"]["wakeDissipation"] = 1
    a = oracle.Case(fx)
    a.init()
    from volcanor_b200.run_case import CaseDriver
    d = CaseDriver(fx, ctx)
    d.init()
    ir = 1 if two_rotors else 0
    worst, forms = 0.0, set()
    for it in range(14 if two_rotors else 25):
        a.step()
        d.step()
        for k in range(d.nr):
            fa, fb = a.force_nondim(k), d.force_nondim(k)
            ga, gb = a.rotor(k).vec(0), d.gamvec(k)
            worst = max(worst, abs(fb[0] / fa[0] - 1.0), float(np.max(np.abs(gb - ga)) / np.max(np.abs(ga))))
        forms.add(ctx.rotor_info(ir, True)["shared_active"])
    print(f"non-uniform streamwiseCoreVec ({'wing + rotor' if two_rotors else 'caradonna'}): max rel err {worst:.2e}, forms seen {sorted(forms)}")
    assert worst < 1e-8
    assert 2 in forms                                   # the dual form did the rotor's sweeps
    if two_rotors:
        assert ctx.rotor_info(0, True)["shared_active"] == 1    # the wing's wake: merged form


@pytest.mark.gpu
def test_product_driver_through_a_multi_gpu_handle(ctx, oracle):
    """The same driver, the same calls, a handle made by vlc_create_multi: bit-identical history under a fixed source split."""
    import torch

    import volcanor_b200 as vb
    fx = _fixture("elevateTest")
    n = 2
    g = vb.Context(devices=list(range(n)) if torch.cuda.device_count() >= n else [0] * n)
    try:
        out = []
        for c in (ctx, g):
            c.set_tuning(0, 2)
            try:
                out.append(_run(c, fx, 40)[1])
            finally:
                c.set_tuning(0, 0)
        assert np.array_equal(out[0], out[1])
    finally:
        g.close()


@pytest.mark.gpu
def test_run_case_cli_writes_the_reference_file_format(tmp_path):
    """python -m volcanor_b200.run_case on a case DIRECTORY written from the fixture (config.nml + geom01.nml): the file it
    writes parses back to the golden rows."""
    from volcanor_b200 import casefile, run_case
    fx = _fixture("katzNplotkin_AR04")
    case = tmp_path / "kp.case"
    case.mkdir()
    def nml(group, d):
        rows = []
        for k, v in d.items():
            if k == "grid" or isinstance(v, str) and k != "geometryFile":
                continue
            vals = v if isinstance(v, list) else [v]
            rows.append(f"{k} = " + ", ".join(repr(x) if not isinstance(x, str) else f"'{x}'" for x in vals))
        return f"&{group}\n" + "\n".join(rows) + "\n/\n"
    (case / "config.nml").write_text(nml("VERSION", {"fileFormatVersion": 0.5}) + nml("PARAMS", fx["config"]))
    (case / "geom01.nml").write_text(nml("GEOMPARAMS", {k: v for k, v in fx["geom"][0].items() if k != "grid"}))
    res = run_case.run(case, nt=20, quiet=True)
    txt = (Path(res["out"]) / "r01ForceNonDim.csv").read_text().splitlines()
    assert txt[0] == casefile.HEADER and len(txt) == 22
    ref = fx["ref_ForceNonDim"]["rows"]
    for it in (0, 1, 10, 20):
        got = float(txt[1 + it][5:20])
        assert abs(got - ref[it][1]) <= 1.01 * 10.0 ** (np.floor(np.log10(abs(ref[it][1]))) - 6), (it, got, ref[it][1])
