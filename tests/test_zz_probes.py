"""Velocity probes (`probes2file`, libPostprocess.f90:333-361): file formats on CPU, probe velocities
through the C ABI against the oracle on the GPU."""
import json
from pathlib import Path

import numpy as np
import pytest

from volcanor_b200 import probes

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_fortran_e15_7_matches_the_reference_format():
    # values as gfortran prints them with E15.7 (leading zero form), incl. rounding carried into the next decade
    assert probes.fortran_e(0.0) == "  0.0000000E+00"
    assert probes.fortran_e(1.0) == "  0.1000000E+01"
    assert probes.fortran_e(-1.25) == " -0.1250000E+01"
    assert probes.fortran_e(0.99999996) == "  0.1000000E+01"
    assert probes.fortran_e(9.87654321e-5) == "  0.9876543E-04"
    assert probes.fortran_e(-3.14159265e10) == " -0.3141593E+11"
    from oracle import casefile                                  # the formatter pinned to the golden force files
    rng = np.random.default_rng(5)
    assert probes.fortran_e(9.9999995e3) == "  0.9999999E+04"   # the double is 9999.99949999..., not a tie: no carry
    for x in np.concatenate([rng.standard_normal(500) * 10.0 ** rng.integers(-12, 12, 500), [0.5, 0.05, 9.9999995e3]]):
        assert probes.fortran_e(x) == casefile.fortran_e15_7(float(x)), x


def test_probes_in_and_csv_round_trip(tmp_path):
    (tmp_path / "probes.in").write_text("3\n0.0 1.0 2.0  0.0 0.0 0.0\n 1.5d0, -2.0, 0.25, 1.0, 0.0, -1.0\n-1 -1 -1 0 0 0\n")
    p, v = probes.read_probes(tmp_path / "probes.in")
    assert p.shape == (3, 3) and np.array_equal(p[1], [1.5, -2.0, 0.25]) and np.array_equal(v[1], [1.0, 0.0, -1.0])
    probes.write_probes(tmp_path / "probes00010.csv", v + 0.125, p)
    lines = (tmp_path / "probes00010.csv").read_text().splitlines()
    assert lines[0] == "".join(h.rjust(15) for h in "uvwxyz") and len(lines) == 4 and all(len(l) == 90 for l in lines)
    back = np.array([[float(l[15 * k:15 * (k + 1)]) for k in range(6)] for l in lines[1:]])
    assert np.allclose(back[:, :3], v + 0.125) and np.allclose(back[:, 3:], p)
    with pytest.raises(ValueError):
        (tmp_path / "bad.in").write_text("2\n0 0 0 0 0 0\n")
        probes.read_probes(tmp_path / "bad.in")


@pytest.mark.gpu
def test_probe_velocities_vs_oracle(oracle, tmp_path):
    """Two rotors (wing + rotor with far wake), moving probes: vel = probeVel + sum_ir [vind_bywing + vind_bywake]."""
    import volcanor_b200 as vb
    from tests.test_oracle_case import two_body_case
    from tests.test_zz_gpu_cp_stage import _define
    c = oracle.Case(two_body_case())
    c.init()
    for _ in range(10):
        c.step()
    ctx = vb.Context(0)
    rots = [c.rotor(ir) for ir in range(c.nr)]
    for ir, r in enumerate(rots):
        _define(ctx, r, ir)
    rng = np.random.default_rng(11)
    probe, probeVel, t = rng.uniform(-2.0, 4.0, (64, 3)), rng.uniform(-1.0, 1.0, (64, 3)), 0.37
    vel, loc = probes.probe_velocities(ctx, len(rots), probe, probeVel, t)
    ref, scale = probeVel.copy(), 0.0
    for r in rots:
        a, b = r.vind_points(0, loc), r.vind_points(1, loc)
        ref = (ref + a) + b
        scale = max(scale, float(np.abs(a).max()), float(np.abs(b).max()))
    assert np.array_equal(loc, probe + probeVel * t)
    assert np.max(np.abs(vel - ref)) < 1e-12 * 50.0 * scale, (float(np.max(np.abs(vel - ref))), scale)
    # inflow2file: -zAxis inflow at the section points of the rotor's blades (main.f90:771)
    rot = rots[1]
    secCP = np.stack([rot.sec(ib, "secCP", 3) for ib in range(rot.nb)])
    d = np.array([0.0, 0.0, -1.0])
    got = probes.inflow_velocities(ctx, len(rots), secCP, d)
    P = secCP.reshape(-1, 3)
    refi = np.zeros(P.shape[0])
    for r in rots:
        refi = refi + r.vind_points(0, P) @ d
        refi = refi - r.vind_points(3, P) @ d
        refi = refi + r.vind_points(1, P) @ d
    assert got.shape == (rot.nb, rot.ns)
    assert np.max(np.abs(got.ravel() - refi)) < 1e-12 * 50.0 * max(scale, float(np.abs(refi).max())), float(np.max(np.abs(got.ravel() - refi)))
    # row a6 complete: the chordwise vortices (classdef.f90:1398-1418), and bound + chordwise = the whole wing
    for ir, r in enumerate(rots):
        ch, bd, wh = ctx.rotor_vind_bywing_chordwiseVortices(ir, loc), ctx.rotor_vind_bywing_boundVortices(ir, loc), \
            ctx.rotor_vind_bywing(ir, loc)
        s6 = 50.0 * max(float(np.abs(r.vind_points(4, loc)).max()), float(np.abs(r.vind_points(3, loc)).max()))
        assert np.max(np.abs(ch - r.vind_points(4, loc))) < 1e-12 * s6
        assert np.max(np.abs(ch + bd - wh)) < 1e-12 * s6
    out = probes.probes2file(ctx, len(rots), tmp_path, "00010", probe, probeVel, t)
    assert out.name == "probes00010.csv" and len(out.read_text().splitlines()) == 65
    ctx.close()
