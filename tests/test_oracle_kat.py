"""Pins the CPU oracle against the reference's own known-answer tests (CPU only)."""
import numpy as np
import pytest

from tests import refgeom


def _oracle_aic(oracle, rec):
    ns, nc, _ = rec.shape
    r = oracle.Rotor(1, nc, ns, 2, 0)
    r.wiP(0)[...] = rec
    assert r.calcAIC() == 0
    return r, r.AIC().copy()


def test_struct_layout_matches_reference_records(oracle):
    import ctypes
    lib = oracle.load()
    # 96 / 400 / 832 / 104 bytes: classdef.f90:57-79, 81-104, 106-179, 198-220
    r = oracle.Rotor(1, 2, 3, 4, 5)
    assert r.wiP(0).shape == (3, 2, 104) and r.waN(0).shape == (3, 4, 50) and r.waF(0).shape == (5, 13)


def test_aic_wing1x3_15_digits(oracle):
    """tests/wing1x3_test.f90:83-85 (reference tolerance 1e-6; the literals carry 15 digits)."""
    _, A = _oracle_aic(oracle, refgeom.wing1x3())
    assert np.max(np.abs(A - refgeom.AIC_WING1X3)) < 1e-6          # the reference's own assertion
    assert np.max(np.abs(A / refgeom.AIC_WING1X3 - 1.0)) < 5e-13   # all printed digits


def test_aic_wing1x2(oracle):
    """tests/wing1x2_test.f90:164-168"""
    _, A = _oracle_aic(oracle, refgeom.wing1x2())
    assert np.max(np.abs(A - refgeom.AIC_WING1X2)) < 1e-6
    assert np.max(np.abs(A / refgeom.AIC_WING1X2 - 1.0)) < 2e-7    # 8 printed digits


def test_gamvec_wing1x3(oracle):
    """tests/wing1x3_test.f90:104-136: RHS = -velCP.nCap after 7 deg pitch; gamVec = AIC_inv*RHS."""
    r, _ = _oracle_aic(oracle, refgeom.wing1x3())
    th = np.deg2rad(7.0)
    ncap = np.array([np.sin(th), 0.0, np.cos(th)])  # asserted by the reference test at :109-111
    rhs = -np.full(3, np.dot([6.0, 0.0, 0.0], ncap))
    g = r.AIC(inverse=True) @ rhs
    assert np.max(np.abs(g - np.array([-0.240131, -0.249833, -0.240131]))) < 1e-6


def test_inv2_and_matmulAX_libmath_kat(oracle):
    """tests/libMath_test.f90:64-95"""
    lib = oracle.load()
    A = np.array([[1.0, 1.0, 2.0], [1.0, 2.0, 1.0], [2.0, 1.0, 1.0]], order="F")
    Ainv = np.empty((3, 3), order="F")
    assert lib.orc_inv2(3, A.ctypes.data, Ainv.ctypes.data) == 0
    assert np.allclose(Ainv @ A, np.eye(3), atol=1e-14)
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 52, 104):
        M = np.asfortranarray(rng.normal(size=(n, n)) + n * np.eye(n))
        Mi = np.empty((n, n), order="F")
        assert lib.orc_inv2(n, M.ctypes.data, Mi.ctypes.data) == 0
        assert np.allclose(Mi, np.linalg.inv(M), rtol=1e-10, atol=1e-12)
        x = rng.normal(size=n)
        y = np.empty(n)
        lib.orc_matmulAX(n, n, M.ctypes.data, x.ctypes.data, y.ctypes.data)
        assert np.allclose(y, M @ x, rtol=1e-13, atol=1e-13)
    S = np.zeros((3, 3), order="F")
    assert lib.orc_inv2(3, S.ctypes.data, Ainv.ctypes.data) != 0   # 'Matrix is numerically singular!'


def test_vf_vind_guards(oracle):
    """classdef.f90:498: zero on the filament axis / end points; rVc = 0 is legal (:3504-3506)."""
    lib = oracle.load()
    f = np.zeros(12)
    f[0:3] = [0, 0, 0]
    f[3:6] = [1, 0, 0]
    f[9] = 0.0
    v = np.empty(3)
    for P in ([0, 0, 0], [1, 0, 0], [0.5, 0, 0], [3, 0, 0]):
        P = np.array(P, dtype=float)
        lib.orc_vf_vind(f.ctypes.data, P.ctypes.data, v.ctypes.data)
        assert np.all(v == 0.0)
    P = np.array([0.5, 1.0, 0.0])
    lib.orc_vf_vind(f.ctypes.data, P.ctypes.data, v.ctypes.data)
    # finite straight segment, h = 1: v = (cos a1 - cos a2)/(4 pi h) along +z
    expect = (2 * 0.5 / np.sqrt(1.25)) / (4 * np.pi)
    assert abs(v[2] - expect) < 1e-15 and v[0] == 0 and v[1] == 0


def test_flat_equals_structured(oracle):
    """The flat helper used for GPU parity and the structured source loops agree to rounding."""
    from volcanor_b200 import synth
    lat = synth.multirotor(3000, seed=3, n_rotor=1, nb=1, S=6, F=5, with_wing=False)[0]
    R, S, F = lat.R, lat.S, lat.F
    r = oracle.Rotor(1, 1, S, R, F)
    from tests.helpers import lattice_to_rotor
    lattice_to_rotor(lat, r, 0)
    r.set_rows(1, 1)
    P = np.random.default_rng(1).uniform(-1.5, 1.5, size=(50, 3))
    Vs = r.vind_points(1, P)
    p1, p2, rvc, gam, flag = lat.flatten()
    Vf = oracle.vind_flat(p1, p2, rvc, gam, flag, P)
    _, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, P)
    assert np.max(np.abs(Vs - Vf) / Vabs.max()) < 1e-14


def test_bound_plus_chordwise_vortices_equal_the_whole_wing(oracle):
    """classdef.f90:1376-1418: vind_bywing_boundVortices (filaments 2, 4 minus the trailing-edge row) and
    vind_bywing_chordwiseVortices (filaments 1, 3 plus the trailing-edge row) partition the wing's rings, so their sum is
    vind_bywing up to the order of summation."""
    import json
    from pathlib import Path
    fx = json.loads((Path(__file__).resolve().parent / "golden" / "caradonna.json").read_text())
    fx["geom"][0]["nNwake"] = 6
    c = oracle.Case(fx)
    c.init()
    for _ in range(3):
        c.step()
    r = c.rotor(0)
    P = np.random.default_rng(2).uniform(-1.5, 1.5, (40, 3))
    whole, bound, chord = r.vind_points(0, P), r.vind_points(3, P), r.vind_points(4, P)
    assert np.max(np.abs(chord)) > 0 and np.max(np.abs(bound)) > 0
    assert np.max(np.abs(bound + chord - whole)) < 1e-13 * max(np.max(np.abs(bound)), np.max(np.abs(chord)))
