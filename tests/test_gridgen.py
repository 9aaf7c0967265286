"""The Eulerian grid tool: program gridgen (src/gridgen.f90) = the same pair kernel on the cell
centres of a Cartesian grid.  CPU: the oracle's restatement against a direct numpy evaluation of the file's formulas;
GPU: vlc_gridgen against the oracle."""
import numpy as np
import pytest

from tests.helpers import lattice_to_rotor
from volcanor_b200 import synth


def _filament_file(oracle, seed=0):
    """Record sets as filaments2file would write them: wing rings, near-wake rings, TE filaments, far filaments."""
    rng = np.random.default_rng(seed)
    from tests import refgeom
    wing = refgeom.flat_wing_records(3, 5, 1.0, 4.0, 0.1, [-10.0, 0, 0], 0.01, 0.04)
    wing[:, :, 48] = rng.uniform(-1, 1, size=(5, 3))
    vrWing = wing[:, :, :50].reshape(-1, 50).copy()
    lat = synth._helix_lattice(rng, np.array([0.0, 0.0, -0.3]), 1.0, 0.1, 5, 9, 6, psi0=0.2, sense=1.0)
    lat.gam[1, 2] = 0.0
    lat.gam[2, 3] = 1e-17           # gridgen applies no |gam| > eps rule
    ro = oracle.Rotor(1, 3, 5, 9, 6)
    lattice_to_rotor(lat, ro, 0)
    vrN = ro.waN(0).reshape(-1, 50).copy()
    te = ro.waN(0)[:, -1, 12:24].copy()                       # vf(2) of the last row
    gte = -ro.waN(0)[:, -1, 48].copy()
    vfF = ro.waF(0)[:, :12].copy()
    gF = ro.waF(0)[:, 12].copy()
    return vrWing, vrN, te, gte, vfF, gF


def test_oracle_gridgen_matches_the_file_formulas(oracle):
    vrWing, vrN, te, gte, vfF, gF = _filament_file(oracle)
    nx, ny, nz = 7, 5, 4
    lo, hi, vel = np.array([-2.0, -1.5, -1.0]), np.array([2.0, 1.5, 0.5]), np.array([3.0, 0.0, -1.0])
    gc, vc = oracle.gridgen(nx, ny, nz, lo, hi, vel, vrWing, vrN, te, gte, vfF, gF)
    xs = [np.arange(n) * ((h - l) / (n - 1)) + l for n, l, h in zip((nx, ny, nz), lo, hi)]
    mid = [0.5 * (x[:-1] + x[1:]) for x in xs]
    G = np.stack(np.meshgrid(mid[2], mid[1], mid[0], indexing="ij")[::-1], axis=-1)
    assert gc.shape == (nz - 1, ny - 1, nx - 1, 3) and np.max(np.abs(gc - G)) < 1e-15
    # direct evaluation: last wing ring only (gridgen.f90:122 assigns), all wake rings, TE and far filaments
    P = gc.reshape(-1, 3)

    def rings(rec):
        p1 = np.concatenate([rec[:, 12 * f:12 * f + 3] for f in range(4)])
        p2 = np.concatenate([rec[:, 12 * f + 3:12 * f + 6] for f in range(4)])
        rvc = np.concatenate([rec[:, 12 * f + 9] for f in range(4)])
        return p1, p2, rvc, np.tile(rec[:, 48], 4)

    parts = [rings(vrWing[-1:]), rings(vrN), (te[:, 0:3], te[:, 3:6], te[:, 9], gte), (vfF[:, 0:3], vfF[:, 3:6], vfF[:, 9], gF)]
    p1, p2, rvc, gam = (np.concatenate([p[k] for p in parts]) for k in range(4))
    V = oracle.vind_flat(p1, p2, rvc, gam, None, P) + vel
    assert np.max(np.abs(vc.reshape(-1, 3) - V)) < 1e-13 * np.max(np.abs(V))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(7, 5, 4), (2, 2, 2), (33, 18, 9)])
def test_vlc_gridgen_vs_oracle(ctx, oracle, shape):
    vrWing, vrN, te, gte, vfF, gF = _filament_file(oracle, seed=shape[0])
    nx, ny, nz = shape
    lo, hi, vel = np.array([-2.0, -1.5, -1.0]), np.array([2.0, 1.5, 0.5]), np.array([3.0, 0.0, -1.0])
    gc_o, vc_o = oracle.gridgen(nx, ny, nz, lo, hi, vel, vrWing, vrN, te, gte, vfF, gF)
    gc, vc = ctx.gridgen(nx, ny, nz, lo, hi, vel, vrWing, vrN, te, gte, vfF, gF)
    assert np.array_equal(gc, gc_o)                                   # same arithmetic for the cell centres
    scale = np.max(np.abs(vc_o - vel))
    assert np.max(np.abs(vc - vc_o)) < 1e-12 * max(scale, 1.0) * 50
    # empty file: free stream only
    z50, z12, z1 = np.zeros((0, 50)), np.zeros((0, 12)), np.zeros(0)
    _, v0 = ctx.gridgen(nx, ny, nz, lo, hi, vel, z50, z50, z12, z1, z12, z1)
    assert np.all(v0 == vel)


@pytest.mark.gpu
def test_vlc_gridgen_rejects_inverted_box(ctx):
    import volcanor_b200 as vb
    z50, z12, z1 = np.zeros((0, 50)), np.zeros((0, 12)), np.zeros(0)
    with pytest.raises(vb.VlcError):
        ctx.gridgen(3, 3, 3, [1.0, 0, 0], [0.0, 1, 1], [0, 0, 0], z50, z50, z12, z1, z12, z1)   # gridgen.f90:43-45


@pytest.mark.gpu
def test_vlc_gridgen_slices_cover_the_full_call(ctx, oracle):
    """vlc_gridgen_slice (one process per GPU takes a contiguous slice of the cell list): same cell centres bit for bit,
    velocities as the full call up to the summation order of the source split (1e-13 of the velocity scale)."""
    import volcanor_b200 as vb
    vrWing, vrN, te, gte, vfF, gF = _filament_file(oracle, seed=9)
    nx, ny, nz = 12, 9, 6
    lo, hi, vel = np.array([-2.0, -1.5, -1.0]), np.array([2.0, 1.5, 0.5]), np.array([3.0, 0.0, -1.0])
    args = (nx, ny, nz, lo, hi, vel, vrWing, vrN, te, gte, vfF, gF)
    gc, vc = ctx.gridgen(*args)
    m = (nx - 1) * (ny - 1) * (nz - 1)
    scale = max(float(np.max(np.abs(vc - vel))), 1.0)
    for world in (2, 3):
        per = (m + world - 1) // world
        parts = [ctx.gridgen_slice(*args, r * per, min(per, m - r * per)) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), gc.reshape(-1, 3))
        assert np.max(np.abs(np.concatenate([p[1] for p in parts]) - vc.reshape(-1, 3))) < 1e-13 * 50 * scale
    g0, v0 = ctx.gridgen_slice(*args, 7, 0)
    assert g0.shape == (0, 3) and v0.shape == (0, 3)
    with pytest.raises(vb.VlcError, match="cell slice"):
        ctx.gridgen_slice(*args, m - 3, 4)
