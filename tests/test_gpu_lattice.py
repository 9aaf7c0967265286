"""GPU parity of the device-resident lattice path (tier 3) and its shared-node sweep kernel (bs_lattice.cuh) against
the CPU oracle evaluating the REFERENCE's enumeration (ring by ring, 4 filaments per ring, classdef.f90:1450-1469).

The shared-node kernel regroups that sum (one evaluation per lattice node and per unique edge); results must agree
within the per-call tolerance 1e-12 of the velocity scale, for targets that coincide with lattice nodes (the wake
nodes themselves: every target lies on up to 4 edges, all of which must be skipped exactly like classdef.f90:498),
for skipped rings (|gam| <= eps), with and without a far wake, and when the two copies of a shared edge carry
different core radii, in which case the device falls back to the flat enumeration by itself.
"""
import numpy as np
import pytest

from tests.helpers import scaled_err
from volcanor_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _upload(ctx, lats, set_=3):
    import torch
    keep = []
    for i, l in enumerate(lats):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        nodes, gam, rvc4 = t(l.nodes), t(l.gam), t(l.rvc4)
        far = t(l.far_nodes) if l.F > 0 else None
        gF = t(l.gamF) if l.F > 0 else None
        rF = t(l.rvcF) if l.F > 0 else None
        ctx.pack_lattice_dev(set_, i > 0, l.R, l.S, nodes, gam, rvc4, l.F, far, gF, rF)
        keep += [nodes, gam, rvc4, far, gF, rF]
    ctx.sync()
    return keep


def _sweep(ctx, P, set_=3):
    import torch
    dP = torch.from_numpy(np.ascontiguousarray(P)).cuda()
    dV = torch.full_like(dP, float("nan"))
    ctx.vind_dev(set_, P.shape[0], dP, dV)
    ctx.sync()
    return dV.cpu().numpy()


def _check(ctx, oracle, lats, P, tol=TOL):
    keep = _upload(ctx, lats)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    assert ctx.num_sources(3) == rvc.size                   # reference enumeration count
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, P)
    Vl, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, P)
    ctx.set_shared_nodes(True)
    Vs = _sweep(ctx, P)
    ctx.set_shared_nodes(False)
    Vf = _sweep(ctx, P)
    ctx.set_shared_nodes(True)
    assert np.all(np.isfinite(Vs)) and np.all(np.isfinite(Vf))
    es, ef = scaled_err(Vs, Vo, Vabs), scaled_err(Vf, Vo, Vabs)
    assert es < tol and ef < tol, (es, ef)
    # neither GPU form is further from the long-double sum than the reference-order double sum is
    assert scaled_err(Vs, Vl, Vabs) < 5 * max(scaled_err(Vo, Vl, Vabs), 1e-15)
    return Vs, Vf, Vo, Vabs


@pytest.mark.parametrize("kw", [dict(n=3000, n_rotor=1, nb=1, S=4, F=0, with_wing=False),
                                dict(n=6000, n_rotor=1, nb=2, S=8, F=8, with_wing=True),
                                dict(n=20000, n_rotor=4, nb=2, S=8, F=16, with_wing=True)])
def test_wake_nodes_as_targets(ctx, oracle, kw):
    """The convection sweep itself: targets are the lattice's own nodes (libCommon.f90:133-145)."""
    kw = dict(kw)
    lats = synth.multirotor(kw.pop("n"), seed=11, **kw)
    _check(ctx, oracle, lats, synth.targets_all(lats))


def test_off_lattice_targets_and_source_splits(ctx, oracle):
    lats = synth.multirotor(8000, seed=5, n_rotor=2, nb=2, S=6, F=5)
    P = np.random.default_rng(2).uniform(-4, 4, size=(333, 3))
    for nsplit in (0, 1, 3, 16):
        ctx.set_tuning(0, nsplit)
        try:
            _check(ctx, oracle, lats, P)
        finally:
            ctx.set_tuning(0, 0)


@pytest.mark.parametrize("W", [1, 2, 3, 4])
@pytest.mark.parametrize("T", [1, 2, 3])
def test_strip_widths_and_targets_per_thread(ctx, oracle, W, T):
    """Launch shapes of the shared-node kernel (strip width W: ns not a multiple of W leaves a partly empty last strip;
    T targets per thread) only change speed and summation order -- for lattices and for uploaded rotor records."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    ctx.set_lattice_tuning(W, T)
    try:
        lats = synth.multirotor(7000, seed=21, n_rotor=2, nb=2, S=7, F=6, with_wing=True)
        _check(ctx, oracle, lats, synth.targets_all(lats))
        assert ctx.set_info(3)["strip_width"] == W
        ro = _make_rotor_pair(ctx, oracle, seed=6, ns=5, nNwake=7, nFwake=4, rowNear=2, rowFar=2)
        P = np.random.default_rng(3).uniform(-1.5, 1.5, size=(300, 3))
        s = _tol_scale(ro, P) * 50
        assert np.max(np.abs(ctx.rotor_vind_bywake(0, P) - ro.vind_points(1, P))) < TOL * s
        assert np.max(np.abs(ctx.rotor_vind(0, P, True) - ro.vind_points(2, P, True))) < TOL * s
    finally:
        ctx.set_lattice_tuning(0, 0)


def test_far_field_targets_per_target_tolerance(ctx, oracle):
    """Targets 10 / 100 / 1000 rotor radii away from the wake, mixed into a batch with wake-node targets: their own
    velocity scale sum|terms| is 1e-2 ... 1e-6 of the batch maximum, so only the PER-TARGET measure
    (tests/helpers.py:scaled_err) can see an error there -- a dropped edge or a wrong merged strength on a
    low-influence target (round-1 review, item 6).  The pair formula r0.(r1/|r1| - r2/|r2|) loses log2(r/L) bits to
    cancellation IN THE REFERENCE TOO (classdef.f90:499): its own double sum is 5e-15 from the exact sum at 100 R and
    2e-11 at 1e6 R (measured with the long-double oracle).  Hence: 1e-12 per target out to 1000 R (measured r02b: lattice
    kernel 8e-16 / 8e-15 / 7e-14 at 10 / 100 / 1000 R, flat kernel 6e-15 / 4e-14 / 3e-13), and for the 1e6 R group -- where
    the reference itself is not reproducible to 1e-12 -- a bar relative to the reference-order double sum's own distance
    from the exact sum: factor 5 for the lattice kernel (measured 1.1), factor 50 for the flat kernel (measured 25: it
    uses the per-source r0 = p2 - p1 and G r0.r2 = G r0.r1 - G|r0|^2, exact in exact arithmetic, while the reference forms
    r0 = r1 - r2 from the two ROUNDED differences P - p, which at |P| = 1e6 carry an absolute error of 1e-10 each)."""
    from tests.helpers import scaled_err_batch
    lats = synth.multirotor(20000, seed=31, n_rotor=4, nb=2, S=8, F=16, with_wing=True)
    rng = np.random.default_rng(5)
    d = rng.normal(size=(600, 3))
    far = 100.0 * d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(0.8, 1.5, size=(600, 1))
    groups = [("wake nodes", synth.targets_all(lats)[::7]), ("10 R", 0.1 * far[:200]), ("100 R", far),
              ("1000 R", 10.0 * far[:200])]
    P = np.concatenate([g[1] for g in groups])
    Vs, Vf, Vo, Vabs = _check(ctx, oracle, lats, P)          # asserts 1e-12 PER TARGET on the whole batch, both kernels
    off = 0
    for name, Pg in groups:
        sl = slice(off, off + Pg.shape[0])
        off += Pg.shape[0]
        print(f"{name:>10}: scale {Vabs[sl].max():.2e}  per-target error lattice {scaled_err(Vs[sl], Vo[sl], Vabs[sl]):.2e} "
              f"flat {scaled_err(Vf[sl], Vo[sl], Vabs[sl]):.2e}  (batch-scaled form: "
              f"{scaled_err_batch(Vs[sl], Vo[sl], Vabs):.1e})")
    assert Vabs[-200:].max() < 1e-5 * Vabs[:groups[0][1].shape[0]].max()   # the far targets ARE invisible to the batch form
    # 1e6 R: cancellation of ~26 bits in the reference's own formula
    Px = 1e4 * far[:100]
    keep = _upload(ctx, lats)
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, Px)
    Vl, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, Px)
    ref_err = scaled_err(Vo, Vl, Vabs)
    for shared in (True, False):
        ctx.set_shared_nodes(shared)
        try:
            e = scaled_err(_sweep(ctx, Px), Vl, Vabs)
        finally:
            ctx.set_shared_nodes(True)
        print(f"     1e6 R: vs exact sum: {'lattice' if shared else 'flat'} kernel {e:.2e}, reference-order double sum {ref_err:.2e}")
        assert e < (5 if shared else 50) * ref_err


@pytest.mark.parametrize("R,S,F", [(1, 1, 0), (1, 1, 3), (2, 1, 0), (1, 5, 2), (130, 2, 1), (3, 70, 0)])
def test_degenerate_lattice_shapes(ctx, oracle, R, S, F):
    rng = np.random.Generator(np.random.PCG64(R * 100 + S))
    lat = synth._helix_lattice(rng, np.zeros(3), 1.0, 0.1, S, R, F, psi0=0.3, sense=1.0)
    P = np.concatenate([lat.targets(), rng.uniform(-2, 2, size=(17, 3))])
    _check(ctx, oracle, [lat], P)


def test_skipped_rings_and_zero_strength(ctx, oracle):
    """classdef.f90:1452: rings with |gam| <= eps contribute nothing; merged edge strengths use Gamma' = 0 for them."""
    lats = synth.multirotor(5000, seed=9, n_rotor=1, nb=2, S=6, F=4, with_wing=False)
    rng = np.random.default_rng(0)
    for l in lats:
        l.gam[rng.uniform(size=l.gam.shape) < 0.2] = 0.0
        l.gam[rng.uniform(size=l.gam.shape) < 0.1] = 1e-17
    _check(ctx, oracle, lats, synth.targets_all(lats))
    for l in lats:
        l.gam[...] = 1e-17
        l.gamF[...] = 0.0
    keep = _upload(ctx, lats)
    V = _sweep(ctx, synth.targets_all(lats))
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, synth.targets_all(lats))
    # only the horseshoe correction (no gam rule, classdef.f90:1460-1463) is left: O(1e-17)
    assert np.max(np.abs(V - Vo)) < 1e-28


def test_unmergeable_core_radii_fall_back_to_flat_enumeration(ctx, oracle):
    """Non-uniform streamwise cores: vf3 of ring (r, j) and vf1 of ring (r, j+1) are the same edge with different
    rVc (the reference keeps both).  The pack kernel detects it and the device runs the flat kernel."""
    lats = synth.multirotor(6000, seed=4, n_rotor=1, nb=2, S=6, F=4, with_wing=False)
    for l in lats:
        l.rvc4[:, :, 2] *= 1.0 + 0.05 * np.arange(l.S)[:, None]      # vf3 differs from the neighbour's vf1
    P = synth.targets_all(lats)
    Vs, Vf, Vo, Vabs = _check(ctx, oracle, lats, P)
    assert np.array_equal(Vs, Vf)          # same kernel, same records, same split -> bitwise equal
    # and a mergeable set packed afterwards into the same slot uses the lattice kernel again
    lats2 = synth.multirotor(6000, seed=4, n_rotor=1, nb=2, S=6, F=4, with_wing=False)
    Vs2, Vf2, _, Vabs2 = _check(ctx, oracle, lats2, P)
    assert not np.array_equal(Vs2, Vf2) and scaled_err(Vs2, Vf2, Vabs2) < TOL


@pytest.mark.parametrize("ns,nNwake,W", [(8, 12, 0), (6, 9, 0), (5, 7, 0), (7, 10, 2)])
def test_nonuniform_streamwise_cores_use_the_dual_form(ctx, oracle, ns, nNwake, W):
    """Rotor records shed with a non-uniform streamwiseCoreVec (classdef.f90:3841; dissipate_wake keeps vf(3)%rVc = vf(1)%rVc
    of the same ring, :4371-4372): the two copies of every interior streamwise edge differ.  Round 1 abandoned the lattice
    kernel for the whole set; now check_rings_kernel selects the DUAL form (rotor_info: shared_active == 2): same nodes, same
    spanwise edges, 33-instruction streamwise edges with both core radii.  Per-call parity against the oracle's ring-by-ring
    sum, for every source loop that contains the wake, with tail strips (ns = 5, 6, 7) and a forced strip width."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    ctx.rotors_clear()
    ctx.set_lattice_tuning(W, 0)
    try:
        ro = _make_rotor_pair(ctx, oracle, seed=13, ns=ns, nNwake=nNwake, nFwake=4, rowNear=2, rowFar=2)
        for pred in (False, True):
            for ib in range(ro.nb):
                w = ro.waN(ib, pred)
                for j in range(ns):
                    w[j, :, 9] *= 1.0 + 0.25 * j             # vf(1)%rVc by column
                    w[j, :, 24 + 9] = w[j, :, 9]              # vf(3)%rVc = vf(1)%rVc of the same ring
                ctx.rotor_put_nwake(0, ib, w, pred)
        assert ctx.rotor_info(0)["shared_active"] == 2 and ctx.rotor_info(0, True)["shared_active"] == 2
        rng = np.random.default_rng(3)
        nodes = ro.waN(0)[:, 1:, 12:15].reshape(-1, 3)
        P = np.concatenate([rng.uniform(-1.5, 1.5, size=(300, 3)), nodes,
                            0.5 * (ro.waN(1)[:, 1:, 0:3] + ro.waN(1)[:, 1:, 12:15]).reshape(-1, 3)])
        s = _tol_scale(ro, P) * 50
        got = ctx.rotor_vind_bywake(0, P)
        assert np.max(np.abs(got - ro.vind_points(1, P))) < TOL * s
        assert np.max(np.abs(ctx.rotor_vind_bywake(0, P, True) - ro.vind_points(1, P, True))) < TOL * s
        assert np.max(np.abs(ctx.rotor_vind(0, P) - ro.vind_points(2, P))) < TOL * s
        # the flat enumeration of the same records agrees to rounding, and is a different summation (not bitwise)
        ctx.set_shared_nodes(False)
        try:
            flat = ctx.rotor_vind_bywake(0, P)
        finally:
            ctx.set_shared_nodes(True)
        assert np.max(np.abs(flat - got)) < TOL * s and not np.array_equal(flat, got)
        # one spanwise copy out of step (the transient between shiftwake and dissipate_wake): the whole set goes flat
        w = ro.waN(0).copy()
        w[2, 3, 36 + 9] *= 1.5
        ctx.rotor_put_nwake(0, 0, w)
        assert ctx.rotor_info(0)["shared_active"] == 0
        ro.waN(0)[...] = w
        assert np.max(np.abs(ctx.rotor_vind_bywake(0, P) - ro.vind_points(1, P))) < TOL * s
    finally:
        ctx.set_lattice_tuning(0, 0)
        ctx.rotors_clear()


def test_lattice_state_kernels_vs_oracle(ctx, oracle):
    """dissipate (classdef.f90:4364-4393), convect (:1531), AB2/AM2 (main.f90:1032-1034, :1094-1096) on device arrays."""
    import torch
    rng = np.random.default_rng(1)
    R, S = 7, 5
    rvc4 = rng.uniform(0.01, 0.05, size=(S, R, 4))
    gam = rng.uniform(-1, 1, size=(S, R))
    d_r, d_g = torch.from_numpy(rvc4.copy()).cuda(), torch.from_numpy(gam.copy()).cuda()
    a, nu, k, dt = 5.0, 1.8e-5, 0.3, 2e-3
    ctx.dissipate_lattice_dev(R, S, d_r, d_g, a, nu, k, dt)
    ctx.sync()
    exp = rvc4.copy()
    g2 = 4.0 * 1.2564 * a * nu * dt
    exp[:, :, 0] = np.sqrt(rvc4[:, :, 0] ** 2 + g2)
    exp[:, :, 2] = exp[:, :, 0]
    exp[:, :, 1] = np.sqrt(rvc4[:, :, 1] ** 2 + g2)
    exp[:, 1:, 3] = exp[:, :-1, 1]
    assert np.max(np.abs(d_r.cpu().numpy() - exp)) < 1e-17
    assert np.max(np.abs(d_g.cpu().numpy() - gam * np.exp(-k * dt))) < 1e-16
    x, v, v1 = rng.normal(size=(50, 3)), rng.normal(size=(50, 3)), rng.normal(size=(50, 3))
    dx, dv, dv1 = (torch.from_numpy(t.copy()).cuda() for t in (x, v, v1))
    out = torch.empty_like(dv)
    ctx.convect_dev(50, dx, dv, dt)
    ctx.ab2_dev(50, dv, dv1, out)
    ctx.sync()
    assert np.array_equal(dx.cpu().numpy(), x + v * dt)
    assert np.array_equal(out.cpu().numpy(), 0.5 * (3.0 * v - v1))
    ctx.am2_dev(50, dv, dv1, out)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), (v + v1) * 0.5)


# ------------------------------------------------------------------ tier 2: reference records -> shared-node form

def test_rotor_records_shared_and_flat_forms_agree_with_oracle(ctx, oracle):
    """vlc_rotor_vind_bywake / vlc_rotor_vind on uploaded waN records use the shared-node form when the records
    describe a lattice (they do after blade_wake_continuity); both forms must match the oracle."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    P = np.random.default_rng(3).uniform(-1.5, 1.5, size=(300, 3))
    for rows in ((1, 1), (3, 2), (7, 5)):
        ro = _make_rotor_pair(ctx, oracle, seed=6, nNwake=7, nFwake=4, rowNear=rows[0], rowFar=rows[1])
        s = _tol_scale(ro, P) * 50
        res = {}
        for shared in (True, False):
            ctx.set_shared_nodes(shared)
            try:
                res[shared] = (ctx.rotor_vind_bywake(0, P), ctx.rotor_vind_bywake(0, P, True), ctx.rotor_vind(0, P))
            finally:
                ctx.set_shared_nodes(True)
            assert np.max(np.abs(res[shared][0] - ro.vind_points(1, P))) < TOL * s
            assert np.max(np.abs(res[shared][1] - ro.vind_points(1, P, True))) < TOL * s
            assert np.max(np.abs(res[shared][2] - ro.vind_points(2, P))) < TOL * s
        if rows[0] < 7:
            assert not np.array_equal(res[True][0], res[False][0])     # really two different kernels


def test_rotor_records_that_are_not_a_lattice_fall_back(ctx, oracle):
    """One corner copy moved: the records no longer describe a lattice, the reference (and the oracle) still sum
    them filament by filament; the device detects the mismatch and uses the flat enumeration."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    ro = _make_rotor_pair(ctx, oracle, seed=12, nNwake=6, nFwake=3, rowNear=2, rowFar=2)
    w = ro.waN(0)
    w[2, 3, 12 * 1 + 0] += 1e-3          # vf(2)%fc(1,1) of ring (row 4, col 3) only: breaks continuity with vf(1)%fc(:,2)
    ctx.rotor_put_nwake(0, 0, w)
    P = np.random.default_rng(4).uniform(-1.5, 1.5, size=(100, 3))
    s = _tol_scale(ro, P) * 50
    Va = ctx.rotor_vind_bywake(0, P)
    ctx.set_shared_nodes(False)
    try:
        Vb = ctx.rotor_vind_bywake(0, P)
    finally:
        ctx.set_shared_nodes(True)
    assert np.max(np.abs(Va - ro.vind_points(1, P))) < TOL * s
    assert np.array_equal(Va, Vb)


def test_far_wake_state_kernels(ctx):
    """Far-wake core growth / decay (classdef.f90:4397-4404, Fwake_decay :975-980) and strain (:4410-4422, :505-521)
    on flat device arrays."""
    import torch
    rng = np.random.default_rng(5)
    n = 1000
    rvc, gam = rng.uniform(0.01, 0.1, n), rng.normal(size=n)
    d_r, d_g = torch.from_numpy(rvc.copy()).cuda(), torch.from_numpy(gam.copy()).cuda()
    a, nu, k, dt = 5000.0, 1.81e-5, 0.2, 2.8e-3
    ctx.dissipate_dev(n, d_r, n, d_g, a, nu, k, dt)
    ctx.sync()
    assert np.array_equal(d_r.cpu().numpy(), np.sqrt(rvc * rvc + 4.0 * 1.2564 * a * nu * dt))
    assert np.max(np.abs(d_g.cpu().numpy() / (gam * np.exp(-k * dt)) - 1.0)) < 4e-16
    p1, p2 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    l0, rvc0 = rng.uniform(0.5, 2.0, n), rng.uniform(0.01, 0.1, n)
    out = torch.zeros(n, dtype=torch.float64).cuda()
    ctx.strain_dev(n, torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda(), torch.from_numpy(l0).cuda(),
                   torch.from_numpy(rvc0).cuda(), out)
    ctx.sync()
    d = p1 - p2
    lc = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    assert np.max(np.abs(out.cpu().numpy() / (rvc0 * np.sqrt(l0 / lc)) - 1.0)) < 4e-16


@pytest.mark.parametrize("rows,cols", [(1, 3), (2, 2), (7, 6), (40, 1)])
def test_vel_order2_vs_oracle(ctx, oracle, rows, cols):
    """fdScheme 1 averaging (libCommon.f90:213-258) bit for bit against the oracle's restatement."""
    import torch
    rng = np.random.default_rng(rows * 10 + cols)
    vn, vp = rng.normal(size=(cols, rows, 3)), rng.normal(size=(cols, rows, 3))
    ref = np.empty_like(vn)
    lib = oracle.load()
    lib.orc_vel_order2_Nwake(vn.ctypes.data, vp.ctypes.data, rows, cols, ref.ctypes.data)
    out = torch.zeros(cols, rows, 3, dtype=torch.float64).cuda()
    ctx.vel_order2_dev(rows, cols, torch.from_numpy(vn).cuda(), torch.from_numpy(vp).cuda(), out)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), ref)


def test_full_size_workload_properties_and_sampled_oracle(ctx, oracle):
    """BASELINE.json's benchmark size (synthetic multirotor wake, 1 000 192 filaments x 258 176 targets): the whole
    sweep through the shared-node kernel, checked (1) against the CPU oracle on 96 sampled targets (reference
    enumeration, double and long double), (2) against the flat kernel on 4096 sampled targets, (3) for exact linearity
    in the circulation (doubling every Gamma doubles every velocity bit for bit)."""
    import torch
    lats = synth.multirotor(1_000_000, seed=12345)
    keep = _upload(ctx, lats)
    P = synth.targets_all(lats)
    m = P.shape[0]
    info = ctx.set_info(3)
    assert info["filaments"] == 1000192 and m == 258176 and info["shared_active"] == 1 and info["strip_width"] == 4
    V = _sweep(ctx, P)
    assert np.all(np.isfinite(V))
    rng = np.random.default_rng(0)
    # (1) oracle on a sample (6e7 pair evaluations on the CPU)
    idx = np.sort(rng.choice(m, size=96, replace=False))
    p1, p2, rvc, gam, flag = synth.flatten_all(lats)
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, P[idx])
    Vl, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, P[idx])
    e = scaled_err(V[idx], Vo, Vabs)
    el, eo = scaled_err(V[idx], Vl, Vabs), scaled_err(Vo, Vl, Vabs)
    print(f"1e6 filaments: scaled error vs oracle {e:.3e}; vs long double: GPU {el:.3e}, reference-order double sum {eo:.3e}")
    assert e < TOL and el < 5 * max(eo, 1e-15)
    # (2) flat kernel (reference enumeration) on a larger sample
    idx2 = np.sort(rng.choice(m, size=4096, replace=False))
    ctx.set_shared_nodes(False)
    try:
        Vf = _sweep(ctx, P[idx2])
    finally:
        ctx.set_shared_nodes(True)
    scale = np.max(Vabs)
    assert np.max(np.abs(Vf - V[idx2])) < TOL * scale
    # (3) linearity
    for l in lats:
        l.gam *= 2.0
        l.gamF *= 2.0
    keep2 = _upload(ctx, lats)
    V2 = _sweep(ctx, P)
    assert np.array_equal(V2, 2.0 * V)


@pytest.mark.parametrize("ns,main,tail", [(5, 4, 1), (6, 4, 2), (7, 4, 3), (8, 4, 0), (9, 4, 1), (10, 4, 2), (13, 4, 1)])
def test_tail_strips_cover_column_counts_that_are_not_multiples_of_four(ctx, oracle, ns, main, tail):
    """A rotor's lattice with ns columns is covered by floor(ns/4) strips of width 4 plus ONE strip of width ns mod 4
    (no padded columns) -- a second, small launch of the same kernel.  Same sums as the reference's ring-by-ring
    enumeration: compared with the oracle on random targets and on the wake's own nodes (guarded pairs), for the current
    and the predicted set; forcing one width with vlc_set_lattice_tuning gives the same velocities."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    ctx.set_lattice_tuning(5, 0)       # the tail-strip cover also for a wake this small (default: only from 2e4 rings on)
    ro = _make_rotor_pair(ctx, oracle, seed=40 + ns, ns=ns, nNwake=9, nFwake=4, rowNear=2, rowFar=2)
    rng = np.random.default_rng(ns)
    nodes = np.concatenate([ro.waN(ib)[:, 1:, 12:15].reshape(-1, 3) for ib in range(ro.nb)])     # corner 2 of the active rings
    P = np.concatenate([rng.uniform(-1.5, 1.5, size=(300, 3)), nodes])
    s = _tol_scale(ro, P) * 50
    res = {}
    for pred in (False, True):
        info = ctx.rotor_info(0, pred)
        assert info["shared_active"] == 1 and info["strip_width"] == main and info["tail_strip_width"] == tail, info
        assert info["lattice_records"] == ro.nb * (ns // 4 + (1 if tail else 0)) * (8 + 1), info
        res[pred] = ctx.rotor_vind_bywake(0, P, pred)
        assert np.max(np.abs(res[pred] - ro.vind_points(1, P, pred))) < TOL * s
    assert np.max(np.abs(ctx.rotor_vind(0, P) - ro.vind_points(2, P))) < TOL * s
    try:
        ctx.set_lattice_tuning(2, 0)                 # one width for every strip (the last one padded when ns is odd)
        ro = _make_rotor_pair(ctx, oracle, seed=40 + ns, ns=ns, nNwake=9, nFwake=4, rowNear=2, rowFar=2)
        info = ctx.rotor_info(0)
        assert info["strip_width"] == 2 and info["tail_strip_width"] == 0 and info["shared_active"] == 1, info
        assert np.max(np.abs(ctx.rotor_vind_bywake(0, P) - res[False])) < TOL * s
    finally:
        ctx.set_lattice_tuning(0, 0)
