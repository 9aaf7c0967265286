"""Prescribed far wake (classdef.f90:222-236 pFwake_class, :998-1066 pFwake_update, :5170-5218
rotor_updatePrescribedWake, hook :4826-4828 in rotor_convectwake).  PARITY UNPINNED: no shipped case and no reference test
enables it (prescWakeAfterTruncNt = 0 everywhere), so these tests pin the oracle's restatement against a second, independent
numpy restatement written from the same source lines, against the structural properties the source guarantees, and check
that the staged orchestration (the library's resident mode keeps this generator on the host) reproduces the
driver's inline statement bit for bit, 240 helix filaments included."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
NPF, FW = 240, 13


def _np_pfwake_update(waF, hub, dpsi, pitch0, radius0):
    """pFwake_update (classdef.f90:998-1066) with numpy; waF (n, 13) records: fc(:,1) = [0:3], fc(:,2) = [3:6], rVc = [9],
    gam = [12].  Returns (records (240, 13) with only fc / rVc / gam filled, helixPitch, helixRadius)."""
    two_pi = 2.0 * (np.arctan(1.0) * 4.0)
    n = waF.shape[0]
    anchor = waF[-1, 0:3]
    radius = 0.0
    pitch = 0.0
    for i in range(n):                                           # :1022-1030, summation order of the source
        radius = radius + np.sqrt(waF[i, 1] ** 2 + waF[i, 0] ** 2)
        if i < n - 1:
            pitch = pitch + waF[i, 2] - waF[i + 1, 2]
    pitch = abs(pitch) * (-two_pi / dpsi) / (n - 1)
    radius = radius / n
    pitch = 0.5 * pitch + (1 - 0.5) * pitch0                     # :1035-1038
    radius = 0.5 * radius + (1 - 0.5) * radius0
    d_theta = np.arctan2(anchor[1], anchor[0])
    dz = anchor[2] - hub[2]
    theta = -1.0 * (np.arange(NPF + 1) * ((two_pi * 10.0) / NPF) + 0.0)   # linspace libMath.f90:138-157, clockwise
    coords = np.stack([radius * np.cos(theta + d_theta), radius * np.sin(theta + d_theta),
                       pitch * np.abs(theta) / two_pi + dz])     # (3, 241)
    out = np.zeros((NPF, FW))
    out[:, 3:6] = (hub[:, None] + coords[:, :NPF]).T             # assignP(2, ...) :1057
    out[:, 0:3] = (hub[:, None] + coords[:, 1:]).T               # assignP(1, ...) :1058
    out[0, 3:6] = anchor                                         # :1062
    out[:, 12] = waF[-1, 12]
    out[:, 9] = waF[-1, 9]
    return out, pitch, radius


def _call_update(oracle, pf, pitch, radius, waF, hub, axis, dpsi):
    lib = oracle.load()
    p, r = C.c_double(pitch), C.c_double(radius)
    waF = np.ascontiguousarray(waF)
    hub = np.ascontiguousarray(hub, dtype=np.float64)
    axis = np.ascontiguousarray(axis, dtype=np.float64)
    rc = lib.orc_pfwake_update(pf.ctypes.data, C.addressof(p), C.addressof(r), waF.ctypes.data, waF.shape[0],
                               hub.ctypes.data, axis.ctypes.data, dpsi)
    return rc, p.value, r.value


def _helical_far_wake(rng, n, hub):
    psi = -np.arange(n)[::-1] * np.deg2rad(12.0) + 0.3           # last record = the newest... any order is legal input
    rad = 1.9 + 0.05 * rng.standard_normal(n)
    waF = np.zeros((n, FW))
    waF[:, 0] = hub[0] * 0 + rad * np.cos(psi)                   # the source measures the radius from the origin, not the hub
    waF[:, 1] = rad * np.sin(psi)
    waF[:, 2] = hub[2] - 0.8 - 0.07 * np.arange(n)[::-1] + 0.002 * rng.standard_normal(n)
    waF[1:, 3:6] = waF[:-1, 0:3]
    waF[0, 3:6] = waF[0, 0:3] + [0.1, 0.0, 0.05]
    waF[:, 9] = 0.03 + 0.01 * rng.random(n)
    waF[:, 8] = waF[:, 9]
    waF[:, 12] = -1.3 + 0.1 * rng.random(n)
    return waF


@pytest.mark.parametrize("n,seed", [(2, 1), (7, 2), (40, 3)])
def test_pfwake_update_against_numpy_restatement(oracle, n, seed):
    rng = np.random.default_rng(seed)
    hub = np.array([0.0, 0.0, 0.4])
    waF = _helical_far_wake(rng, n, hub)
    dpsi = 130.9 * 1.92e-3
    pf = np.full((NPF, FW), 7.0)                                  # members the update must leave alone keep the 7
    rc, pitch, radius = _call_update(oracle, pf, 0.0, 0.0, waF, hub, [0, 0, 1], dpsi)
    assert rc == 0
    ref, pitch_np, radius_np = _np_pfwake_update(waF, hub, dpsi, 0.0, 0.0)
    assert pitch == pitch_np and radius == radius_np              # plain sums and products: bit-identical
    assert pitch < 0 and abs(2 * radius - 1.9) < 0.1              # descending along -z; half the fit: relaxed from 0
    for sl in (slice(0, 3), slice(3, 6)):                         # cos / sin: libm vs numpy, a few ulp of the radius
        assert np.max(np.abs(pf[:, sl] - ref[:, sl])) <= 8 * np.finfo(float).eps * (abs(radius) + 10 * abs(pitch) + 1.0)
    assert np.array_equal(pf[:, 9], ref[:, 9]) and np.array_equal(pf[:, 12], ref[:, 12])
    assert np.all(pf[:, [6, 7, 8, 10, 11]] == 7.0)                # l0, lc, rVc0, age, ageAzimuthal untouched (:1064-1065)
    # structure: starts at the anchor, contiguous polyline, 10 revolutions of 15 degree filaments at the fitted radius
    assert np.array_equal(pf[0, 3:6], waF[-1, 0:3])
    assert np.array_equal(pf[1:, 3:6], pf[:-1, 0:3])
    r_xy = np.hypot(pf[:, 0] - hub[0], pf[:, 1] - hub[1])
    assert np.allclose(r_xy, radius, rtol=1e-13)
    ang = np.unwrap(np.arctan2(pf[:, 1], pf[:, 0]))
    assert np.allclose(np.diff(ang), -np.deg2rad(15.0), atol=1e-12)
    assert np.allclose(np.diff(pf[:, 2]), pitch / 24.0, rtol=1e-10)
    # relaxation against the previous fit (:1035-1038): a second call moves half of the remaining way
    rc, pitch2, radius2 = _call_update(oracle, pf, pitch, radius, waF, hub, [0, 0, 1], dpsi)
    assert rc == 0
    assert pitch2 == 0.5 * (2 * pitch) + 0.5 * pitch and radius2 == 0.5 * (2 * radius) + 0.5 * radius


def test_pfwake_update_refuses_tilted_shaft(oracle):
    """error stop "Prescribed far wake only implemented for shaft along Z-axis" (classdef.f90:1010-1012)."""
    rng = np.random.default_rng(5)
    hub = np.zeros(3)
    waF = _helical_far_wake(rng, 5, hub)
    pf = np.zeros((NPF, FW))
    rc, _, _ = _call_update(oracle, pf, 0.0, 0.0, waF, hub, [0.0, 1e-3, 1.0], 0.25)
    assert rc != 0 and not pf.any()


def rotor_fx(small=True):
    """Case fixture for these tests.  small: a three-blade axisymmetric hovering rotor with a coarse lattice (tutorials/
    caradonna.case with nb = 3, nc = 3, ns = 6): steps in milliseconds on the CPU.  Otherwise tests/elevateTest.case: five
    blades from a PLOT3D grid, 16 x 16 panels each."""
    if not small:
        return json.loads((GOLDEN / "elevateTest.json").read_text())
    fx = json.loads((GOLDEN / "caradonna.json").read_text())
    fx["config"]["nt"] = 40
    fx["geom"][0].update(nb=3, nc=3, ns=6, axisymmetrySwitch=1)
    return fx


def _with_prescribed_wake(gen):
    def m(fx):
        g = fx["geom"][0]
        g["nNwake"] = 6                      # roll-up, far wake and truncation inside a short window
        g["wakeTruncateNt"] = 10
        g["prescWakeAfterTruncNt"] = 2       # prescWakeNt = 12: the helix is attached from step 13 on (:3013-3017, :4826)
        g["prescWakeGenNt"] = gen            # 0: fitted to rows rowFar..nFwakeEnd; 2: to the last three rows (:5180-5184)
    return m


@pytest.mark.parametrize("fd,gen,small", [(3, 0, False), (3, 2, True), (1, 0, True), (0, 0, True), (5, 2, True)])
def test_prescribed_wake_staged_orchestration_equals_inline_time_loop(oracle, fd, gen, small):
    from tests.test_staged_hooks import _lib
    fx = rotor_fx(small)                                             # axisymmetric: copy + rotate of :5203-5216
    _with_prescribed_wake(gen)(fx)
    fx["config"]["fdScheme"] = fd
    plain = json.loads(json.dumps(fx))
    plain["geom"][0]["prescWakeAfterTruncNt"] = 0
    lib = _lib()
    a, b = oracle.Case(fx), oracle.Case(fx)
    c = oracle.Case(plain) if (fd, gen) == (3, 0) else None   # the same case without the helix: one comparison is enough
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    assert h
    runs = [x for x in (a, b, c) if x is not None]
    for x in runs:
        x.init()
    assert a.rotor(0).nFwake == 4
    differs = False
    for it in range(18):
        for x in runs:
            x.step()
        assert np.array_equal(a.force_nondim(0), b.force_nondim(0)), (fd, it + 1)
        assert a.pairs_last_step == b.pairs_last_step
        if c is None:
            continue
        if it + 1 <= 13:   # attached at the end of step 13 (iter > prescWakeNt = 12): first felt by the RHS of step 14
            assert np.array_equal(a.force_nondim(0), c.force_nondim(0)), (fd, it + 1)
        else:
            differs = differs or not np.array_equal(a.force_nondim(0), c.force_nondim(0))
    assert differs or c is None                                               # the 240 filaments per blade are live sources
    ra, rb = a.rotor(0), b.rotor(0)
    two_pi_5 = 2.0 * np.pi / ra.nb
    for ib in range(ra.nb):
        for pred in (False, True):
            assert np.array_equal(ra.wapF(ib, pred), rb.wapF(ib, pred)), (fd, "wapF", ib, pred)
            assert np.array_equal(ra.waF(ib, pred), rb.waF(ib, pred))
            assert np.array_equal(ra.waN(ib, pred), rb.waN(ib, pred))
        w = ra.wapF(ib)
        # one strength for the whole helix (the last far filament's at the time of the update; the roll-up / shiftFwake that
        # end the step have moved the far wake on since)
        assert np.all(np.abs(w[:, 12]) > 0) and np.all(w[:, 12] == w[0, 12]) and np.all(w[:, 9] == w[0, 9])
        assert np.array_equal(w[1:, 3:6], w[:-1, 0:3])
        if ib > 0:                                                        # blades 2..nb: rotated copies of blade 1's
            w0 = ra.wapF(0)
            a0 = np.arctan2(w0[:, 1], w0[:, 0])
            ai = np.arctan2(w[:, 1], w[:, 0])
            d = np.angle(np.exp(1j * (ai - a0 - two_pi_5 * ib)))
            assert np.max(np.abs(d)) < 1e-12 and np.allclose(w[:, 2], w0[:, 2], rtol=0, atol=1e-13)
    lib.case_gpu_hooks_free(h)


# ------------------------------------------------------------------------------------------------------- wake burst
# rotor%burst_wake() (classdef.f90:4911-4917 -> :2306-2339): PARITY UNPINNED as well (wakeBurst = 0 in every shipped case).

def _kinked_far_wake(rng, n):
    """A far-wake chain (fc2(i+1) = fc1(i)) along -z with a few sharp kinks."""
    nodes = np.zeros((n + 1, 3))
    step = np.array([0.0, 0.0, -0.1])
    for i in range(1, n + 1):
        d = step + 0.004 * rng.standard_normal(3)
        if i in (3, 4, n - 2):
            d = d + np.array([0.12, -0.05, 0.03])              # a kink well beyond the limit used below
        nodes[i] = nodes[i - 1] + d
    waF = np.zeros((n, FW))
    waF[:, 3:6] = nodes[:-1]                                    # fc(:,2): the newer end
    waF[:, 0:3] = nodes[1:]                                     # fc(:,1): the older end; fc2(i+1) = fc1(i) (wake_continuity)
    waF[:, 9] = 0.02 + 0.001 * np.arange(n)
    waF[:, 12] = -1.0
    return waF


def test_burst_wake_oracle_properties_and_the_products_kernel(oracle):
    rng = np.random.default_rng(21)
    n, limit, core = 12, 0.05, 0.77
    waF = _kinked_far_wake(rng, n)
    assert np.array_equal(waF[1:, 3:6], waF[:-1, 0:3])          # the chain rule of the far wake (wake_continuity)
    from tests.test_kernels_emul import emul_lib
    olib, lib = oracle.load(), emul_lib()
    olib.orc_burst_pair.restype = C.c_int
    olib.orc_burst_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    # the skew by an independent formula: angle between successive segments, as a fraction of pi
    seg = waF[:, 3:6] - waF[:, 0:3]
    cosang = np.einsum("ij,ij->i", seg[:-1], -seg[1:]) / (np.linalg.norm(seg[:-1], axis=1) * np.linalg.norm(seg[1:], axis=1))
    skew = np.abs(np.arccos(np.clip(cosang, -1, 1)) - np.pi) / np.pi
    want = skew >= limit
    assert 2 <= want.sum() < n - 1 and np.min(np.abs(skew - limit)) > 1e-3      # clear decisions, both kinds
    got = np.array([olib.orc_burst_pair(waF[i].ctypes.data, waF[i + 1].ctypes.data, limit) for i in range(n - 1)], dtype=bool)
    assert np.array_equal(got, want)
    # straight chain: skew 0, nothing bursts; limit 0 bursts everything (>=)
    straight = waF.copy()
    straight[:, 0:3] = np.arange(n, 0, -1)[:, None] * np.array([0.0, 0.0, 0.1])
    straight[:, 3:6] = (np.arange(n, 0, -1)[:, None] + 1) * np.array([0.0, 0.0, 0.1])
    assert not olib.orc_burst_pair(straight[2].ctypes.data, straight[3].ctypes.data, 1e-7)
    assert olib.orc_burst_pair(straight[2].ctypes.data, straight[3].ctypes.data, 0.0)
    # the product's kernel (emulated, tests/native/kernels_emul.cpp): two blades, rowFar = 3 (rows 1, 2 inactive)
    two = np.stack([waF, waF[:, :]]).copy()
    two[1, :, 9] += 0.5
    before = two.copy()
    lib.emul_burst_wake(2, n, 3, limit, core, two.ctypes.data)
    for ib in range(2):
        hit = np.zeros(n, dtype=bool)
        for i in range(2, n - 1):                               # 0-based rows rowFar-1 .. nFwake-2
            if want[i]:
                hit[i] = hit[i + 1] = True
        assert np.array_equal(two[ib, hit, 9], np.full(hit.sum(), core))
        assert np.array_equal(two[ib, ~hit, 9], before[ib, ~hit, 9])
        keep = [k for k in range(FW) if k != 9]
        assert np.array_equal(two[ib][:, keep], before[ib][:, keep])


@pytest.mark.parametrize("fd", [3])
def test_wake_burst_staged_orchestration_equals_inline_time_loop(oracle, fd):
    """The driver's `mod(iter, wakeBurst)` statement (main.f90:490-497) through the staged orchestration: bit-identical to
    the inline loop, and the burst really changes far-wake core radii (and with them the loads)."""
    from tests.test_staged_hooks import _lib
    fx = rotor_fx()
    g = fx["geom"][0]
    g["nNwake"], g["wakeTruncateNt"], g["skewLimit"] = 6, 14, 0.004
    fx["config"]["fdScheme"] = fd
    fx["config"]["wakeBurst"] = 2
    plain = json.loads(json.dumps(fx))
    plain["config"]["wakeBurst"] = 0
    lib = _lib()
    a, b, c = oracle.Case(fx), oracle.Case(fx), oracle.Case(plain)
    b.init_rotors()
    h = lib.case_cpu_staged_hooks_install(b.h, b.nr)
    assert h
    for x in (a, b, c):
        x.init()
    for it in range(16):
        for x in (a, b, c):
            x.step()
        assert np.array_equal(a.force_nondim(0), b.force_nondim(0)), (fd, it + 1)
    ra, rb, rc = a.rotor(0), b.rotor(0), c.rotor(0)
    chord = g["chord"]
    burst = 0
    for ib in range(ra.nb):
        assert np.array_equal(ra.waF(ib), rb.waF(ib)) and np.array_equal(ra.waN(ib), rb.waN(ib))
        burst += int(np.sum(ra.waF(ib)[:, 9] == chord))
        assert not np.any(rc.waF(ib)[:, 9] == chord)
    assert burst >= 2
    assert not np.array_equal(a.force_nondim(0), c.force_nondim(0))
    lib.case_gpu_hooks_free(h)


# ------------------------------------------------------------ bodies shared by the GPU tests and their CPU emulation
# (tests/test_zzz_gpu_first_run.py calls them with a volcanor_b200.Context; here they run on EmulatedWakeContext: the
# product's kernels compiled for the host, so what the GPU tests upload, call and compare is exercised without a GPU)

def check_update_prescribed_wake(ctx, oracle, gen, axisym):
    from tests.test_zz_gpu_cp_stage import _define, _developed
    fx = rotor_fx()
    _with_prescribed_wake(gen)(fx)
    fx["geom"][0]["axisymmetrySwitch"] = axisym
    case = _developed(oracle, fx, 15)
    rot = case.rotor(0)
    p = rot.params()
    olib = oracle.load()
    _define(ctx, rot, 0)
    ctx.rotor_set_frame(0, p["shaftAxis"], p["hubCoords"])
    for ib in range(rot.nb):
        ctx.rotor_put_nwake(0, ib, rot.waN(ib, True), predicted=True)
        ctx.rotor_put_fwake(0, ib, rot.waF(ib, True), predicted=True)
    zero = np.zeros(2)
    for pred in (False, True):
        for ib in range(rot.nb):
            olib.orc_rotor_set_pfHelix(rot.h, ib, int(pred), zero.ctypes.data)
        for rep in range(2):
            dt = 0.0137 * (rep + 1)
            assert olib.orc_rotor_updatePrescribedWake(rot.h, dt, b"P" if pred else b"C") == 0
            ctx.rotor_updatePrescribedWake(0, p["omegaSlow"] * dt, gen, "P" if pred else "C")
            for ib in range(rot.nb):
                w, hx = ctx.rotor_get_pfwake(0, ib, pred)
                ref, hr = rot.wapF(ib, pred), np.zeros(2)
                olib.orc_rotor_get_pfHelix(rot.h, ib, int(pred), hr.ctypes.data)
                assert np.array_equal(hx, hr), (pred, rep, ib, hx, hr)
                assert np.array_equal(w[:, 9], ref[:, 9]) and np.array_equal(w[:, 12], ref[:, 12])
                assert np.all(np.abs(w[:, 12]) > 0)
                scale = abs(hr[1]) + 10 * abs(hr[0]) + np.max(np.abs(p["hubCoords"]))
                assert np.max(np.abs(w[:, 0:6] - ref[:, 0:6])) < 1e-13 * scale, (pred, rep, ib)


def check_burst_wake(ctx, oracle, fx):
    from tests.test_zz_gpu_cp_stage import _define, _developed
    fx["config"]["wakeBurst"] = 0
    case = _developed(oracle, fx, 15)
    rot = case.rotor(0)
    row0 = rot.dims()["rowFar"] - 1
    skews = []
    for ib in range(rot.nb):
        w = rot.waF(ib)[row0:]
        seg = w[:, 3:6] - w[:, 0:3]
        cosang = np.einsum("ij,ij->i", seg[:-1], -seg[1:]) / (np.linalg.norm(seg[:-1], axis=1) * np.linalg.norm(seg[1:], axis=1))
        skews += list(np.abs(np.arccos(np.clip(cosang, -1, 1)) - np.pi) / np.pi)
    sk = np.sort(np.array(skews))
    assert len(sk) >= 4
    k = int(np.nonzero(np.diff(sk) > 1e-6 * sk[-1])[0][-1])       # the last clear gap: only the sharpest kinks burst
    limit = 0.5 * (sk[k] + sk[k + 1])
    _define(ctx, rot, 0)
    core = 0.77
    ctx.rotor_burst_wake(0, limit, core)
    olib = oracle.load()
    olib.orc_burst_pair.restype = C.c_int
    olib.orc_burst_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    changed = 0
    for ib in range(rot.nb):
        before = rot.waF(ib).copy()
        ref = before.copy()
        for i in range(row0, rot.nFwake - 1):
            if olib.orc_burst_pair(before[i].ctypes.data, before[i + 1].ctypes.data, limit):
                ref[i, 9] = ref[i + 1, 9] = core
        got = ctx.rotor_get_fwake(0, ib, rot.nFwake)
        assert np.array_equal(got, ref), ib
        changed += int(np.sum(ref[:, 9] != before[:, 9]))
    assert changed >= 2 and changed < rot.nb * rot.nFwake


def check_calc_skew(ctx, oracle, axisym):
    """vlc_rotor_calc_skew against rotor%calc_skew() of the oracle (classdef.f90:4919-4936) on a developed near wake: the
    records BIT-IDENTICAL (sums, products, one sqrt, one division); values in [0, 1], 0 only where gam is 0."""
    from tests.test_zz_gpu_cp_stage import _define, _developed
    fx = rotor_fx()
    g = fx["geom"][0]
    g["nNwake"], g["wakeTruncateNt"], g["axisymmetrySwitch"] = 6, 10, axisym
    case = _developed(oracle, fx, 4)                                      # rowNear = 3: rows 1, 2 are not yet shed
    rot = case.rotor(0)
    assert rot.dims()["rowNear"] == 3
    _define(ctx, rot, 0)
    ctx.rotor_calc_skew(0)
    oracle.load().orc_rotor_calc_skew(rot.h)
    for ib in range(rot.nb):
        got, ref = ctx.rotor_get_nwake(0, ib, rot.nNwake, rot.ns), rot.waN(ib)
        # rows rowNear.. only: vlc_rotor_put_nwake transfers the active rows, the device's rows 1, 2 were never written
        assert np.array_equal(got[:, 2:, :], ref[:, 2:, :]), (ib, np.argwhere(got[:, 2:, :] != ref[:, 2:, :])[:4])
        sk = ref[:, 2:, 49]
        assert np.all((sk >= 0) & (sk <= 1)) and np.any(sk > 0)
        assert np.all((sk > 0) | (np.abs(ref[:, 2:, 48]) <= np.finfo(float).eps) | (sk == 0))
        assert np.all(ref[:, :2, 49] == 0)                               # inactive rows untouched
    if axisym:
        assert np.array_equal(rot.waN(rot.nb - 1)[:, 2:, 49], rot.waN(0)[:, 2:, 49])


class EmulatedWakeContext:
    """Stand-in for volcanor_b200.Context: its own copies of the wake records; the work is done by the product's KERNELS
    compiled for the host and run thread by thread (tests/native/kernels_emul.cpp, tests/test_kernels_emul.py)."""

    def __init__(self):
        from tests.test_kernels_emul import emul_lib
        self.lib = emul_lib()
        self.r = {}

    def rotor_define(self, ir, nb, nc, ns, nNwake, nFwake, surfaceType=1):
        self.r[ir] = dict(nb=nb, nFwake=nFwake, waF=np.zeros((2, nb, nFwake, FW)), wapF=np.zeros((2, nb, NPF, FW)),
                          helix=np.zeros((2, nb, 2)), nbConvect=nb, axisym=0, rowFar=nFwake + 1, axis=np.array([0.0, 0.0, 1.0]),
                          hub=np.zeros(3))

    def rotor_set_wake_params(self, ir, nbConvect, axisymmetrySwitch, *rest):
        self.r[ir].update(nbConvect=nbConvect, axisym=axisymmetrySwitch)

    def rotor_set_rows(self, ir, rowNear, rowFar):
        self.r[ir]["rowFar"] = rowFar
        self.r[ir]["rowNear"] = rowNear

    def rotor_set_frame(self, ir, shaftAxis, hubCoords):
        self.r[ir].update(axis=np.array(shaftAxis, dtype=float), hub=np.array(hubCoords, dtype=float))

    def rotor_put_wing(self, ir, ib, wiP):
        pass

    def rotor_put_nwake(self, ir, ib, waN, predicted=False):
        r = self.r[ir]
        if "waN" not in r:
            r["waN"] = np.zeros((2, r["nb"]) + np.asarray(waN).shape)
        r["waN"][int(predicted), ib] = waN

    def rotor_get_nwake(self, ir, ib, nNwake, ns, predicted=False):
        return self.r[ir]["waN"][int(predicted), ib].copy()

    def rotor_calc_skew(self, ir):
        r = self.r[ir]
        ns, nNwake = r["waN"].shape[2], r["waN"].shape[3]
        self.lib.emul_calc_skew(r["nb"], r["nbConvect"], r["axisym"], ns, nNwake, r["rowNear"], r["waN"][0].ctypes.data)

    def rotor_put_fwake(self, ir, ib, waF, predicted=False):
        self.r[ir]["waF"][int(predicted), ib] = waF

    def rotor_get_fwake(self, ir, ib, nFwake, predicted=False):
        return self.r[ir]["waF"][int(predicted), ib].copy()

    def rotor_get_pfwake(self, ir, ib, predicted=False):
        return self.r[ir]["wapF"][int(predicted), ib].copy(), self.r[ir]["helix"][int(predicted), ib].copy()

    def rotor_updatePrescribedWake(self, ir, deltaPsi, prescWakeGenNt=0, wakeType="C"):
        from oracle import pyoracle
        r, s = self.r[ir], {"C": 0, "P": 1}[wakeType]
        olib = pyoracle.load()
        olib.orc_getTransformAxis.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        nb = r["nb"]
        T, rotate = np.zeros((nb, 9)), np.zeros(nb, dtype=np.int32)
        two_pi = 2.0 * (np.arctan(1.0) * 4.0)
        for ib in range(1, nb):
            off = two_pi / nb * ib
            rotate[ib] = abs(off) > np.finfo(float).eps
            olib.orc_getTransformAxis(off, r["axis"].ctypes.data, T[ib].ctypes.data)
        rc = self.lib.emul_updatePrescribedWake(nb, r["nbConvect"], r["axisym"], r["nFwake"], r["rowFar"], prescWakeGenNt, deltaPsi,
                                     r["hub"].ctypes.data, T.ctypes.data, rotate.ctypes.data, r["waF"][s].ctypes.data,
                                     r["wapF"][s].ctypes.data, r["helix"][s].ctypes.data)
        assert rc == 0

    def rotor_burst_wake(self, ir, skewLimit, largeCoreRadius):
        r = self.r[ir]
        self.lib.emul_burst_wake(r["nb"], r["nFwake"], r["rowFar"], skewLimit, largeCoreRadius, r["waF"][0].ctypes.data)


@pytest.mark.parametrize("gen,axisym", [(0, 1), (2, 0)])
def test_body_of_the_gpu_update_test_on_emulated_kernels(oracle, gen, axisym):
    check_update_prescribed_wake(EmulatedWakeContext(), oracle, gen, axisym)


def test_body_of_the_gpu_burst_test_on_emulated_kernels(oracle):
    fx = rotor_fx()
    g = fx["geom"][0]
    g["nNwake"], g["wakeTruncateNt"], g["skewLimit"] = 6, 14, 0.004
    check_burst_wake(EmulatedWakeContext(), oracle, fx)


@pytest.mark.parametrize("axisym", [1, 0])
def test_body_of_the_gpu_skew_test_on_emulated_kernels(oracle, axisym):
    check_calc_skew(EmulatedWakeContext(), oracle, axisym)


def test_skew_oracle_against_an_independent_formula(oracle):
    """|cos| of the angle between the bimedians (lines joining the midpoints of opposite sides) of a quadrilateral ring."""
    rng = np.random.default_rng(8)
    olib = oracle.load()
    olib.orc_vr_skew.restype = C.c_double
    olib.orc_vr_skew.argtypes = [C.c_void_p]
    for trial in range(20):
        corners = rng.standard_normal((4, 3)) + np.array([[0, 0, 0], [2, 0, 0], [2, 2, 0], [0, 2, 0]])
        if trial == 0:
            corners = np.array([[0.0, 0, 0], [1, 0, 0], [1, 3, 0], [0, 3, 0]])      # rectangle: bimedians orthogonal
        ring = np.zeros(50)
        for n in range(4):
            ring[12 * n:12 * n + 3] = corners[n]
            ring[12 * n + 3:12 * n + 6] = corners[(n + 1) % 4]
        ring[48] = -0.7
        m = [0.5 * (corners[n] + corners[(n + 1) % 4]) for n in range(4)]       # side midpoints 12, 23, 34, 41
        b1, b2 = m[2] - m[0], m[3] - m[1]
        want = abs(b1 @ b2) / (np.linalg.norm(b1) * np.linalg.norm(b2))
        got = olib.orc_vr_skew(ring.ctypes.data)
        assert abs(got - want) < 1e-14, (trial, got, want)
        if trial == 0:
            assert got == 0.0
        ring[48] = 0.0
        assert olib.orc_vr_skew(ring.ctypes.data) == 0.0
