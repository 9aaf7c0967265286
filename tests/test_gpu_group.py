"""GPU tests of the multi-GPU data plane BEHIND the C ABI (round-1 review, task 2): one handle made by
vlc_create_multi looks like a single-GPU context to the caller -- the reference's call sites (main.f90:814-841 ->
libCommon.f90:114-171) stay as they are -- while the library replicates the state on every member, shards the targets of
every sweep and all-gathers the velocity slices of the resident wake sweep.

On a box with ONE GPU the group is made of several members on device 0 (slices exchanged with peer copies); with two or
more GPUs the same tests run on distinct devices and the exchange is ncclAllGather.  With a fixed source split
(vlc_set_tuning) the summation order of a target does not depend on which member evaluates it, so the results are
BIT-IDENTICAL to a single-GPU context; with the automatic split they agree to rounding (the split is chosen per launch
from the number of targets in it).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.helpers import scaled_err
from volcanor_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [0] * n


@pytest.fixture(scope="module", params=[2, 3])
def gctx(request):
    import volcanor_b200 as vb
    c = vb.Context(devices=_devices(request.param))
    yield c
    c.close()


def test_group_reports_its_partition(gctx):
    import torch
    info = gctx.comm_info()
    assert info["world"] in (2, 3) and info["rank"] == 0
    assert info["transport"] == ("nccl" if torch.cuda.device_count() >= info["world"] else "peer")


@pytest.mark.parametrize("n,m", [(1000, 77), (5000, 2049), (20000, 3)])
def test_group_flat_sweep_vs_single_gpu_and_oracle(ctx, gctx, oracle, n, m):
    """vlc_set_sources + vlc_vind through the group handle: per-target 1e-12 against the oracle, bit-identical to the
    single-GPU context under a fixed source split, to rounding under the automatic one."""
    p1, p2, rvc, gam, flag, P = synth.random_filaments(n, m, seed=n + m)
    Vo = oracle.vind_flat(p1, p2, rvc, gam, flag, P)
    _, Vabs = oracle.vind_flat_ld(p1, p2, rvc, gam, flag, P)
    for nsplit in (0, 5):
        out = []
        for c in (ctx, gctx):
            c.set_tuning(0, nsplit)
            try:
                c.set_sources(0, p1, p2, rvc, gam, flag)
                out.append(c.vind(0, P))
            finally:
                c.set_tuning(0, 0)
        assert scaled_err(out[1], Vo, Vabs) < 1e-12
        if nsplit:
            assert np.array_equal(out[0], out[1])
        else:
            assert scaled_err(out[0], out[1], Vabs) < 1e-13


def test_group_rotor_call_sites(ctx, gctx, oracle):
    """The tier-2 call sites of the shim through the group handle: vind_bywing / bywake / boundVortices / chordwise,
    vind_onNwake_byRotor, vind_onFwake_byRotor (libCommon.f90:114-211), calcAIC + solve."""
    from tests.test_gpu_parity import _make_rotor_pair, _tol_scale
    ro = _make_rotor_pair(gctx, oracle, seed=8, rowNear=3, rowFar=2)
    _make_rotor_pair(ctx, oracle, seed=8, rowNear=3, rowFar=2)
    P = np.random.default_rng(7).uniform(-1.5, 1.5, size=(333, 3))
    s = _tol_scale(ro, P) * 50
    for c in (ctx, gctx):
        c.set_tuning(0, 3)
    try:
        for kind, fn in [(0, "rotor_vind_bywing"), (1, "rotor_vind_bywake"), (3, "rotor_vind_bywing_boundVortices"),
                         (4, "rotor_vind_bywing_chordwiseVortices")]:
            got = getattr(gctx, fn)(0, P)
            assert np.max(np.abs(got - ro.vind_points(kind, P))) < 1e-12 * s, fn
            assert np.array_equal(got, getattr(ctx, fn)(0, P)), fn
        for pred in (False, True):
            for ib in range(ro.nb):
                ref = ro.vind_onNwake_byRotor(ro, ib, 3, pred)
                got = gctx.vind_onNwake_byRotor(0, ro.waN(ib, pred), ro.nNwake - 2, ro.ns, ro.nNwake, pred, offset_records=2)
                assert got.shape == ref.shape and np.max(np.abs(got - ref)) < 1e-12 * s
                assert np.array_equal(got, ctx.vind_onNwake_byRotor(0, ro.waN(ib, pred), ro.nNwake - 2, ro.ns, ro.nNwake, pred,
                                                                    offset_records=2))
                gotf = gctx.vind_onFwake_byRotor(0, ro.waF(ib, pred), ro.nFwake - 1, pred, offset_records=1)
                assert np.max(np.abs(gotf - ro.vind_onFwake_byRotor(ro, ib, 2, pred))) < 1e-12 * s
    finally:
        for c in (ctx, gctx):
            c.set_tuning(0, 0)
    A = gctx.rotor_calcAIC(0, ro.N)
    assert ro.calcAIC() == 0
    assert np.max(np.abs(A - ro.AIC())) < 1e-12 * np.max(np.abs(A))
    assert np.array_equal(A, ctx.rotor_calcAIC(0, ro.N))
    rhs = np.random.default_rng(1).uniform(-1, 1, ro.N)
    assert np.array_equal(gctx.rotor_solve(0, rhs), ctx.rotor_solve(0, rhs))


@pytest.mark.parametrize("name,nsteps,cp", [("katzNplotkin_AR04", 160, False), ("elevateTest", 150, True),
                                            ("caradonna", 40, True)])
def test_group_resident_case_is_bit_identical_to_one_gpu(ctx, gctx, oracle, name, nsteps, cp):
    """A whole case with the wake resident on the device(s), driven by the SAME single-process driver through the same
    hook table: handle of one GPU vs group handle.  Fixed source split -> force histories identical bit for bit, every
    step; golden files reproduced to 7 digits through the group; members end with bit-identical wakes (the all-gather
    keeps the replicas in step)."""
    from tests.test_gpu_resident import _resident_hooks, _short_caradonna, _step
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    if name == "caradonna":
        _short_caradonna(fx)
    hist = []
    for c in (ctx, gctx):
        c.set_tuning(0, 2)
        try:
            case = oracle.Case(fx)
            lib, h = _resident_hooks(case, c)
            if cp:
                import ctypes as C
                lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
                assert lib.case_hooks_enable_cp(h) == 0, c.lib.vlc_last_error(c.h)
            case.init()
            f = [case.force_nondim(0).copy()]
            for it in range(nsteps):
                _step(case, lib, h, c, it + 1)
                f.append(case.force_nondim(0).copy())
            hist.append(np.array(f))
            lib.case_gpu_hooks_free(h)
        finally:
            c.set_tuning(0, 0)
    assert np.array_equal(hist[0], hist[1]), float(np.max(np.abs(hist[0] - hist[1])))
    if "ref_ForceNonDim" in fx and name != "caradonna":
        ref = np.array(fx["ref_ForceNonDim"]["rows"])[:nsteps + 1]
        ulp = 10.0 ** (np.floor(np.log10(np.abs(ref[:, 1]))) - 6)
        dev = np.abs(hist[1][:, 0] - ref[:, 1]) / ulp
        print(f"{name}: {nsteps} steps through a group of {gctx.comm_info()['world']} ({gctx.comm_info()['transport']}): "
              f"max deviation from the golden file {dev.max():.2f} units of the 7th digit")
        assert dev.max() <= 1.0


def test_group_shares_source_splits_of_the_collocation_point_stage(ctx, oracle, monkeypatch):
    """SURVEY 8e, "RHS: shard sources and sum partials": with NCCL between distinct devices the collocation-point sweeps of a
    group share their SOURCE splits out (each member a part of every slot range, all-reduce of the partial buffer, the same
    fixed-order reduce) -- here forced for every sweep by VLC_RHS_SHARE_MIN_PAIRS=0.  The all-reduce adds zeros to the one
    contribution of each slot, so the force history of a case is bit-identical to the one-GPU run and to a group that does
    not share (the wake sweeps, sharded by targets, need the fixed split for that)."""
    import torch
    import volcanor_b200 as vb
    from tests.test_gpu_resident import _resident_hooks, _step
    import ctypes as C
    if torch.cuda.device_count() < 2:
        pytest.skip("needs NCCL between two devices")
    fx = json.loads((GOLDEN / "elevateTest.json").read_text())

    def run(c, nsteps=40):
        c.set_tuning(0, 4)
        try:
            case = oracle.Case(fx)
            lib, h = _resident_hooks(case, c)
            lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
            assert lib.case_hooks_enable_cp(h) == 0, c.lib.vlc_last_error(c.h)
            case.init()
            n0 = c.launch_count
            f = [case.force_nondim(0).copy()]
            for it in range(nsteps):
                _step(case, lib, h, c, it + 1)
                f.append(case.force_nondim(0).copy())
            lib.case_gpu_hooks_free(h)
            return np.array(f), c.launch_count - n0
        finally:
            c.set_tuning(0, 0)

    ref, _ = run(ctx)
    out = {}
    for share in ("0", "-1"):
        monkeypatch.setenv("VLC_RHS_SHARE_MIN_PAIRS", share)
        g = vb.Context(devices=[0, 1])
        try:
            assert g.comm_info()["transport"] == "nccl"
            out[share] = run(g)
        finally:
            g.close()
    assert np.array_equal(out["0"][0], ref) and np.array_equal(out["-1"][0], ref)
    assert out["0"][1] > out["-1"][1], "no all-reduce was launched: the source splits were not shared"


def test_group_refuses_device_pointer_entry_points(gctx):
    """A device pointer belongs to one device: the tier-3 / _dev entry points return VLC_ERR_STATE on a group handle."""
    import torch

    import volcanor_b200 as vb
    x = torch.zeros(30, dtype=torch.float64, device="cuda:0")
    with pytest.raises(vb.VlcError, match="ONE device"):
        gctx.convect_dev(10, x, x, 0.1)
    with pytest.raises(vb.VlcError, match="ONE device"):
        gctx.vind_dev(0, 10, x, x)
    with pytest.raises(vb.VlcError, match="ONE device"):
        gctx.wake_sweep_slice(False, 0, 0, x)
    with pytest.raises(vb.VlcError, match="caller's stream"):
        gctx.set_stream(torch.cuda.current_stream().cuda_stream)
    gctx.set_stream(None, use_own=True)


def test_group_gridgen_matches_single_gpu(ctx, gctx, oracle):
    """program gridgen through the group handle: every member takes a slice of the cell list (gridgen.f90:116-139)."""
    rng = np.random.default_rng(3)
    lat = synth._helix_lattice(rng, np.zeros(3), 1.0, 0.1, 6, 9, 5, psi0=0.2, sense=1.0)
    p1, p2, rvc, gam, flag = lat.flatten()
    n = rvc.size
    vf = np.zeros((n, 12))
    vf[:, 0:3], vf[:, 3:6], vf[:, 8], vf[:, 9] = p1, p2, rvc, rvc
    z = np.zeros((0, 50))
    args = (9, 8, 7, [-2.0, -2.0, -2.0], [2.0, 2.0, 1.0], [1.0, 0.0, 0.0], z, z, vf, gam, np.zeros((0, 12)), np.zeros(0))
    out = []
    for c in (ctx, gctx):
        c.set_tuning(0, 2)
        try:
            out.append(c.gridgen(*args))
        finally:
            c.set_tuning(0, 0)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    gc, vc = oracle.gridgen(*args)
    assert np.array_equal(gc, out[1][0]) and np.max(np.abs(vc - out[1][1])) < 1e-12 * np.max(np.abs(vc))


def test_group_of_one_is_a_plain_context():
    import volcanor_b200 as vb
    c = vb.Context(devices=[0])
    assert c.comm_info() == {"world": 1, "rank": 0, "transport": "single"}
    p1, p2, rvc, gam, flag, P = synth.random_filaments(100, 10, seed=1)
    c.set_sources(0, p1, p2, rvc, gam, flag)
    assert np.all(np.isfinite(c.vind(0, P)))
    c.close()


def test_library_owned_communicator_one_process_per_gpu():
    """vlc_comm_init_rank: two processes, one GPU each, the driver calls plain vlc_wake_sweep and the library all-gathers
    (NCCL).  Needs two GPUs (NCCL refuses two ranks on one device); tests/test_gpu_resident.py covers the one-GPU box with
    the explicit slice / scatter entry points."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.test_gpu_resident import _run_sharded
    out = _run_sharded("katzNplotkin_AR04", 160, extra=("--lib-comm",))
    print(out)
    assert out["ok"] and out["ranks_identical"] and out["golden_max_dev_7th_digit"] <= 1.0, out
    assert out["exchange"] == "nccl inside the library"
