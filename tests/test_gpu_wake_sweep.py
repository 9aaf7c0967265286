"""vlc_wake_sweep on several source rotors at once (the path bench.py times): the rotors' packed sets laid side by side and
swept as ONE set (capi.cu: build_ws_combined), against the oracle's flat evaluation of the reference enumeration, per
target.  Uniform cores (merged form), non-uniform streamwise cores on every rotor (dual form), mixed (flat enumeration for
that launch), rotors whose strip widths differ (rotor-by-rotor sweeps), and the same through a multi-GPU handle."""
import numpy as np
import pytest

from tests.helpers import scaled_err
from volcanor_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(c, rotors):
    c.rotors_clear()
    for ir, r in enumerate(rotors):
        c.rotor_define(ir, r["nb"], 1, r["ns"], r["nNwake"], r["nFwake"], 2)      # surfaceType 2: no wing sources
        c.rotor_set_wake_params(ir, r["nb"], 0, 0, 0, 1, r["ns"], 1.0, 5.0, 0.0, 0.0)
        c.rotor_set_rows(ir, 1, 1)
        for ib in range(r["nb"]):
            c.rotor_put_nwake(ir, ib, r["waN"][ib])
            if r["nFwake"]:
                c.rotor_put_fwake(ir, ib, r["waF"][ib])


def _velocities(c, rotors):
    c.wake_sweep(False)
    out = []
    for ir, r in enumerate(rotors):
        for ib in range(r["nb"]):
            vn, vf = c.rotor_get_wakevel(ir, ib, 0, r["nNwake"], r["ns"], r["nFwake"])
            out += [vn.reshape(-1, 3), vf]
    return np.concatenate(out)


def _oracle(oracle, rotors):
    """flat filament list of the records themselves (so that edited core radii are what the oracle sees)"""
    p1, p2, rvc, gam, flag = [], [], [], [], []
    for r in rotors:
        for ib in range(r["nb"]):
            w, f = r["waN"][ib], r["waF"][ib]                     # (ns, nNwake, 50), (nFwake, 13)
            for k in range(4):
                p1.append(w[:, :, 12 * k:12 * k + 3]); p2.append(w[:, :, 12 * k + 3:12 * k + 6]); rvc.append(w[:, :, 12 * k + 9])
            # enumeration: rings j outer / i inner / filament 1..4
            P1 = np.stack(p1[-4:], axis=2).reshape(-1, 3); P2 = np.stack(p2[-4:], axis=2).reshape(-1, 3)
            RV = np.stack(rvc[-4:], axis=2).reshape(-1); G = np.repeat(w[:, :, 48].reshape(-1), 4)
            del p1[-4:], p2[-4:], rvc[-4:]
            parts = [(P1, P2, RV, G, np.ones(G.size, np.uint8))]
            if r["nFwake"]:
                last = w[:, -1, :]                                # horseshoe correction: -vf(2) of the last near row
                parts.append((last[:, 12:15], last[:, 15:18], last[:, 21], -last[:, 48], np.zeros(last.shape[0], np.uint8)))
                parts.append((f[:, 0:3], f[:, 3:6], f[:, 9], f[:, 12], np.ones(f.shape[0], np.uint8)))
            for a, b, c_, d, e in parts:
                p1.append(a); p2.append(b); rvc.append(c_); gam.append(d); flag.append(e)
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs))
    return cat(p1), cat(p2), cat(rvc), cat(gam), cat(flag)


def _check(c, oracle, rotors, lats, expect_forms):
    _setup(c, rotors)
    V = _velocities(c, rotors)
    forms = [c.rotor_info(ir)["shared_active"] for ir in range(len(rotors))]
    assert forms == expect_forms, forms
    src = _oracle(oracle, rotors)
    P = synth.targets_all(lats)
    assert P.shape == V.shape and src[2].size == sum(c.rotor_info(ir)["filaments"] for ir in range(len(rotors)))
    Vo = oracle.vind_flat(*src, P)
    _, Vabs = oracle.vind_flat_ld(*src, P)
    e = scaled_err(V, Vo, Vabs)
    assert np.all(np.isfinite(V)) and e < TOL, e
    return V, e


def _slope(rotors, which, slope=0.05):
    for ir in which:
        r = rotors[ir]
        f = 1.0 + slope * np.arange(r["ns"])[:, None]
        for w in r["waN"]:
            for k in (8, 9, 24 + 8, 24 + 9):
                w[:, :, k] *= f


@pytest.mark.parametrize("case", ["merged", "dual", "mixed", "widths differ"])
def test_combined_wake_sweep_vs_oracle(ctx, oracle, case):
    lats = synth.multirotor(24000, seed=3, n_rotor=3, nb=2, S=8, F=6, with_wing=True)
    rotors = synth.rotors_from_lattices(lats)
    assert [r["nb"] for r in rotors] == [2, 2, 2, 1]
    expect = [1, 1, 1, 1]
    if case == "dual":
        _slope(rotors, range(4))
        expect = [2, 2, 2, 2]
    elif case == "mixed":
        _slope(rotors, [1])
        expect = [1, 2, 1, 1]
    elif case == "widths differ":                                   # another rotor with 6 columns: strips of width 2 there
        extra = synth.rotors_from_lattices(synth.multirotor(3000, seed=4, n_rotor=1, nb=2, S=6, F=4, with_wing=False))
        lats = lats + synth.multirotor(3000, seed=4, n_rotor=1, nb=2, S=6, F=4, with_wing=False)
        rotors = rotors + extra
        expect = [1, 1, 1, 1, 1]
    V, e = _check(ctx, oracle, rotors, lats, expect)
    print(f"{case}: {V.shape[0]} wake-node targets x {sum(r['nb'] for r in rotors)} blades' wakes in one vlc_wake_sweep: per-target error {e:.2e}")


def test_combined_wake_sweep_through_a_group_is_bit_identical(ctx, oracle):
    import torch

    import volcanor_b200 as vb
    lats = synth.multirotor(24000, seed=3, n_rotor=3, nb=2, S=8, F=6, with_wing=True)
    rotors = synth.rotors_from_lattices(lats)
    g = vb.Context(devices=[0, 1, 2] if torch.cuda.device_count() >= 3 else [0, 0, 0])
    try:
        out = []
        for c in (ctx, g):
            c.set_tuning(0, 3)
            try:
                out.append(_check(c, oracle, rotors, lats, [1, 1, 1, 1])[0])
            finally:
                c.set_tuning(0, 0)
        assert np.array_equal(out[0], out[1])
    finally:
        g.close()
        ctx.rotors_clear()
