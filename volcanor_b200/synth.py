"""Synthetic multirotor wakes (BASELINE.json configs[4]) -- workload generator.

Host-side numpy only; produces the same data for the GPU path, the CPU oracle and the benchmark:
per blade a helical near-wake lattice of (R+1) x (S+1) nodes whose rings follow the reference's corner
order (vr_assignP, src/classdef.f90:569-592), a far-wake chain of F filaments from the tip, and one
fixed wing with a flat wake.  Circulations, core radii and node jitter are seeded (PCG64).

A lattice is described as in include/volcanor_b200.h (tier 3):
  nodes (S+1, R+1, 3)  [= Fortran (3, R+1, S+1)], gam (S, R), rvc4 (S, R, 4), far_nodes (F+1, 3), gamF, rvcF.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Lattice:
    nodes: np.ndarray       # (S+1, R+1, 3)
    gam: np.ndarray         # (S, R)
    rvc4: np.ndarray        # (S, R, 4)
    far_nodes: np.ndarray   # (F+1, 3) or (0, 3)
    gamF: np.ndarray        # (F,)
    rvcF: np.ndarray        # (F,)
    meta: dict = field(default_factory=dict)

    @property
    def R(self):
        return self.gam.shape[1]

    @property
    def S(self):
        return self.gam.shape[0]

    @property
    def F(self):
        return self.gamF.shape[0]

    def n_filaments(self) -> int:
        return 4 * self.R * self.S + ((self.S + self.F) if self.F > 0 else 0)

    def flatten(self):
        """Filaments in the reference's enumeration (blade_vind_bywake, classdef.f90:1450-1469):
        rings j outer / i inner / filament 1..4, then -vf2 of the last row (horseshoe), then the far chain.
        Returns p1, p2 (n,3), rvc, gam, wake_flag."""
        S, R = self.S, self.R
        nd = self.nodes
        c1 = nd[:-1, :-1]  # corner 1 = (r, j)
        c2 = nd[:-1, 1:]   # corner 2 = (r+1, j)
        c3 = nd[1:, 1:]    # corner 3 = (r+1, j+1)
        c4 = nd[1:, :-1]   # corner 4 = (r, j+1)
        corners = np.stack([c1, c2, c3, c4, c1], axis=2)  # (S, R, 5, 3)
        p1 = corners[:, :, 0:4].reshape(-1, 3)
        p2 = corners[:, :, 1:5].reshape(-1, 3)
        rvc = self.rvc4.reshape(-1)
        gam = np.repeat(self.gam.reshape(-1), 4)
        flag = np.ones(gam.size, dtype=np.uint8)
        if self.F > 0:
            hp1 = nd[:-1, R]  # vf2 of last row: corner 2 -> corner 3
            hp2 = nd[1:, R]
            p1 = np.concatenate([p1, hp1, self.far_nodes[1:]])
            p2 = np.concatenate([p2, hp2, self.far_nodes[:-1]])
            rvc = np.concatenate([rvc, self.rvc4[:, R - 1, 1], self.rvcF])
            gam = np.concatenate([gam, -self.gam[:, R - 1], self.gamF])
            flag = np.concatenate([flag, np.zeros(S, np.uint8), np.ones(self.F, np.uint8)])
        return (np.ascontiguousarray(p1), np.ascontiguousarray(p2), np.ascontiguousarray(rvc),
                np.ascontiguousarray(gam), flag)

    def targets(self) -> np.ndarray:
        """Convected nodes = vind_onNwake targets (corner 2 of every ring + corner 3 of the last column,
        libCommon.f90:133-145) then vind_onFwake targets (fc(:,1) of each far filament, :190-195)."""
        t = self.nodes[:, 1:, :].reshape(-1, 3)
        if self.F > 0:
            t = np.concatenate([t, self.far_nodes[1:]])
        return np.ascontiguousarray(t)


def _helix_lattice(rng, hub, radius, chord, S, R, F, psi0, sense, dpsi_deg=5.0, gamma0=1.0, jitter=1e-3,
                   delta=5.0, nu=1.8e-5, omega=100.0):
    dpsi = np.deg2rad(dpsi_deg)
    pitch = 0.1 * radius
    th = np.linspace(0.0, np.pi, S + 1)
    rj = radius * (0.2 + 0.8 * 0.5 * (1.0 - np.cos(th)))          # cosine-spaced on [0.2, 1] R
    i = np.arange(R + 1)
    psi = psi0 + sense * i * dpsi
    nodes = np.empty((S + 1, R + 1, 3))
    nodes[:, :, 0] = hub[0] + rj[:, None] * np.cos(psi)[None, :]
    nodes[:, :, 1] = hub[1] + rj[:, None] * np.sin(psi)[None, :]
    nodes[:, :, 2] = hub[2] - pitch * (i * dpsi)[None, :] / (2 * np.pi)
    nodes += rng.uniform(-jitter * radius, jitter * radius, size=nodes.shape)
    jj = np.arange(1, S + 1)
    gam = gamma0 * np.sin(np.pi * (jj - 0.5) / S)[:, None] * (1.0 + 0.1 * rng.uniform(-1, 1, size=(S, R)))
    age = (np.arange(R) + 1) * dpsi / omega
    stream = np.sqrt((0.04 * chord) ** 2 + 4 * 1.2564 * delta * nu * age)
    rvc4 = np.empty((S, R, 4))
    rvc4[:, :, 0] = stream[None, :]
    rvc4[:, :, 2] = stream[None, :]
    rvc4[:, :, 1] = 0.04 * chord
    rvc4[:, :, 3] = 0.04 * chord
    if F > 0:
        k = np.arange(F + 1)
        psif = psi[-1] + sense * k * dpsi
        far = np.empty((F + 1, 3))
        far[:, 0] = hub[0] + radius * np.cos(psif)
        far[:, 1] = hub[1] + radius * np.sin(psif)
        far[:, 2] = nodes[-1, -1, 2] - pitch * (k * dpsi) / (2 * np.pi)
        far[0] = nodes[-1, -1]  # tip of the last near row
        far[1:] += rng.uniform(-jitter * radius, jitter * radius, size=(F, 3))
        gmin = gam[:, -1][np.argmax(np.abs(gam[:, -1]))]
        gamF = np.full(F, gmin)
        rvcF = np.full(F, stream[-1])
    else:
        far = np.zeros((0, 3))
        gamF = np.zeros(0)
        rvcF = np.zeros(0)
    return Lattice(nodes, gam, rvc4, far, gamF, rvcF, {"kind": "rotor-blade", "hub": list(map(float, hub))})


def _wing_lattice(rng, origin, span, chord, S, R, jitter=1e-3, gamma0=0.5):
    y = origin[1] + span * 0.5 * (1.0 - np.cos(np.linspace(0, np.pi, S + 1))) - span / 2
    x = origin[0] + chord * 0.25 * np.arange(R + 1)
    nodes = np.empty((S + 1, R + 1, 3))
    nodes[:, :, 0] = x[None, :]
    nodes[:, :, 1] = y[:, None]
    nodes[:, :, 2] = origin[2]
    nodes += rng.uniform(-jitter * chord, jitter * chord, size=nodes.shape)
    jj = np.arange(1, S + 1)
    gam = gamma0 * np.sin(np.pi * (jj - 0.5) / S)[:, None] * (1.0 + 0.1 * rng.uniform(-1, 1, size=(S, R)))
    rvc4 = np.full((S, R, 4), 0.04 * chord)
    return Lattice(nodes, gam, rvc4, np.zeros((0, 3)), np.zeros(0), np.zeros(0), {"kind": "wing"})


def multirotor(n_filaments: int, seed: int = 12345, n_rotor: int = 4, nb: int = 2, S: int = 32, F: int = 64,
               with_wing: bool = True) -> list[Lattice]:
    """Lattices of a 4-rotor + wing configuration with ~n_filaments filaments in total."""
    rng = np.random.Generator(np.random.PCG64(seed))
    radius, chord = 1.0, 0.1
    n_blades = n_rotor * nb
    wing_S, wing_R = (S, max(2, min(64, n_filaments // (40 * S)))) if with_wing else (0, 0)
    wing_n = 4 * wing_S * wing_R
    per_blade = max(0, n_filaments - wing_n) / max(n_blades, 1)
    R = max(2, int(round((per_blade - S - F) / (4 * S))))
    hubs = [np.array([2.5 * np.cos(2 * np.pi * k / n_rotor), 2.5 * np.sin(2 * np.pi * k / n_rotor), 0.0])
            for k in range(n_rotor)]
    out = []
    for k, hub in enumerate(hubs):
        sense = 1.0 if k % 2 == 0 else -1.0
        for b in range(nb):
            out.append(_helix_lattice(rng, hub, radius, chord, S, R, F, psi0=2 * np.pi * b / nb, sense=sense))
    if with_wing:
        out.append(_wing_lattice(rng, np.array([0.0, 0.0, 0.5]), 4.0, 0.5, wing_S, wing_R))
    return out


def flatten_all(lattices: list[Lattice]):
    parts = [l.flatten() for l in lattices]
    cat = lambda k: np.ascontiguousarray(np.concatenate([p[k] for p in parts]))
    return cat(0), cat(1), cat(2), cat(3), cat(4)


def targets_all(lattices: list[Lattice]) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([l.targets() for l in lattices]))


def random_filaments(n: int, m: int, seed: int = 0, scale: float = 1.0):
    """Unstructured random test set (plus a few targets placed exactly on filament end points / axes)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = rng.uniform(-1, 1, size=(n, 3)) * scale
    p2 = p1 + rng.uniform(-0.2, 0.2, size=(n, 3)) * scale
    rvc = rng.uniform(0.0, 0.05, size=n) * scale
    gam = rng.uniform(-1, 1, size=n)
    flag = (rng.uniform(size=n) < 0.5).astype(np.uint8)
    gam[rng.uniform(size=n) < 0.05] = 0.0
    gam[rng.uniform(size=n) < 0.02] = 1e-17
    P = rng.uniform(-1.2, 1.2, size=(m, 3)) * scale
    k = min(m, n, 16)
    if k > 0:
        P[:k] = p1[:k]                       # on an end point: c == 0 exactly
        if m > 2 * k:
            P[k:2 * k] = p2[:k]
        if m > 3 * k:
            P[2 * k:3 * k] = 0.5 * (p1[:k] + p2[:k])  # on the filament itself
    return p1, p2, rvc, gam, flag, P


# ---- the same wakes in the reference's own record layout (what the Fortran driver holds and the shim hands over) ----

VF, VR, FW = 12, 50, 13   # doubles per vf_class / vr_class (= Nwake_class) / Fwake_class record (classdef.f90:57-104, :181-220)


def lattice_records(lat: Lattice):
    """One blade's wake as the reference stores it: waN (S, R, 50) = Fortran waN(R, S) of Nwake_class records (ring corners
    per vr_assignP, classdef.f90:569-592: filament k runs corner k -> k+1; rVc0 = rVc; gam at member 48) and waF (F, 13)
    Fwake_class records (fc(:,1) = downstream end, fc(:,2) = upstream end; gam at member 12)."""
    S, R = lat.S, lat.R
    nd = lat.nodes
    c = [nd[:-1, :-1], nd[:-1, 1:], nd[1:, 1:], nd[1:, :-1]]
    waN = np.zeros((S, R, VR))
    for f in range(4):
        a, b = c[f], c[(f + 1) % 4]
        waN[:, :, VF * f + 0:VF * f + 3] = a
        waN[:, :, VF * f + 3:VF * f + 6] = b
        ln = np.linalg.norm(b - a, axis=2)
        waN[:, :, VF * f + 6] = ln          # l0
        waN[:, :, VF * f + 7] = ln          # lc
        waN[:, :, VF * f + 8] = lat.rvc4[:, :, f]
        waN[:, :, VF * f + 9] = lat.rvc4[:, :, f]
    waN[:, :, 48] = lat.gam
    waF = np.zeros((lat.F, FW))
    if lat.F > 0:
        waF[:, 0:3] = lat.far_nodes[1:]
        waF[:, 3:6] = lat.far_nodes[:-1]
        ln = np.linalg.norm(lat.far_nodes[1:] - lat.far_nodes[:-1], axis=1)
        waF[:, 6] = ln
        waF[:, 7] = ln
        waF[:, 8] = lat.rvcF
        waF[:, 9] = lat.rvcF
        waF[:, 12] = lat.gamF
    return waN, waF


def rotors_from_lattices(lattices: list[Lattice]) -> list[dict]:
    """Group the blades' lattices into rotors (blades of one hub; the fixed wing is a one-bladed rotor without far wake):
    [{nb, ns, nNwake, nFwake, waN: [per blade (ns, nNwake, 50)], waF: [per blade (nFwake, 13)], lattices}] in the order
    of flatten_all / targets_all, so that source and target counts are those of the flat enumeration."""
    rotors, key_of = [], {}
    for l in lattices:
        key = (l.meta.get("kind"), tuple(l.meta.get("hub", ())), l.R, l.S, l.F)
        if l.meta.get("kind") != "rotor-blade" or key not in key_of:
            key_of[key] = len(rotors)
            rotors.append({"nb": 0, "ns": l.S, "nNwake": l.R, "nFwake": l.F, "waN": [], "waF": [], "lattices": []})
        r = rotors[key_of[key]]
        waN, waF = lattice_records(l)
        r["nb"] += 1
        r["waN"].append(waN)
        r["waF"].append(waF)
        r["lattices"].append(l)
    return rotors
