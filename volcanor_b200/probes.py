"""The caller side of the reference's velocity probes (`switches%probe`, src/main.f90:83-93, :792-797;
`probes2file`, src/libPostprocess.f90:333-361): `probes.in` reader, the probe velocities through the batched point
calls of the C ABI, and `Results/probesNNNNN.csv` in the reference's format; the blade inflow of `inflow2file`
(libPostprocess.f90:900-945) through the same calls.

    vel(:, i) = probeVel(:, i) + sum over rotors of [vind_bywing(P_i) + vind_bywake(P_i)],   P_i = probe(:, i) + probeVel(:, i)*t

The reference evaluates the two `rotor%vind_*` functions point by point; here every rotor takes ALL probe locations in
one `vlc_rotor_vind_bywing` / `vlc_rotor_vind_bywake` call each (the same sweeps as the collocation-point loops) and the
terms are added in the reference's order.  Host-side numpy only around the library calls; no CPU path for the velocities.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def read_probes(path) -> tuple[np.ndarray, np.ndarray]:
    """`probes.in` (main.f90:85-92, list-directed reads): the count, then one line `x y z u v w` per probe.
    Returns (probe, probeVel), each (n, 3)."""
    tok = Path(path).read_text().replace(",", " ").split()
    n = int(tok[0])
    vals = np.array([float(t.replace("d", "e").replace("D", "e")) for t in tok[1:1 + 6 * n]], dtype=np.float64)
    if vals.size != 6 * n:
        raise ValueError(f"{path}: {n} probes announced, {vals.size // 6} complete lines found")
    vals = vals.reshape(n, 6)
    return vals[:, :3].copy(), vals[:, 3:].copy()


def probe_velocities(ctx, n_rotors: int, probe, probeVel, t: float) -> tuple[np.ndarray, np.ndarray]:
    """(vel, probeLocation), each (n, 3), from the rotors 0..n_rotors-1 the context holds (current wake)."""
    probe = np.ascontiguousarray(probe, dtype=np.float64).reshape(-1, 3)
    probeVel = np.ascontiguousarray(probeVel, dtype=np.float64).reshape(-1, 3)
    loc = probe + probeVel * t                      # libPostprocess.f90:347
    vel = probeVel.copy()
    for ir in range(n_rotors):                      # :349-353, terms added left to right
        vel = (vel + ctx.rotor_vind_bywing(ir, loc)) + ctx.rotor_vind_bywake(ir, loc)
    return vel, loc


def inflow_velocities(ctx, n_rotors: int, secCP, directionVector) -> np.ndarray:
    """The inflow along `directionVector` at the section points of one rotor's blades (`inflow2file`,
    libPostprocess.f90:900-928): secCP (nb, ns, 3) -> inflowVel (nb, ns) = sum over rotors of
    dot(vind_bywing - vind_bywing_boundVortices + vind_bywake, directionVector), terms added in the reference's order."""
    secCP = np.ascontiguousarray(secCP, dtype=np.float64)
    P = secCP.reshape(-1, 3)
    d = np.asarray(directionVector, dtype=np.float64).reshape(3)

    def dot(v):  # dot_product of each row with d, unfused
        return (v[:, 0] * d[0] + v[:, 1] * d[1]) + v[:, 2] * d[2]
    inflow = np.zeros(P.shape[0])
    for ir in range(n_rotors):                      # :913-921
        inflow = inflow + dot(ctx.rotor_vind_bywing(ir, P))
        inflow = inflow - dot(ctx.rotor_vind_bywing_boundVortices(ir, P))
        inflow = inflow + dot(ctx.rotor_vind_bywake(ir, P))
    return inflow.reshape(secCP.shape[:-1])


def fortran_e(x: float, width: int = 15, digits: int = 7) -> str:
    """One value in Fortran Ew.d (E15.7): [-]0.dddddddE+xx right-justified."""
    x = float(x)
    if x == 0.0:
        return ("0." + "0" * digits + "E+00").rjust(width)
    if not np.isfinite(x):
        return ("NaN" if x != x else ("Infinity" if x > 0 else "-Infinity")).rjust(width)
    mant, exp = f"{abs(x):.{digits - 1}E}".split("E")      # correctly rounded d.ddddddE+xx
    e = int(exp) + 1
    body = f"0.{mant.replace('.', '')}E{'+' if e >= 0 else '-'}{abs(e):02d}"
    return (("-" if x < 0 else "") + body).rjust(width)


def write_probes(path, vel, loc) -> None:
    """`probesNNNNN.csv` (libPostprocess.f90:343-358): header 6(A15), one row 6(E15.7) per probe: u v w x y z."""
    rows = ["".join(h.rjust(15) for h in ("u", "v", "w", "x", "y", "z"))]
    for v, p in zip(np.asarray(vel).reshape(-1, 3), np.asarray(loc).reshape(-1, 3)):
        rows.append("".join(fortran_e(a) for a in (*v, *p)))
    Path(path).write_text("\n".join(rows) + "\n")


def probes2file(ctx, n_rotors: int, results_dir, timestamp: str, probe, probeVel, t: float) -> Path:
    """= call probes2file(timestamp, probe, probeVel, rotor, t) (main.f90:795)."""
    vel, loc = probe_velocities(ctx, n_rotors, probe, probeVel, t)
    out = Path(results_dir) / f"probes{timestamp}.csv"
    write_probes(out, vel, loc)
    return out
