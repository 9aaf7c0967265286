"""Hand-built flat-wing geometry following rotor_init for geometryFile='0' (src/classdef.f90:3152-3247 panels,
:3311-3367 vortex rings at quarter-panel shift, :3388-3400 TE shed, :3402-3417 CP/nCap, :3498-3519 core radii).
Used to pin the oracle's pair kernel against the reference's AIC known-answer tests before the full case
driver is involved (tests/wing1x3_test.f90:83-85, tests/wing1x2_test.f90:164-165)."""
import numpy as np

WP, VR, VF = 104, 50, 12


def cosspace(a, b, n):  # libMath.f90:159-171
    th = np.arange(n) * (np.pi / (n - 1))
    return a + (b - a) * 0.5 * (1.0 - np.cos(th))


def linspace(a, b, n):  # libMath.f90:138-157
    dx = (b - a) / (n - 1)
    return np.arange(n) * dx + a


def flat_wing_records(nc, ns, chord, span, root_cut, velBody, dt, spanwiseCore, spanSpacing=2, chordSpacing=1):
    """wingpanel records (ns, nc, 104) of an un-pitched fixed wing (Omega = 0)."""
    xVec = linspace(-chord, 0.0, nc + 1) if chordSpacing == 1 else cosspace(-chord, 0.0, nc + 1)
    yVec = (cosspace if spanSpacing == 2 else linspace)(root_cut * span, span, ns + 1)
    rec = np.zeros((ns, nc, WP))
    PC = np.zeros((ns, nc, 4, 3))
    for j in range(ns):
        for i in range(nc):
            PC[j, i, 0] = [xVec[i], yVec[j], 0.0]
            PC[j, i, 1] = [xVec[i + 1], yVec[j], 0.0]
            PC[j, i, 2] = [xVec[i + 1], yVec[j + 1], 0.0]
            PC[j, i, 3] = [xVec[i], yVec[j + 1], 0.0]
    corners = np.zeros((ns, nc, 4, 3))
    for j in range(ns):
        for i in range(nc):
            if i < nc - 1:
                xs = [(PC[j, i, 1, 0] - PC[j, i, 0, 0]) * 0.25, (PC[j, i + 1, 1, 0] - PC[j, i, 1, 0]) * 0.25,
                      (PC[j, i + 1, 2, 0] - PC[j, i, 2, 0]) * 0.25, (PC[j, i, 2, 0] - PC[j, i, 3, 0]) * 0.25]
            else:
                xs = [(PC[j, i, 1, 0] - PC[j, i, 0, 0]) * 0.25, 0.0, 0.0, (PC[j, i, 2, 0] - PC[j, i, 3, 0]) * 0.25]
            for n in range(4):
                corners[j, i, n] = PC[j, i, n] + [xs[n], 0.0, 0.0]
    velShed = 0.3 * np.linalg.norm(velBody)          # :3394
    corners[:, nc - 1, 1, 0] += velShed * dt          # :3397-3398 (sign(1,Omega=0) = +1)
    corners[:, nc - 1, 2, 0] += velShed * dt
    dx = np.linalg.norm(PC[:, :, 1] - PC[:, :, 0], axis=-1)
    dy = np.linalg.norm(PC[:, :, 2] - PC[:, :, 1], axis=-1)
    dxdymin = min(dx.min(), dy.min())
    core = min(spanwiseCore * chord, dxdymin * 0.1)   # :3499-3502 (spanwiseCore already * chord, :3133)
    for j in range(ns):
        for i in range(nc):
            r = rec[j, i]
            for f in range(4):
                r[VF * f + 0:VF * f + 3] = corners[j, i, f]
                r[VF * f + 3:VF * f + 6] = corners[j, i, (f + 1) % 4]
                r[VF * f + 8] = core
                r[VF * f + 9] = core
            if i == nc - 1:
                r[VF * 1 + 8] = r[VF * 1 + 9] = spanwiseCore * chord   # :3513
            r[52:64] = PC[j, i].reshape(-1)
            cp = ((PC[j, i, 0] + PC[j, i, 3]) * 0.25 + (PC[j, i, 1] + PC[j, i, 2]) * 0.75) * 0.5  # :794-795
            n = np.cross(PC[j, i, 2] - PC[j, i, 0], PC[j, i, 3] - PC[j, i, 1])                    # :812-813
            r[64:67] = cp
            r[67:70] = n / np.linalg.norm(n)
    return rec


AIC_WING1X3 = np.array([[3.18902463326559, -0.135541879282409, -0.002934863377599825],
                        [-0.02762714554230721, 2.97991717943654, -0.02762714554230732],
                        [-0.002934863377599823, -0.135541879282409, 3.18902463326559]])  # wing1x3_test.f90:83-85
AIC_WING1X2 = np.array([[1.1223476, -0.092667149], [-0.092667149, 1.1223476]])          # wing1x2_test.f90:164-165


def wing1x3():
    """tests/wing1x3_test.f90:17-78"""
    return flat_wing_records(nc=1, ns=3, chord=0.3, span=2.0, root_cut=0.0, velBody=[-6.0, 0, 0], dt=0.00625,
                             spanwiseCore=0.04)


def wing1x2():
    """tests/wing1x2_test.f90:17-77"""
    return flat_wing_records(nc=1, ns=2, chord=1.0, span=2.0, root_cut=0.0, velBody=[-10.0, 0, 0], dt=0.00625,
                             spanwiseCore=0.04)
