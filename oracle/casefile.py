"""TEST INFRASTRUCTURE: what the oracle's case driver needs around the reference's case directories.  The file formats
themselves (namelists, PLOT3D, the force history row) are read / written by volcanor_b200/casefile.py; here only
`filaments_from_case`, which gathers the filaments file's contents from the oracle's rotors."""
from __future__ import annotations

import numpy as np

from volcanor_b200.casefile import (HEADER, force_nondim_line, fortran_e15_7, parse_namelist, read_case,  # noqa: F401
                                    read_plot3d)


def filaments_from_case(case) -> dict:
    """What filaments2file gathers from rotor(:) (libPostprocess.f90:363-453), in its loop order: the wing's vortex
    rings, ALL nNwake rows of the near wake, vf(2) of the last near row with -gam, the active far filaments.  Like the
    reference it refuses to run before the far wake has developed (:383-385)."""
    vrW, vrN, vfT, gT, vfF, gF = [], [], [], [], [], []
    rotors = [case.rotor(ir) for ir in range(case.nr)]
    for r in rotors:
        d = r.dims()
        if d["rowFar"] > r.nFwake:
            raise RuntimeError("ERROR: Use filaments2file() only after development of far wake")
    for r in rotors:
        for ib in range(r.nb):
            vrW.append(r.wiP(ib)[:, :, :50].reshape(-1, 50))          # icol outer, irow inner = storage order
    for r in rotors:
        for ib in range(r.nb):
            vrN.append(r.waN(ib).reshape(-1, 50))
    for r in rotors:
        for ib in range(r.nb):
            last = r.waN(ib)[:, r.nNwake - 1, :]
            vfT.append(last[:, 12:24])
            gT.append(last[:, 48] * (-1.0))
    for r in rotors:
        rowFar = r.dims()["rowFar"]
        for ib in range(r.nb):
            vfF.append(r.waF(ib)[rowFar - 1:, :12])
            gF.append(r.waF(ib)[rowFar - 1:, 12])
    cat = lambda xs, w: np.concatenate(xs).copy() if xs else np.zeros((0, w) if w > 1 else (0,))
    return {"vrWing": cat(vrW, 50), "vrNwake": cat(vrN, 50), "vfNwakeTE": cat(vfT, 12), "gamNwakeTE": cat(gT, 1),
            "vfFwake": cat(vfF, 12), "gamFwake": cat(gF, 1)}
