"""The reference's case-directory formats on the caller side of the path: readers for `config.nml` /
`geomNN.nml` and PLOT3D geometry files, and the writer of one row of the force history file.

`read_case(dir)` parses `config.nml` / `geomNN.nml` (Fortran namelists, libCommon.f90:51-108, classdef.f90:2541-2767;
a PLOT3D `geometryFile` becomes the in-memory `grid`, rotor_plot3dtoblade classdef.f90:3957-3984) into a dict
{"name", "config": {...}, "geom": [{...}, ...]} keyed by the namelist variable names.  `force_nondim_line` formats one
row of `Results/rNNForceNonDim.csv` exactly as force2file does (libPostprocess.f90:824-838: `format(A, 9(E15.7))`); the
golden files of the reference round-trip byte for byte (tests/golden/make_fixtures.py:check_formatter).
File formats only: the program flow that consumes them is the reference driver's (main.f90), not this library's.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

HEADER = " iter          CL/CT          CD/CQ            CLu            CDi            CD0            CDu            CFx            CFy            CFz"


def parse_namelist(path: Path) -> dict:
    """Minimal Fortran-namelist reader: `key = v1, v2, ...` lines inside &GROUP ... / blocks; `!` starts a comment."""
    out: dict = {}
    group = None
    for raw in Path(path).read_text().splitlines():
        line = raw.split("!")[0].strip()
        if not line:
            continue
        if line.startswith("&"):
            group = line[1:].strip()
            continue
        if line == "/":
            group = None
            continue
        if "=" not in line or group is None:
            continue
        key, val = [s.strip() for s in line.split("=", 1)]
        vals = []
        for it in [v.strip() for v in val.rstrip(",").split(",") if v.strip()]:
            it = it.strip("'\"")
            try:
                vals.append(int(it))
            except ValueError:
                try:
                    vals.append(float(it.replace("d", "e").replace("D", "e")))
                except ValueError:
                    vals.append(it)
        out[key] = vals[0] if len(vals) == 1 else vals
    return out


def read_plot3d(path: Path, nc: int, ns: int) -> list:
    """nx ny nz, then x(i,j), y(i,j), z(i,j) with i fastest; returned flat as grid(3, nx, ny) column-major."""
    tok = Path(path).read_text().split()
    nx, ny, nz = int(tok[0]), int(tok[1]), int(tok[2])
    if nz != 1 or nx != nc + 1 or ny != ns + 1:
        raise ValueError("ERROR: Wrong or conflicting data in PLOT3D file")       # classdef.f90:3973-3977
    a = np.array([float(t) for t in tok[3:3 + 3 * nx * ny]]).reshape(3, ny, nx)    # [comp, j, i]
    return [float(x) for x in np.transpose(a, (1, 2, 0)).reshape(-1)]


def read_case(case_dir) -> dict:
    case_dir = Path(case_dir)
    cfg = parse_namelist(case_dir / "config.nml")
    if str(cfg.pop("fileFormatVersion", "0.5")) not in ("0.5",):
        raise ValueError("ERROR: config.nml template version does not match")     # libCommon.f90:71-73
    geoms = []
    for ir in range(1, int(cfg.get("nr", 1)) + 1):
        g = parse_namelist(case_dir / f"geom{ir:02d}.nml")
        g.pop("fileFormatVersion", None)
        gf = str(g.get("geometryFile", "0"))
        if not gf.startswith("0"):
            g["grid"] = read_plot3d(case_dir / gf, int(g["nc"]), int(g["ns"]))
        geoms.append(g)
    return {"name": case_dir.name, "config": cfg, "geom": geoms}


def fortran_e15_7(x: float) -> str:
    """One value in Fortran E15.7: 0.dddddddE+xx right-justified in 15 columns (correctly rounded from the exact
    binary value, like gfortran's formatted write)."""
    if x == 0.0 or not math.isfinite(x):
        return "  0.0000000E+00" if x == 0.0 else f"{x:>15}"
    mant, exp = f"{abs(x):.6E}".split("E")
    e = int(exp) + 1
    return f"{'-' if x < 0 else ''}0.{mant.replace('.', '')}E{'+' if e >= 0 else '-'}{abs(e):02d}".rjust(15)


def force_nondim_line(it: int, f) -> str:
    return f"{it:05d}" + "".join(fortran_e15_7(float(v)) for v in f)
