#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the reference tree (run in the build container,
where /root/reference exists; the GPU box only reads the committed JSON files).

  python tests/golden/make_fixtures.py [/root/reference]

For every case: the namelist inputs (config.nml, geomNN.nml -> plain dicts keyed by the namelist variable
names; a PLOT3D geometry file becomes the array "grid") and, when the reference ships them, its golden
results (referenceResults/*.ref: the rNNForceNonDim.csv history and the sectional distribution).
The unit-test known answers (tests/*_test.f90) are transcribed in tests/refgeom.py / tests/test_oracle_case.py
next to the citation of the line they come from.
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def parse_namelist(path: Path) -> dict:
    """Minimal Fortran-namelist reader: `key = v1, v2, ...` lines inside &GROUP ... / blocks; `!` comments."""
    out: dict = {}
    group = None
    for raw in path.read_text().splitlines():
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
        items = [v.strip() for v in val.rstrip(",").split(",") if v.strip()]
        vals = []
        for it in items:
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
    """rotor_plot3dtoblade (classdef.f90:3957-3984): nx ny nz then x(i,j), y(i,j), z(i,j), i fastest.
    Returned flat in the order grid(3, nx, ny) column-major (xyz fastest)."""
    tok = path.read_text().split()
    nx, ny, nz = int(tok[0]), int(tok[1]), int(tok[2])
    assert nz == 1 and nx == nc + 1 and ny == ns + 1, (nx, ny, nz, nc, ns)
    a = np.array([float(t) for t in tok[3:3 + 3 * nx * ny]]).reshape(3, ny, nx)   # [comp, j, i]
    grid = np.transpose(a, (1, 2, 0))                                              # [j, i, comp] = (3, nx, ny) col-major
    return [float(x) for x in grid.reshape(-1)]


def case_fixture(case_dir: Path, name: str, n_rotors: int = 1, results: bool = True) -> dict:
    cfg = parse_namelist(case_dir / "config.nml")
    cfg.pop("fileFormatVersion", None)
    geoms = []
    for ir in range(1, n_rotors + 1):
        g = parse_namelist(case_dir / f"geom{ir:02d}.nml")
        g.pop("fileFormatVersion", None)
        gf = str(g.get("geometryFile", "0"))
        if not gf.startswith("0"):
            g["grid"] = read_plot3d(case_dir / gf, int(g["nc"]), int(g["ns"]))
        geoms.append(g)
    fx = {"name": name, "source": f"{case_dir.relative_to(case_dir.parents[1])}", "config": cfg, "geom": geoms}
    ref = case_dir / "referenceResults"
    if results and ref.exists():
        hist = np.loadtxt(ref / "r01ForceNonDim.csv.ref", skiprows=1)
        fx["ref_ForceNonDim"] = {"columns": ["iter", "CL/CT", "CD/CQ", "CLu", "CDi", "CD0", "CDu", "CFx", "CFy", "CFz"],
                                 "rows": hist.tolist(), "format": "E15.7 (libPostprocess.f90:838)"}
        dist = sorted(ref.glob("r01b01ForceDist*.csv.ref"))
        if dist:
            lines = dist[0].read_text().splitlines()
            cols = lines[0].split()
            rows = [[float(x) for x in l.split()] for l in lines[1:] if l.strip()]
            fx["ref_ForceDist"] = {"file": dist[0].name, "columns": cols, "rows": rows}
            # every sectional distribution the reference ships (libPostprocess.f90:849-881, format 16(E15.7)), by time step
            fx["ref_ForceDists"] = []
            for f in dist:
                ls = f.read_text().splitlines()
                fx["ref_ForceDists"].append({"file": f.name, "iter": int(re.search(r"ForceDist(\d+)", f.name).group(1)),
                                             "columns": ls[0].split(),
                                             "rows": [[float(x) for x in l.split()] for l in ls[1:] if l.strip()]})
        pj = ref / "r01Params.json.ref"
        if pj.exists():
            fx["ref_Params"] = json.loads(pj.read_text())
    return fx


def check_formatter(ref: Path):
    """oracle/casefile.py's E15.7 writer must reproduce the reference's golden text byte for byte."""
    sys.path.insert(0, str(HERE.parent.parent))
    from oracle import casefile
    for f in ref.glob("tests/*.case/referenceResults/r01ForceNonDim.csv.ref"):
        lines = f.read_text().splitlines()
        assert lines[0] == casefile.HEADER, f
        for l in lines[1:]:
            vals = [float(l[5 + 15 * k:5 + 15 * (k + 1)]) for k in range(9)]
            assert casefile.force_nondim_line(int(l[:5]), vals) == l, (f, l)
        print(f"formatter round-trips {f.relative_to(ref)} ({len(lines) - 1} rows)")


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    check_formatter(ref)
    jobs = [
        (ref / "tests" / "katzNplotkin-AR04.case", "katzNplotkin_AR04", True),
        (ref / "tests" / "elevateTest.case", "elevateTest", True),
        (ref / "tutorials" / "simplewing.case", "simplewing", False),
        (ref / "tutorials" / "caradonna.case", "caradonna", False),
        (ref / "tutorials" / "tr1208.case", "tr1208", False),
    ]
    for d, name, res in jobs:
        if not d.exists():
            print("skip", d)
            continue
        fx = case_fixture(d, name, results=res)
        out = HERE / f"{name}.json"
        out.write_text(json.dumps(fx, indent=None, separators=(",", ":")) + "\n")
        print(f"wrote {out} ({out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
