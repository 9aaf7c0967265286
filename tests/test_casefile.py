"""The caller side of the path: reading the reference's case directories and writing its force history
file.  The rows the driver restatement writes must be TEXTUALLY identical to the reference's golden
r01ForceNonDim.csv.ref (format(A, 9(E15.7)), libPostprocess.f90:838) wherever the values agree to the printed digit."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import casefile

GOLDEN = Path(__file__).resolve().parent / "golden"


def _ref_lines(fx):
    """The golden file as text, rebuilt from the fixture's numbers (the formatter round-trips the reference's own text
    exactly: checked against /root/reference by tests/golden/make_fixtures.py when the fixtures are made)."""
    return [casefile.HEADER] + [casefile.force_nondim_line(int(r[0]), r[1:]) for r in fx["ref_ForceNonDim"]["rows"]]


def _write_namelists(fx, d: Path):
    """An input deck in the reference's namelist syntax from a fixture dict."""
    def block(name, kv):
        rows = [f"&{name}"]
        for k, v in kv.items():
            if k == "grid":
                continue
            if isinstance(v, list):
                v = ", ".join(repr(x) for x in v)
            rows.append(f"{k} = {v}   ! comment")
        return "\n".join(rows + ["/", ""])
    (d / "config.nml").write_text(block("VERSION", {"fileFormatVersion": 0.5}) + block("PARAMS", fx["config"]))
    for ir, g in enumerate(fx["geom"], 1):
        g = dict(g)
        if "grid" in g:
            grid = np.array(g["grid"]).reshape(g["ns"] + 1, g["nc"] + 1, 3)       # [j, i, comp]
            txt = f"{g['nc'] + 1} {g['ns'] + 1} 1\n" + "\n".join(
                " ".join(repr(float(grid[j, i, c])) for i in range(g["nc"] + 1)) for c in range(3) for j in range(g["ns"] + 1))
            (d / "blade.xyz").write_text(txt + "\n")
            g["geometryFile"] = "blade.xyz"
        (d / f"geom{ir:02d}.nml").write_text(block("VERSION", {"fileFormatVersion": 0.15}) + block("ALL", g))


def test_fortran_e15_7_formatter():
    assert casefile.fortran_e15_7(0.2047732e1) == "  0.2047732E+01"
    assert casefile.fortran_e15_7(-0.6325490e-2) == " -0.6325490E-02"
    assert casefile.fortran_e15_7(0.0) == "  0.0000000E+00"
    assert casefile.fortran_e15_7(0.99999996) == "  0.1000000E+01"       # rounding carries into the next decade
    assert casefile.fortran_e15_7(1.951962e-4) == "  0.1951962E-03"
    assert casefile.force_nondim_line(7, [0.1] * 9).startswith("00007  0.1000000E+00")


@pytest.mark.parametrize("name,nsteps", [("katzNplotkin_AR04", 40), ("elevateTest", 40)])
def test_case_directory_round_trip_and_golden_text(oracle, tmp_path, name, nsteps):
    fx = json.loads((GOLDEN / f"{name}.json").read_text())
    _write_namelists(fx, tmp_path)
    fx2 = casefile.read_case(tmp_path)
    assert fx2["config"]["nt"] == fx["config"]["nt"] and fx2["geom"][0]["ns"] == fx["geom"][0]["ns"]
    if "grid" in fx["geom"][0]:
        assert np.array_equal(np.array(fx2["geom"][0]["grid"]), np.array(fx["geom"][0]["grid"]))
    c = oracle.Case(fx2)
    c.init()
    lines = [casefile.HEADER, casefile.force_nondim_line(0, c.force_nondim(0))]
    for it in range(1, nsteps + 1):
        c.step()
        lines.append(casefile.force_nondim_line(it, c.force_nondim(0)))
    ref = _ref_lines(fx)[:nsteps + 2]
    same = sum(a == b for a, b in zip(lines, ref))
    # CL/CT, CFx, CFz columns (1, 7, 9) textually; a last-digit flip is allowed on at most 2 % of the rows
    cols = lambda l: (l[5:20], l[95:110], l[125:140])
    same_cols = sum(cols(a) == cols(b) for a, b in zip(lines, ref))
    assert same_cols >= 0.98 * len(ref), (same_cols, len(ref))
    print(f"{name}: {same}/{len(ref)} rows byte-identical to the golden file, {same_cols}/{len(ref)} in the CL/CFx/CFz columns")
