"""File formats around the Eulerian grid tool (caller side): filamentsNNNNN.dat as filaments2file
writes it (src/libPostprocess.f90:455-464), gridconfig.nml (src/gridgen.f90:27-40), gridNNNNN.tec (:150-164); and the
program itself (volcanor_b200.gridgen.run) on a case directory, on the GPU."""
import json
import struct
from pathlib import Path

import numpy as np
import pytest

from volcanor_b200 import gridgen as gg

GOLDEN = Path(__file__).resolve().parent / "golden"
GRIDCONFIG = """&VERSION
fileFormatVersion = 0.2
/
! -----------------------------------------------
&INPUTS
! NO. OF NODES FOR GRID
nx = 9
ny = 7
nz = 5
xyzMin = -1.5, -1.5, -1.2
xyzMax = 1.5, 1.5d0, 0.3
vel = 0.0, -6.0, -3.0
! RANGE OF FILAMENT FILES TO BE READ
fileRangeStart = 24
fileRangeStep = 1
fileRangeEnd = 24
/
"""


def _arrays(seed=3, n=(6, 20, 4, 5)):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n[0], 50)), rng.standard_normal((n[1], 50)), rng.standard_normal((n[2], 12)),
            rng.standard_normal(n[2]), rng.standard_normal((n[3], 12)), rng.standard_normal(n[3]))


def test_filaments_file_layout_and_round_trip(tmp_path):
    a = _arrays()
    p = tmp_path / "filaments00024.dat"
    gg.write_filaments(p, *a)
    raw = p.read_bytes()
    # gfortran unformatted sequential: <len> body <len>; four default-integer records first
    assert struct.unpack_from("<iii", raw, 0) == (4, 6, 4) and struct.unpack_from("<iii", raw, 12) == (4, 20, 4)
    assert struct.unpack_from("<iii", raw, 24) == (4, 4, 4) and struct.unpack_from("<iii", raw, 36) == (4, 5, 4)
    assert struct.unpack_from("<i", raw, 48)[0] == 8 * 50 * (6 + 20)
    assert len(raw) == 4 * 12 + (8 + 8 * 50 * 26) + (8 + 8 * (12 * 4 + 4)) + (8 + 8 * (12 * 5 + 5))
    assert np.frombuffer(raw, dtype="<f8", count=50, offset=52).tolist() == a[0][0].tolist()   # vrWing(1) leads record 5
    f = gg.read_filaments(p)
    for got, want in zip((f["vrWing"], f["vrNwake"], f["vfNwakeTE"], f["gamNwakeTE"], f["vfFwake"], f["gamFwake"]), a):
        assert np.array_equal(got, want)
    # empty far wake / corrupted marker
    gg.write_filaments(p, a[0], a[1], a[2], a[3], np.zeros((0, 12)), np.zeros(0))
    assert gg.read_filaments(p)["vfFwake"].shape == (0, 12)
    p.write_bytes(raw[:60] + b"\x00" + raw[61:-4] + struct.pack("<i", 1))
    with pytest.raises(ValueError):
        gg.read_filaments(p)


def test_gridconfig_and_tecplot(tmp_path):
    (tmp_path / "gridconfig.nml").write_text(GRIDCONFIG)
    cfg = gg.read_gridconfig(tmp_path / "gridconfig.nml")
    assert (cfg["nx"], cfg["ny"], cfg["nz"]) == (9, 7, 5) and cfg["xyzMax"] == [1.5, 1.5, 0.3] and cfg["vel"] == [0.0, -6.0, -3.0]
    assert (cfg["fileRangeStart"], cfg["fileRangeStep"], cfg["fileRangeEnd"]) == (24, 1, 24)
    (tmp_path / "bad.nml").write_text(GRIDCONFIG.replace("0.2", "0.1"))
    with pytest.raises(ValueError, match="template version"):
        gg.read_gridconfig(tmp_path / "bad.nml")
    (tmp_path / "inv.nml").write_text(GRIDCONFIG.replace("xyzMin = -1.5", "xyzMin = 2.5"))
    with pytest.raises(ValueError, match="XYZmin"):
        gg.read_gridconfig(tmp_path / "inv.nml")
    rng = np.random.default_rng(0)
    v = rng.standard_normal((4, 6, 8, 3))
    out = tmp_path / "grid00024.tec"
    gg.write_tecplot(out, 9, 7, 5, cfg["xyzMin"], cfg["xyzMax"], v)
    head = out.read_text().splitlines()[:5]
    assert head[0].strip() == 'TITLE = "Grid"' and head[2].split()[:4] == ["ZONE", "I=9", "J=7", "K=5"]
    assert head[4].strip() == "VARLOCATION=([4]=CELLCENTERED,[5]=CELLCENTERED,[6]=CELLCENTERED)"
    nx, ny, nz, nodes, vel = gg.read_tecplot(out)
    assert (nx, ny, nz) == (9, 7, 5) and np.array_equal(vel, v)
    assert np.array_equal(nodes, gg.grid_nodes(9, 7, 5, cfg["xyzMin"], cfg["xyzMax"]))
    assert nodes[0, 0, 1, 0] == 1 * (3.0 / 8) + -1.5 and nodes[-1, -1, -1].tolist() == [8 * (3.0 / 8) - 1.5, 6 * (3.0 / 6) - 1.5, 4 * (1.5 / 4) - 1.2]
    with pytest.raises(FileExistsError):                    # status='new' (gridgen.f90:147)
        gg.write_tecplot(out, 9, 7, 5, cfg["xyzMin"], cfg["xyzMax"], v)


def _short_caradonna_with_far_wake(oracle, nsteps=24):
    fx = json.loads((GOLDEN / "caradonna.json").read_text())
    fx["config"]["nt"] = 40
    fx["geom"][0]["nNwake"] = 12
    fx["geom"][0]["wakeTruncateNt"] = 18
    c = oracle.Case(fx)
    c.init()
    for _ in range(nsteps):
        c.step()
    return c


def test_filaments_from_case_follows_filaments2file(oracle):
    from oracle import casefile
    c = oracle.Case(json.loads((GOLDEN / "katzNplotkin_AR04.json").read_text()))
    c.init()
    with pytest.raises(RuntimeError, match="development of far wake"):        # libPostprocess.f90:383-385
        casefile.filaments_from_case(c)
    c = _short_caradonna_with_far_wake(oracle)
    r = c.rotor(0)
    f = casefile.filaments_from_case(c)
    assert f["vrWing"].shape == (2 * 10 * 25, 50) and f["vrNwake"].shape == (2 * 12 * 25, 50)
    assert f["vfNwakeTE"].shape == (2 * 25, 12) and f["vfFwake"].shape == (2 * 6, 12)
    assert np.array_equal(f["vrWing"][10 * 25 + 3], r.wiP(1)[0, 3, :50])       # blade 2, icol 1, irow 4
    assert np.array_equal(f["vrNwake"][12 * 2 + 5], r.waN(0)[2, 5])            # blade 1, icol 3, irow 6
    assert np.array_equal(f["vfNwakeTE"][25 + 4], r.waN(1)[4, 11, 12:24]) and f["gamNwakeTE"][25 + 4] == -r.waN(1)[4, 11, 48]
    assert np.array_equal(f["vfFwake"][6 + 2], r.waF(1)[2, :12]) and f["gamFwake"][6 + 2] == r.waF(1)[2, 12]


@pytest.mark.gpu
def test_gridgen_program_on_a_case_directory(ctx, oracle, tmp_path):
    """filaments2file -> Results/filaments00024.dat -> `program gridgen` on the GPU -> Results/grid00024.tec, against
    the oracle's restatement of the program on the same arrays."""
    from oracle import casefile
    c = _short_caradonna_with_far_wake(oracle)
    f = casefile.filaments_from_case(c)
    (tmp_path / "Results").mkdir()
    (tmp_path / "gridconfig.nml").write_text(GRIDCONFIG)
    gg.write_filaments(tmp_path / "Results" / "filaments00024.dat", f["vrWing"], f["vrNwake"], f["vfNwakeTE"],
                       f["gamNwakeTE"], f["vfFwake"], f["gamFwake"])
    written = gg.run(tmp_path, ctx=ctx, verbose=False)
    assert [p.name for p in written] == ["grid00024.tec"]
    nx, ny, nz, nodes, vel = gg.read_tecplot(written[0])
    cfg = gg.read_gridconfig(tmp_path / "gridconfig.nml")
    _, vo = oracle.gridgen(nx, ny, nz, np.array(cfg["xyzMin"]), np.array(cfg["xyzMax"]), np.array(cfg["vel"]), f["vrWing"],
                           f["vrNwake"], f["vfNwakeTE"], f["gamNwakeTE"], f["vfFwake"], f["gamFwake"])
    scale = np.max(np.abs(vo - np.array(cfg["vel"])))
    err = np.max(np.abs(vel - vo)) / max(scale, 1.0)
    print(f"gridgen on a case directory: {vel.size // 3} cell centres, max scaled error vs the oracle {err:.2e}")
    assert err < 1e-12 * 50
