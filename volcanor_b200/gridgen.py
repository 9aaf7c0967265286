"""The caller side of `program gridgen` (reference src/gridgen.f90): the file formats around vlc_gridgen.

  * `read_filaments` / `write_filaments`: `Results/filamentsNNNNN.dat`, the unformatted sequential file that
    filaments2file writes (src/libPostprocess.f90:363-473) and gridgen reads (src/gridgen.f90:93-111).  Seven records,
    each framed by gfortran's 4-byte little-endian length markers:
        nvrWing | nvrNwake | nvfNwakeTE | nvfFwake            (default integers, 4 bytes)
        vrWing(:), vrNwake(:)                                 (vr_class = 50 doubles each, classdef.f90:81-104)
        vfNwakeTE(:), gamNwakeTE(:)                           (vf_class = 12 doubles each, :57-79; then the doubles)
        vfFwake(:), gamFwake(:)
  * `read_gridconfig`: `gridconfig.nml` (namelists VERSION and INPUTS, gridgen.f90:27-40; template version 0.2).
  * `write_tecplot`: `Results/gridNNNNN.tec` with the reference's header lines and BLOCK data order
    (gridgen.f90:150-164): X, Y, Z at the nodes, U, V, W at the cell centres, x fastest.  The reference writes the
    numbers list-directed (`write(13, *)`), whose column layout is compiler-specific; here they are written with 17
    significant digits, five per line -- any Tecplot/ParaView reader takes both.
  * `run(case_dir)`: the program itself -- for every file of fileRangeStart:fileRangeEnd:fileRangeStep read the
    filaments, evaluate the velocity at the cell centres ON THE GPU (vlc_gridgen; there is no CPU path), add the free
    stream, write the .tec file.  `python -m volcanor_b200.gridgen <case_dir>`.
"""
from __future__ import annotations

import struct
import sys
from pathlib import Path

import numpy as np

VR, VF = 50, 12
TEMPLATE_VERSION = "0.2"          # gridgen.f90:25


def _records(buf: bytes):
    """Split a gfortran unformatted sequential file into its records (4-byte markers before and after each)."""
    out, off = [], 0
    while off < len(buf):
        (n,) = struct.unpack_from("<i", buf, off)
        if n < 0:
            raise ValueError("filaments file: records larger than 2 GiB (gfortran sub-records) are not supported")
        body = buf[off + 4:off + 4 + n]
        (m,) = struct.unpack_from("<i", buf, off + 4 + n)
        if m != n or len(body) != n:
            raise ValueError("filaments file: record markers do not match (not a gfortran unformatted file?)")
        out.append(body)
        off += n + 8
    return out


def read_filaments(path) -> dict:
    """-> {"vrWing": (n,50), "vrNwake": (n,50), "vfNwakeTE": (n,12), "gamNwakeTE": (n,), "vfFwake": (n,12), "gamFwake": (n,)}"""
    rec = _records(Path(path).read_bytes())
    if len(rec) != 7 or any(len(r) != 4 for r in rec[:4]):
        raise ValueError(f"filaments file: expected 4 integer records + 3 data records, found {len(rec)} records")
    nW, nN, nT, nF = (struct.unpack("<i", r)[0] for r in rec[:4])

    def split(body, n1, w1, n2, w2, what):
        a = np.frombuffer(body, dtype="<f8")
        if a.size != n1 * w1 + n2 * w2:
            raise ValueError(f"filaments file: {what} record holds {a.size} doubles, expected {n1 * w1 + n2 * w2}")
        return a[:n1 * w1].reshape(n1, w1).copy(), a[n1 * w1:].reshape((n2, w2) if w2 > 1 else (n2,)).copy()

    vrW, vrN = split(rec[4], nW, VR, nN, VR, "vrWing/vrNwake")
    vfT, gT = split(rec[5], nT, VF, nT, 1, "vfNwakeTE/gamNwakeTE")
    vfF, gF = split(rec[6], nF, VF, nF, 1, "vfFwake/gamFwake")
    return {"vrWing": vrW, "vrNwake": vrN, "vfNwakeTE": vfT, "gamNwakeTE": gT, "vfFwake": vfF, "gamFwake": gF}


def write_filaments(path, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake) -> None:
    """The file filaments2file writes (libPostprocess.f90:455-464), byte for byte for the same arrays."""
    f64 = lambda a, w: np.ascontiguousarray(a, dtype="<f8").reshape(-1, w) if w > 1 else np.ascontiguousarray(a, dtype="<f8").reshape(-1)
    vrWing, vrNwake, vfNwakeTE, vfFwake = f64(vrWing, VR), f64(vrNwake, VR), f64(vfNwakeTE, VF), f64(vfFwake, VF)
    gamNwakeTE, gamFwake = f64(gamNwakeTE, 1), f64(gamFwake, 1)
    if gamNwakeTE.size != vfNwakeTE.shape[0] or gamFwake.size != vfFwake.shape[0]:
        raise ValueError("one circulation per filament")
    bodies = [struct.pack("<i", n) for n in (vrWing.shape[0], vrNwake.shape[0], vfNwakeTE.shape[0], vfFwake.shape[0])]
    bodies += [vrWing.tobytes() + vrNwake.tobytes(), vfNwakeTE.tobytes() + gamNwakeTE.tobytes(),
               vfFwake.tobytes() + gamFwake.tobytes()]
    with open(path, "wb") as fh:
        for b in bodies:
            fh.write(struct.pack("<i", len(b)) + b + struct.pack("<i", len(b)))


def read_gridconfig(path) -> dict:
    """gridconfig.nml: &VERSION fileFormatVersion / &INPUTS nx ny nz xyzMin xyzMax vel fileRangeStart/Step/End /"""
    vals: dict = {}
    for raw in Path(path).read_text().splitlines():
        line = raw.split("!")[0].strip()
        if not line or line.startswith("&") or line == "/" or "=" not in line:
            continue
        key, val = [s.strip() for s in line.split("=", 1)]
        items = [v.strip().strip("'\"") for v in val.rstrip(",").split(",") if v.strip()]
        vals[key.lower()] = items
    if str(vals.get("fileformatversion", [""])[0]) != TEMPLATE_VERSION:
        raise ValueError("ERROR: gridconfig.nml template version does not match")          # gridgen.f90:33-35
    num = lambda k, n, conv: [conv(x.lower().replace("d", "e")) for x in vals[k]][:n]
    cfg = {"nx": num("nx", 1, int)[0], "ny": num("ny", 1, int)[0], "nz": num("nz", 1, int)[0],
           "xyzMin": num("xyzmin", 3, float), "xyzMax": num("xyzmax", 3, float), "vel": num("vel", 3, float),
           "fileRangeStart": num("filerangestart", 1, int)[0], "fileRangeStep": num("filerangestep", 1, int)[0],
           "fileRangeEnd": num("filerangeend", 1, int)[0]}
    if any(a > b for a, b in zip(cfg["xyzMin"], cfg["xyzMax"])):
        raise ValueError("ERROR: All XYZmin values should be greater than XYZmax values")  # gridgen.f90:43-45 (sic)
    return cfg


def grid_nodes(nx, ny, nz, xyzMin, xyzMax):
    """grid(3, nx, ny, nz) of gridgen.f90:62-75 (libMath linspace: a + (b-a)/(n-1)*i) as [iz, iy, ix, 3]."""
    ax = []
    for n, lo, hi in zip((nx, ny, nz), xyzMin, xyzMax):
        ax.append(np.arange(n, dtype=np.float64) * ((hi - lo) / (n - 1)) + lo if n > 1 else np.array([float(lo)]))
    g = np.empty((nz, ny, nx, 3))
    g[..., 0] = ax[0][None, None, :]
    g[..., 1] = ax[1][None, :, None]
    g[..., 2] = ax[2][:, None, None]
    return g


def write_tecplot(path, nx, ny, nz, xyzMin, xyzMax, velCentre) -> None:
    """gridNNNNN.tec (gridgen.f90:150-164).  velCentre: [nz-1, ny-1, nx-1, 3] as vlc_gridgen returns it."""
    g = grid_nodes(nx, ny, nz, xyzMin, xyzMax)
    v = np.asarray(velCentre, dtype=np.float64).reshape(nz - 1, ny - 1, nx - 1, 3)
    with open(path, "x") as fh:                                     # status='new': never overwrite (gridgen.f90:147)
        fh.write(' TITLE = "Grid"\n VARIABLES = "X" "Y" "Z" "U" "V" "W"\n')
        fh.write(f' ZONE I={nx} J={ny} K={nz} T="Data"\n DATAPACKING=BLOCK\n')
        fh.write(" VARLOCATION=([4]=CELLCENTERED,[5]=CELLCENTERED,[6]=CELLCENTERED)\n")
        for block in (g[..., 0], g[..., 1], g[..., 2], v[..., 0], v[..., 1], v[..., 2]):
            flat = block.reshape(-1)                                # x fastest, then y, then z
            pad = (-flat.size) % 5
            rows = np.concatenate([flat, np.zeros(pad)]).reshape(-1, 5)
            lines = [" ".join(f"{x: .16E}" for x in r) for r in rows]
            if pad:
                lines[-1] = " ".join(f"{x: .16E}" for x in rows[-1][:5 - pad])
            fh.write("\n".join(lines) + "\n")


def read_tecplot(path):
    """Inverse of write_tecplot (tests): -> (nx, ny, nz, nodes [nz,ny,nx,3], velCentre [nz-1,ny-1,nx-1,3])."""
    lines = Path(path).read_text().splitlines()
    zone = lines[2].replace("=", " ").split()
    nx, ny, nz = int(zone[zone.index("I") + 1]), int(zone[zone.index("J") + 1]), int(zone[zone.index("K") + 1])
    data = np.array(" ".join(lines[5:]).split(), dtype=np.float64)
    nn, nc = nx * ny * nz, (nx - 1) * (ny - 1) * (nz - 1)
    nodes = np.stack([data[k * nn:(k + 1) * nn].reshape(nz, ny, nx) for k in range(3)], axis=-1)
    vel = np.stack([data[3 * nn + k * nc:3 * nn + (k + 1) * nc].reshape(nz - 1, ny - 1, nx - 1) for k in range(3)], axis=-1)
    return nx, ny, nz, nodes, vel


def sharded_velocities(ctx, cfg, f, world: int = 1, rank: int = 0, gather=None) -> np.ndarray:
    """velCentre (nz-1, ny-1, nx-1, 3) of one filaments record set with the cell list cut into contiguous slices, one
    per rank (one process per GPU; the cells are independent, gridgen.f90:116-139).  `gather(padded, shard)` all-gathers
    the (world*per, 3) host array in place (default: torch.distributed, any backend); world == 1 is the plain call."""
    from .sharding import TargetShard
    nx, ny, nz = cfg["nx"], cfg["ny"], cfg["nz"]
    m = (nx - 1) * (ny - 1) * (nz - 1)
    args = (nx, ny, nz, cfg["xyzMin"], cfg["xyzMax"], cfg["vel"], f["vrWing"], f["vrNwake"], f["vfNwakeTE"], f["gamNwakeTE"],
            f["vfFwake"], f["gamFwake"])
    if world == 1:
        return ctx.gridgen(*args)[1]
    sh = TargetShard(m, world, rank)
    padded = np.zeros((sh.padded, 3))
    if sh.count > 0:
        padded[sh.lo:sh.hi] = ctx.gridgen_slice(*args, sh.lo, sh.count)[1]
    (gather or _allgather_host)(padded, sh)
    return padded[:m].reshape(nz - 1, ny - 1, nx - 1, 3)


def _allgather_host(padded: np.ndarray, sh) -> None:
    """All-gather of the ranks' row slices of a host array through torch.distributed (NCCL needs device tensors)."""
    import torch
    import torch.distributed as dist
    from .sharding import allgather_slices
    t = torch.from_numpy(padded)
    if dist.get_backend() == "nccl":
        d = t.cuda()
        allgather_slices(d, sh)
        t.copy_(d.cpu())
    else:
        allgather_slices(t, sh)


def run(case_dir, ctx=None, verbose=True, world: int | None = None, rank: int | None = None, gather=None) -> list:
    """program gridgen in `case_dir` (reads gridconfig.nml and Results/filamentsNNNNN.dat, writes Results/gridNNNNN.tec).
    Under torchrun (WORLD_SIZE > 1, process group initialised by the caller) every rank evaluates its slice of the cells
    and rank 0 writes the files."""
    import os
    from .api import Context
    case_dir = Path(case_dir)
    cfg = read_gridconfig(case_dir / "gridconfig.nml")
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    own = ctx is None
    ctx = ctx or Context(int(os.environ.get("LOCAL_RANK", "0")))          # raises without a CUDA device: no CPU path
    written = []
    try:
        for k in range(cfg["fileRangeStart"], cfg["fileRangeEnd"] + 1, cfg["fileRangeStep"]):
            stamp = f"{k:05d}"
            f = read_filaments(case_dir / "Results" / f"filaments{stamp}.dat")
            vc = sharded_velocities(ctx, cfg, f, world, rank, gather)
            out = case_dir / "Results" / f"grid{stamp}.tec"
            if rank == 0:
                write_tecplot(out, cfg["nx"], cfg["ny"], cfg["nz"], cfg["xyzMin"], cfg["xyzMax"], vc)
            written.append(out)
            if verbose and rank == 0:
                nfil = 4 * (len(f["vrWing"]) + len(f["vrNwake"])) + len(f["vfNwakeTE"]) + len(f["vfFwake"])
                print(f"grid{stamp}.tec: {(cfg['nx'] - 1) * (cfg['ny'] - 1) * (cfg['nz'] - 1)} cell centres x {nfil} filaments"
                      + (f" on {world} GPUs" if world > 1 else ""))
    finally:
        if own:
            ctx.close()
    return written


if __name__ == "__main__":
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # torchrun: one process per GPU
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    run(sys.argv[1] if len(sys.argv) > 1 else ".")
