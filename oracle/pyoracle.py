"""ctypes binding of the CPU ORACLE (oracle/vlc_oracle.c) -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package volcanor_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
VF, VR, FW, WP, NPF = 12, 50, 13, 104, 240
EPS = 2.220446049250313e-16

_libs: dict[str, C.CDLL] = {}
_vp = C.c_void_p


def build(force: bool = False) -> None:
    """make -C oracle (gcc only; seconds)."""
    need = force or not (_HERE / "libvlc_oracle.so").exists() or not (_HERE / "libvlc_oracle_omp.so").exists()
    if not need:
        newest_src = max(p.stat().st_mtime for p in _HERE.glob("vlc_*.c*")) if list(_HERE.glob("vlc_*.c*")) else 0
        newest_src = max(newest_src, (_HERE / "vlc_oracle.h").stat().st_mtime)
        need = any((_HERE / n).stat().st_mtime < newest_src for n in ("libvlc_oracle.so", "libvlc_oracle_omp.so"))
    if need:
        r = subprocess.run(["make", "-C", str(_HERE)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def build_native(outdir: Path) -> Path | None:
    """-march=native rebuild of the baseline variant on the machine that will time it (bench.py)."""
    out = Path(outdir) / "libvlc_oracle_native.so"
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    cmd = [cc, "-O2", "-march=native", "-fPIC", "-fopenmp", "-std=c99", "-shared", "-o", str(out),
           str(_HERE / "vlc_oracle.c"), "-lm"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    except Exception:
        return None
    return out if r.returncode == 0 and out.exists() else None


def _bind(lib: C.CDLL) -> C.CDLL:
    i32, i64, d = C.c_int, C.c_long, C.c_double
    sig = {
        "orc_vind_flat": (None, [i64, _vp, _vp, _vp, i64, _vp, _vp]),
        "orc_vind_flat_ld": (None, [i64, _vp, _vp, _vp, i64, _vp, _vp, _vp]),
        "orc_num_threads": (i32, []),
        "orc_inv2": (i32, [i32, _vp, _vp]),
        "orc_matmulAX": (None, [i32, i32, _vp, _vp, _vp]),
        "orc_vf_vind": (None, [_vp, _vp, _vp]),
        "orc_vr_vind": (None, [_vp, _vp, _vp]),
        "orc_rotor_new": (_vp, [i32, i32, i32, i32, i32]),
        "orc_rotor_free": (None, [_vp]),
        "orc_rotor_wiP": (_vp, [_vp, i32]),
        "orc_rotor_waN": (_vp, [_vp, i32, i32]),
        "orc_rotor_waF": (_vp, [_vp, i32, i32]),
        "orc_rotor_wapF": (_vp, [_vp, i32, i32]),
        "orc_rotor_vel": (_vp, [_vp, i32, i32]),
        "orc_rotor_AIC": (_vp, [_vp, i32]),
        "orc_rotor_vec": (_vp, [_vp, i32]),
        "orc_rotor_set_rows": (None, [_vp, i32, i32]),
        "orc_rotor_set_params": (None, [_vp, i32, i32, i32, d, d, _vp, _vp, d, d, d, i32, i32]),
        "orc_rotor_vind_points": (None, [_vp, i32, i32, i64, _vp, _vp]),
        "orc_vind_onNwake_byRotor": (None, [_vp, _vp, i32, i32, i32, i32, _vp]),
        "orc_vind_onFwake_byRotor": (None, [_vp, _vp, i32, i32, _vp]),
        "orc_rotor_calcAIC": (i32, [_vp]),
        "orc_rotor_map_gam": (None, [_vp]),
        "orc_rotor_convectwake": (None, [_vp, i32, d, C.c_char]),
        "orc_pfwake_update": (i32, [_vp, _vp, _vp, _vp, i32, _vp, _vp, d]),
        "orc_rotor_updatePrescribedWake": (i32, [_vp, d, C.c_char]),
        "orc_rotor_get_pfHelix": (None, [_vp, i32, i32, _vp]),
        "orc_rotor_set_pfHelix": (None, [_vp, i32, i32, _vp]),
        "orc_rotor_get_presc": (None, [_vp, _vp]),
        "orc_rotor_assignshed": (None, [_vp, C.c_char_p]),
        "orc_rotor_age_wake": (None, [_vp, d]),
        "orc_rotor_dissipate_wake": (None, [_vp, d, d]),
        "orc_rotor_strain_wake": (None, [_vp]),
        "orc_rotor_burst_wake": (None, [_vp]),
        "orc_rotor_calc_skew": (None, [_vp]),
        "orc_rotor_rollup": (None, [_vp]),
        "orc_vel_order2_Nwake": (None, [_vp, _vp, i32, i32, _vp]),
        "orc_vel_order2_Fwake": (None, [_vp, _vp, i32, _vp]),
        "orc_rotor_dims": (None, [_vp, _vp]),
        "orc_rotor_get_params": (None, [_vp, _vp]),
        "orc_gridgen": (None, [i32, i32, i32, _vp, _vp, _vp, i64, _vp, i64, _vp, i64, _vp, _vp, i64, _vp, _vp, _vp, _vp]),
        # case driver (vlc_case.c)
        "orc_case_new": (_vp, [i32]),
        "orc_case_free": (None, [_vp]),
        "orc_case_set_config": (i32, [_vp, C.c_char_p, d]),
        "orc_case_set_geom": (i32, [_vp, i32, C.c_char_p, i32, _vp]),
        "orc_case_set_hooks": (None, [_vp, _vp]),
        "orc_case_init": (i32, [_vp]),
        "orc_case_init_rotors": (i32, [_vp]),
        "orc_case_step": (i32, [_vp]),
        "orc_case_iter": (i32, [_vp]),
        "orc_case_config": (_vp, [_vp]),
        "orc_case_rotor": (_vp, [_vp, i32]),
        "orc_case_error": (C.c_char_p, [_vp]),
        "orc_case_force_nondim": (None, [_vp, i32, _vp]),
        "orc_case_pairs_last_step": (d, [_vp]),
        "orc_rotor_dirLiftDrag": (None, [_vp]),
        "orc_rotor_calc_secAlpha": (None, [_vp]),
        "orc_rotor_calc_force": (None, [_vp, d, d]),
        "orc_rotor_sum_forces": (None, [_vp]),
        "orc_rotor_get_force_params": (None, [_vp, _vp]),
        "orc_rotor_get_file_params": (None, [_vp, _vp]),
        "orc_blade_sec": (_vp, [_vp, i32, C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    return lib


def load(variant: str = "strict", path: Path | None = None) -> C.CDLL:
    key = str(path) if path else variant
    if key in _libs:
        return _libs[key]
    if path is None:
        name = {"strict": "libvlc_oracle.so", "omp": "libvlc_oracle_omp.so"}[variant]
        path = _HERE / name
        if not path.exists():
            build()
    lib = _bind(C.CDLL(str(path)))
    _libs[key] = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def filament_records(p1, p2, rvc) -> np.ndarray:
    """(n, 12) array of vf_class records from flat filament arrays."""
    p1, p2, rvc = _f64(p1).reshape(-1, 3), _f64(p2).reshape(-1, 3), _f64(rvc).ravel()
    rec = np.zeros((rvc.size, VF), dtype=np.float64)
    rec[:, 0:3] = p1
    rec[:, 3:6] = p2
    rec[:, 8] = rvc  # rVc0
    rec[:, 9] = rvc  # rVc
    return rec


def vind_flat(p1, p2, rvc, gam, wake_flag, P, variant="strict", lib=None) -> np.ndarray:
    lib = lib or load(variant)
    rec = filament_records(p1, p2, rvc)
    gam = _f64(gam).ravel()
    P = _f64(P).reshape(-1, 3)
    V = np.empty_like(P)
    fl = None if wake_flag is None else np.ascontiguousarray(wake_flag, dtype=np.uint8)
    lib.orc_vind_flat(rec.shape[0], rec.ctypes.data, gam.ctypes.data, None if fl is None else fl.ctypes.data,
                      P.shape[0], P.ctypes.data, V.ctypes.data)
    return V


def vind_flat_ld(p1, p2, rvc, gam, wake_flag, P, variant="strict"):
    """(V, Vabs): long-double evaluation and the sum of |terms| used to scale tolerances."""
    lib = load(variant)
    rec = filament_records(p1, p2, rvc)
    gam = _f64(gam).ravel()
    P = _f64(P).reshape(-1, 3)
    V = np.empty_like(P)
    A = np.empty_like(P)
    fl = None if wake_flag is None else np.ascontiguousarray(wake_flag, dtype=np.uint8)
    lib.orc_vind_flat_ld(rec.shape[0], rec.ctypes.data, gam.ctypes.data, None if fl is None else fl.ctypes.data,
                         P.shape[0], P.ctypes.data, V.ctypes.data, A.ctypes.data)
    return V, A


def gridgen(nx, ny, nz, xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake,
            variant="strict"):
    """program gridgen (src/gridgen.f90): returns (gridCentre, velCentre), each (nz-1, ny-1, nx-1, 3).
    vrWing / vrNwake: (n, 50) vr_class records; vfNwakeTE / vfFwake: (n, 12) vf_class records + gam arrays."""
    lib = load(variant)
    a = [_f64(x) for x in (xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake)]
    shape = (nz - 1, ny - 1, nx - 1, 3)
    gc, vc = np.empty(shape), np.empty(shape)
    lib.orc_gridgen(nx, ny, nz, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].size // VR, a[3].ctypes.data,
                    a[4].size // VR, a[4].ctypes.data, a[5].size // VF, a[5].ctypes.data, a[6].ctypes.data,
                    a[7].size // VF, a[7].ctypes.data, a[8].ctypes.data, gc.ctypes.data, vc.ctypes.data)
    return gc, vc


class Rotor:
    """Owner of an orc_rotor_t with numpy views onto its reference-layout arrays."""

    def __init__(self, nb, nc, ns, nNwake, nFwake, variant="strict"):
        self.lib = load(variant)
        self.nb, self.nc, self.ns, self.nNwake, self.nFwake = nb, nc, ns, nNwake, nFwake
        self.h = self.lib.orc_rotor_new(nb, nc, ns, nNwake, nFwake)
        self.N = nc * ns * nb
        self.owner = True

    @classmethod
    def from_handle(cls, lib, h):
        """Non-owning wrapper of a rotor that lives inside an orc_case_t."""
        self = cls.__new__(cls)
        self.lib, self.h, self.owner = lib, h, False
        dims = (C.c_int * 10)()
        lib.orc_rotor_dims(h, dims)
        self.nb, self.nc, self.ns, self.nNwake, self.nFwake = dims[0:5]
        self.N = self.nc * self.ns * self.nb
        return self

    def dims(self) -> dict:
        d = (C.c_int * 10)()
        self.lib.orc_rotor_dims(self.h, d)
        return dict(zip(["nb", "nc", "ns", "nNwake", "nFwake", "rowNear", "rowFar", "nbConvect", "nNwakeEnd",
                         "nFwakeEnd"], list(d)))

    def presc(self):
        """(prescWakeNt, prescWakeAfterTruncNt, prescWakeGenNt) in time steps (classdef.f90:400, :3013-3017)."""
        d = (C.c_int * 3)()
        self.lib.orc_rotor_get_presc(self.h, d)
        return tuple(d)

    def params(self) -> dict:
        """What the wake mutators read from rotor_class (orc_rotor_get_params)."""
        o = np.zeros(18)
        self.lib.orc_rotor_get_params(self.h, o.ctypes.data)
        names = ["nbConvect", "axisymmetrySwitch", "ductSwitch", "suppressFwakeSwitch", "rollupStart", "rollupEnd"]
        d = {n: int(o[i]) for i, n in enumerate(names)}
        d.update(Omega=o[6], theta0=o[7], omegaSlow=o[8], apparentViscCoeff=o[9], decayCoeff=o[10], initWakeVel=o[11],
                 shaftAxis=o[12:15].copy(), hubCoords=o[15:18].copy())
        return d

    def sec(self, ib, name, width=1):
        """Sectional array of blade ib by name (vlc_case.c orc_blade_sec): (ns,) or (ns, 3)."""
        p = self.lib.orc_blade_sec(self.h, ib, name.encode())
        if not p:
            raise KeyError(name)
        return self._view(p, (self.ns, 3) if width == 3 else (self.ns,))

    def __del__(self):
        try:
            if self.h and self.owner:
                self.lib.orc_rotor_free(self.h)
                self.h = None
        except Exception:
            pass

    def _view(self, ptr, shape):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape)
        buf = (C.c_double * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape)

    # views are (cols, rows, record) = Fortran (record, rows, cols)
    def wiP(self, ib):
        return self._view(self.lib.orc_rotor_wiP(self.h, ib), (self.ns, self.nc, WP))

    def waN(self, ib, predicted=False):
        return self._view(self.lib.orc_rotor_waN(self.h, ib, int(predicted)), (self.ns, self.nNwake, VR))

    def waF(self, ib, predicted=False):
        return self._view(self.lib.orc_rotor_waF(self.h, ib, int(predicted)), (self.nFwake, FW))

    def wapF(self, ib, predicted=False):
        return self._view(self.lib.orc_rotor_wapF(self.h, ib, int(predicted)), (NPF, FW))

    def vel(self, ib, which):
        """0-3 velNwake, 1, Predicted, Step; 4-7 the same for velFwake; 8, 9 velNwake2 / 3; 10, 11 velFwake2 / 3."""
        if which < 4 or which in (8, 9):
            return self._view(self.lib.orc_rotor_vel(self.h, ib, which), (self.ns + 1, self.nNwake, 3))
        return self._view(self.lib.orc_rotor_vel(self.h, ib, which), (self.nFwake, 3))

    def AIC(self, inverse=False):
        return self._view(self.lib.orc_rotor_AIC(self.h, int(inverse)), (self.N, self.N)).T  # [row, col]

    def vec(self, which):
        return self._view(self.lib.orc_rotor_vec(self.h, which), (self.N,))

    def set_rows(self, rowNear, rowFar):
        self.rowNear, self.rowFar = rowNear, rowFar
        self.lib.orc_rotor_set_rows(self.h, rowNear, rowFar)

    def set_params(self, surfaceType=1, axisymmetrySwitch=0, nbConvect=None, Omega=0.0, omegaSlow=0.0,
                   shaftAxis=(0, 0, 1), hubCoords=(0, 0, 0), theta0=0.0, apparentViscCoeff=1.0, decayCoeff=0.0,
                   rollupStart=1, rollupEnd=None):
        sa, hc = _f64(shaftAxis), _f64(hubCoords)
        self.lib.orc_rotor_set_params(self.h, surfaceType, axisymmetrySwitch,
                                      self.nb if nbConvect is None else nbConvect, Omega, omegaSlow,
                                      sa.ctypes.data, hc.ctypes.data, theta0, apparentViscCoeff, decayCoeff,
                                      rollupStart, self.ns if rollupEnd is None else rollupEnd)

    def vind_points(self, what, P, predicted=False):
        P = _f64(P).reshape(-1, 3)
        V = np.empty_like(P)
        self.lib.orc_rotor_vind_points(self.h, what, int(predicted), P.shape[0], P.ctypes.data, V.ctypes.data)
        return V

    def vind_onNwake_byRotor(self, target: "Rotor", ib, rowNear, predicted=False):
        """vind_onNwake_byRotor(self, target.blade(ib)%waN(rowNear:nNwake, :)) -> (cols+1, rows, 3)."""
        rows = target.nNwake - rowNear + 1
        cols = target.ns
        base = target.lib.orc_rotor_waN(target.h, ib, int(predicted))
        out = np.empty((cols + 1, rows, 3))
        self.lib.orc_vind_onNwake_byRotor(self.h, base + 8 * VR * (rowNear - 1), rows, cols, target.nNwake,
                                          int(predicted), out.ctypes.data)
        return out

    def vind_onFwake_byRotor(self, target: "Rotor", ib, rowFar, predicted=False):
        rows = target.nFwake - rowFar + 1
        base = target.lib.orc_rotor_waF(target.h, ib, int(predicted))
        out = np.empty((max(rows, 0), 3))
        if rows > 0:
            self.lib.orc_vind_onFwake_byRotor(self.h, base + 8 * FW * (rowFar - 1), rows, int(predicted),
                                              out.ctypes.data)
        return out

    def calcAIC(self):
        return self.lib.orc_rotor_calcAIC(self.h)


class OrcConfig(C.Structure):
    """orc_config_t (vlc_case.h)"""
    _fields_ = [("nt", C.c_int), ("nr", C.c_int), ("dt", C.c_double), ("density", C.c_double),
                ("velSound", C.c_double), ("kinematicVisc", C.c_double), ("ntSub", C.c_int), ("ntSubInit", C.c_int),
                ("rotorForcePlot", C.c_int), ("wakeDissipation", C.c_int), ("wakeStrain", C.c_int),
                ("wakeBurst", C.c_int), ("wakeSuppress", C.c_int), ("slowStart", C.c_int), ("slowStartNt", C.c_int),
                ("fdScheme", C.c_int), ("initWakeVelNt", C.c_int)]


class OrcHooks(C.Structure):
    """orc_hooks_t (vlc_case.h): the five hot-path call sites of the driver."""
    VIND_POINTS = C.CFUNCTYPE(C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_long, _vp, _vp)
    VIND_ONN = C.CFUNCTYPE(C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp)
    VIND_ONF = C.CFUNCTYPE(C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp)
    CALC_AIC = C.CFUNCTYPE(C.c_int, _vp, C.c_int, _vp, _vp)
    SOLVE = C.CFUNCTYPE(C.c_int, _vp, C.c_int, _vp, _vp)
    WAKE_STAGE = C.CFUNCTYPE(C.c_int, _vp, C.c_int)   # optional device-resident stages (NULL = CPU restatement)
    CP_RHS_SOLVE = C.CFUNCTYPE(C.c_int, _vp)          # optional collocation-point stage (NULL = the call sites above)
    CP_FORCES = C.CFUNCTYPE(C.c_int, _vp, C.c_int)
    _fields_ = [("user", _vp), ("vind_points", VIND_POINTS), ("vind_onNwake", VIND_ONN), ("vind_onFwake", VIND_ONF),
                ("calcAIC", CALC_AIC), ("solve", SOLVE), ("wake_prestep", WAKE_STAGE), ("wake_convect", WAKE_STAGE),
                ("cp_rhs_solve", CP_RHS_SOLVE), ("cp_forces", CP_FORCES)]


class Case:
    """The oracle's restatement of `program main` (vlc_case.c).  `params` = {"config": {...}, "geom": [{...}, ...]}
    with the namelist variable names of config.nml / geomXX.nml (tests/golden/*.json hold the reference's cases)."""

    def __init__(self, params: dict, variant="strict"):
        self.lib = load(variant)
        geoms = params["geom"]
        self.nr = len(geoms)
        self.h = self.lib.orc_case_new(self.nr)
        self.ignored = []
        for k, v in params["config"].items():
            if self.lib.orc_case_set_config(self.h, k.encode(), float(v)) != 0:
                self.ignored.append(k)
        for ir, g in enumerate(geoms):
            for k, v in g.items():
                if isinstance(v, str):
                    self.ignored.append(k)
                    continue
                a = _f64(np.atleast_1d(v)).ravel()
                if self.lib.orc_case_set_geom(self.h, ir, k.encode(), a.size, a.ctypes.data) != 0:
                    self.ignored.append(k)
        self._hooks = None

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_case_free(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle case driver failed ({rc}): {self.lib.orc_case_error(self.h).decode()}")

    def set_hooks(self, hooks: "OrcHooks | None"):
        self._hooks = hooks  # keep the callbacks alive
        self.lib.orc_case_set_hooks(self.h, C.byref(hooks) if hooks is not None else None)

    def init_rotors(self):
        self._ck(self.lib.orc_case_init_rotors(self.h))

    def init(self):
        self._ck(self.lib.orc_case_init(self.h))

    def step(self):
        self._ck(self.lib.orc_case_step(self.h))

    @property
    def iter(self) -> int:
        return self.lib.orc_case_iter(self.h)

    @property
    def config(self) -> OrcConfig:
        return OrcConfig.from_address(self.lib.orc_case_config(self.h))

    def rotor(self, ir) -> Rotor:
        h = self.lib.orc_case_rotor(self.h, ir)
        if not h:
            raise RuntimeError("rotor does not exist yet: call init_rotors() / init() first")
        return Rotor.from_handle(self.lib, h)

    def force_nondim(self, ir=0) -> np.ndarray:
        out = np.empty(9)
        self.lib.orc_case_force_nondim(self.h, ir, out.ctypes.data)
        return out

    @property
    def pairs_last_step(self) -> float:
        return self.lib.orc_case_pairs_last_step(self.h)
