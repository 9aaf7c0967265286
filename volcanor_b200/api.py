"""ctypes binding of include/volcanor_b200.h plus a small object wrapper.

Names and argument meaning follow the reference's procedures (src/libCommon.f90:114-211,
src/classdef.f90:4151-4479) so that the parity tests read like the reference's own tests.
Every failure of the library raises :class:`VlcError` (the reference `error stop`s).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ROOT = _HERE.parent
_LIB = None

VR_DOUBLES = 50
FWAKE_DOUBLES = 13
WINGPANEL_DOUBLES = 104
NPFWAKE = 240


class VlcError(RuntimeError):
    pass


def lib_path() -> Path:
    # VOLCANOR_B200_LIB selects an experimental build of the same sources (tuning runs); default = the in-tree library
    alt = os.environ.get("VOLCANOR_B200_LIB")
    return Path(alt) if alt else _HERE / "libvolcanor_b200.so"


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -> volcanor_b200/libvolcanor_b200.so (sm_100a only, in-tree so it travels with gpurun)."""
    out = lib_path()
    srcs = [_HERE / "csrc" / "capi.cu"]
    deps = list((_HERE / "csrc").glob("*.cu*")) + [_ROOT / "include" / "volcanor_b200.h"]
    if out.exists() and not force and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++",
           "-shared", "-o", str(out)] + [str(s) for s in srcs] + ["-lcusolver"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise VlcError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr, file=sys.stderr)
    return out


def case_lib_path() -> Path:
    return _HERE / "libvolcanor_case.so"


def build_case_driver(force: bool = False) -> Path:
    """g++ -> volcanor_b200/libvolcanor_case.so: the product-side driver of an unmodified .case directory
    (csrc/case_driver.cpp; host code only -- every stage of the hot path is a call into libvolcanor_b200.so, which it links).
    -ffp-contract=off: the reference's statement order without FMA contraction, like its own -O2 build on x86-64."""
    out = case_lib_path()
    src = _HERE / "csrc" / "case_driver.cpp"
    deps = [src, _ROOT / "include" / "volcanor_b200.h"]
    if out.exists() and not force and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    build_library()
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-shared", "-o", str(out), str(src),
           f"-L{_HERE}", "-lvolcanor_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise VlcError("g++ failed:\n" + r.stdout + r.stderr)
    return out


def _declared_symbols() -> list[str]:
    hdr = (_ROOT / "include" / "volcanor_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vlc_[a-zA-Z0-9_]+)\s*\(", hdr)))


DECLARED_SYMBOLS = _declared_symbols()

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


def _preload_torch_nccl():
    """The library binds NCCL at run time by soname (csrc/group.hpp).  In a Python process that will also import torch,
    the copy that ships with torch (nvidia/nccl/lib, newer than the system's) must be the one in the process -- a system
    libnccl.so.2 loaded first would be picked up by torch's own DT_NEEDED and break its import.  Load torch's copy first
    when there is one; a plain C / Fortran driver never gets here and uses the system library."""
    if "torch" in sys.modules:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for d in spec.submodule_search_locations:
                so = Path(d) / "lib" / "libnccl.so.2"
                if so.exists():
                    C.CDLL(str(so), mode=C.RTLD_GLOBAL)
                    return
    except Exception:
        pass


def load_library() -> C.CDLL:
    """Load the CUDA library.  Fails loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise VlcError(f"{p} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(volcanor_b200 has no CPU fallback)")
    lib = C.CDLL(str(p))
    i64, i32 = C.c_int64, C.c_int
    sig = {
        "vlc_create": (i32, [i32, C.POINTER(_vp)]),
        "vlc_destroy": (i32, [_vp]),
        "vlc_create_multi": (i32, [i32, C.POINTER(i32), C.POINTER(_vp)]),
        "vlc_comm_unique_id": (i32, [_vp]),
        "vlc_comm_init_rank": (i32, [_vp, i32, i32, _vp]),
        "vlc_comm_info": (i32, [_vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "vlc_last_error": (C.c_char_p, [_vp]),
        "vlc_version": (C.c_char_p, []),
        "vlc_set_stream": (i32, [_vp, _vp, i32]),
        "vlc_sync": (i32, [_vp]),
        "vlc_device_info": (i32, [_vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "vlc_set_tuning": (i32, [_vp, i32, i32]),
        "vlc_set_precision": (i32, [_vp, i32]),
        "vlc_launch_count": (i64, [_vp]),
        "vlc_set_sources": (i32, [_vp, i32, i64, _vp, _vp, _vp, _vp, _vp]),
        "vlc_set_sources_dev": (i32, [_vp, i32, i64, _vp, _vp, _vp, _vp, _vp]),
        "vlc_num_sources": (i64, [_vp, i32]),
        "vlc_vind": (i32, [_vp, i32, i64, _vp, _vp]),
        "vlc_vind_dev": (i32, [_vp, i32, i64, _vp, _vp]),
        "vlc_vind_range_dev": (i32, [_vp, i32, i64, i64, i64, _vp, _vp]),
        "vlc_source_tile": (i32, []),
        "vlc_rotor_define": (i32, [_vp, i32, i32, i32, i32, i32, i32, i32]),
        "vlc_rotors_clear": (i32, [_vp]),
        "vlc_rotor_set_rows": (i32, [_vp, i32, i32, i32]),
        "vlc_rotor_put_wing": (i32, [_vp, i32, i32, _vp]),
        "vlc_rotor_put_nwake": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_put_fwake": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_put_pfwake": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_put_wing_gam": (i32, [_vp, i32, i32, _vp]),
        "vlc_rotor_vind_bywing": (i32, [_vp, i32, i64, _vp, _vp]),
        "vlc_rotor_vind_bywake": (i32, [_vp, i32, i32, i64, _vp, _vp]),
        "vlc_rotor_vind_bywing_boundVortices": (i32, [_vp, i32, i64, _vp, _vp]),
        "vlc_rotor_vind_bywing_chordwiseVortices": (i32, [_vp, i32, i64, _vp, _vp]),
        "vlc_rotor_vind": (i32, [_vp, i32, i32, i64, _vp, _vp]),
        "vlc_vind_onNwake_byRotor": (i32, [_vp, i32, _vp, i32, i32, i32, i32, _vp]),
        "vlc_vind_onFwake_byRotor": (i32, [_vp, i32, _vp, i32, i32, _vp]),
        "vlc_rotor_calcAIC": (i32, [_vp, i32, _vp]),
        "vlc_rotor_solve": (i32, [_vp, i32, _vp, _vp]),
        "vlc_rotor_get_AIC_inv": (i32, [_vp, i32, _vp]),
        "vlc_rotor_set_wake_params": (i32, [_vp, i32, i32, i32, i32, i32, i32, i32, C.c_double, C.c_double, C.c_double,
                                            C.c_double]),
        "vlc_rotor_set_frame": (i32, [_vp, i32, _vp, _vp]),
        "vlc_rotor_assignshed": (i32, [_vp, i32, i32]),
        "vlc_rotor_age_wake": (i32, [_vp, i32, C.c_double, C.c_double]),
        "vlc_rotor_dissipate_wake": (i32, [_vp, i32, C.c_double, C.c_double]),
        "vlc_rotor_strain_wake": (i32, [_vp, i32]),
        "vlc_rotor_wake_to_predicted": (i32, [_vp, i32]),
        "vlc_rotor_convectwake": (i32, [_vp, i32, C.c_double, i32]),
        "vlc_rotor_updatePrescribedWake": (i32, [_vp, i32, C.c_double, i32, i32]),
        "vlc_rotor_burst_wake": (i32, [_vp, i32, C.c_double, C.c_double]),
        "vlc_rotor_calc_skew": (i32, [_vp, i32]),
        "vlc_rotor_get_pfwake": (i32, [_vp, i32, i32, i32, _vp, _vp]),
        "vlc_rotor_put_pfwake_helix": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_rollup": (i32, [_vp, i32]),
        "vlc_wake_sweep": (i32, [_vp, i32, i32]),
        "vlc_wake_sweep_count": (i32, [_vp, C.POINTER(i64)]),
        "vlc_wake_sweep_slice": (i32, [_vp, i32, i64, i64, _vp]),
        "vlc_wake_sweep_scatter": (i32, [_vp, i32, i32, _vp]),
        "vlc_rotor_wakevel_op": (i32, [_vp, i32, i32]),
        "vlc_rotor_wakevel_copy": (i32, [_vp, i32, i32, i32]),
        "vlc_rotor_wakevel_lincomb": (i32, [_vp, i32, i32, i32, C.POINTER(i32), _dp, C.c_double]),
        "vlc_rotor_get_nwake": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_get_fwake": (i32, [_vp, i32, i32, i32, _vp]),
        "vlc_rotor_put_wakevel": (i32, [_vp, i32, i32, i32, _vp, _vp]),
        "vlc_rotor_get_wakevel": (i32, [_vp, i32, i32, i32, _vp, _vp]),
        "vlc_rotor_calc_RHS": (i32, [_vp, i32, _vp, _vp]),
        "vlc_rotor_reset_velCP": (i32, [_vp, i32]),
        "vlc_rotor_solve_map_gam": (i32, [_vp, i32, _vp]),
        "vlc_rotor_put_sections": (i32, [_vp, i32, i32, _vp]),
        "vlc_rotor_calc_velCPTotal": (i32, [_vp, i32]),
        "vlc_rotor_calc_force": (i32, [_vp, i32, C.c_double, C.c_double, C.c_double, i32]),
        "vlc_rotor_get_loads": (i32, [_vp, i32, i32, _vp]),
        "vlc_rotor_get_wing": (i32, [_vp, i32, i32, _vp]),
        "vlc_convect_dev": (i32, [_vp, i64, _vp, _vp, C.c_double]),
        "vlc_ab2_dev": (i32, [_vp, i64, _vp, _vp, _vp]),
        "vlc_am2_dev": (i32, [_vp, i64, _vp, _vp, _vp]),
        "vlc_vel_order2_dev": (i32, [_vp, i32, i32, _vp, _vp, _vp]),
        "vlc_dissipate_dev": (i32, [_vp, i64, _vp, i64, _vp, C.c_double, C.c_double, C.c_double, C.c_double]),
        "vlc_dissipate_lattice_dev": (i32, [_vp, i32, i32, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double]),
        "vlc_strain_dev": (i32, [_vp, i64, _vp, _vp, _vp, _vp, _vp]),
        "vlc_pack_lattice_dev": (i32, [_vp, i32, i32, i32, i32, _vp, _vp, _vp, i32, _vp, _vp, _vp]),
        "vlc_pack_lattice": (i32, [_vp, i32, i32, i32, i32, _vp, _vp, _vp, i32, _vp, _vp, _vp]),
        "vlc_set_shared_nodes": (i32, [_vp, i32]),
        "vlc_set_lattice_tuning": (i32, [_vp, i32, i32]),
        "vlc_set_info": (i32, [_vp, i32, C.POINTER(i64)]),
        "vlc_rotor_info": (i32, [_vp, i32, i32, C.POINTER(i64)]),
        "vlc_last_sweep_ms": (i32, [_vp, _dp, _dp]),
        "vlc_lattice_targets_dev": (i32, [_vp, i32, i32, _vp, _vp]),
        "vlc_lattice_scatter_dev": (i32, [_vp, i32, i32, _vp, _vp]),
        "vlc_gridgen": (i32, [_vp, i32, i32, i32, _vp, _vp, _vp, i64, _vp, i64, _vp, i64, _vp, _vp, i64, _vp, _vp, _vp, _vp]),
        "vlc_gridgen_slice": (i32, [_vp, i32, i32, i32, _vp, _vp, _vp, i64, _vp, i64, _vp, i64, _vp, _vp, i64, _vp, _vp, i64, i64,
                                    _vp, _vp]),
        "vlc_event_record": (i32, [_vp, i32]),
        "vlc_event_elapsed_ms": (i32, [_vp, i32, i32, _dp]),
        "vlc_l2_flush": (i32, [_vp]),
        "vlc_sweep_stats": (i32, [_vp, i32, C.POINTER(i64), _dp, _dp, _dp]),
        "vlc_measure_fp64_peak": (i32, [_vp, i32, _dp, _dp]),
        "vlc_measure_fp64_rate": (i32, [_vp, i32, i32, _dp, _dp]),
        "vlc_probe_rsqrt": (i32, [_vp, i64, _vp, _vp, _vp, _vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._vlc_signatures = sig
    _LIB = lib
    return lib


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.size == int(np.prod(shape)), (a.shape, shape)
    return a


def _ptr(a) -> int | None:
    """Raw address of a numpy array, torch tensor (host or device) or int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One library context = one GPU.  Mirrors the reference call sites (see the header)."""

    def __init__(self, device: int = 0, devices=None):
        """device: one GPU.  devices = [d0, d1, ...]: ONE handle over several GPUs of the box (vlc_create_multi): state is
        replicated, sweeps are target-sharded, the wake sweep all-gathers inside the library."""
        self.lib = load_library()
        h = _vp()
        if devices is not None:
            _preload_torch_nccl()
            devs = [int(d) for d in devices]
            arr = (C.c_int * len(devs))(*devs)
            rc = self.lib.vlc_create_multi(len(devs), arr, C.byref(h))
            device = devs[0]
        else:
            rc = self.lib.vlc_create(device, C.byref(h))
        if rc != 0:
            raise VlcError(f"vlc_create failed ({rc}): {self.lib.vlc_last_error(None).decode()}")
        self.h = h
        self.device = device

    # -- multi-GPU data plane behind the ABI ----------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id made by rank 0; the launcher hands it to every rank (comm_init_rank)."""
        _preload_torch_nccl()
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.vlc_comm_unique_id(buf)
        if rc != 0:
            raise VlcError(f"vlc_comm_unique_id failed ({rc}): {lib.vlc_last_error(None).decode()}")
        return buf.raw

    def comm_init_rank(self, world: int, rank: int, unique_id: bytes | None):
        """One process per GPU: join the library-owned communicator; wake_sweep and the host-pointer sweeps become
        collective (same arguments on every rank)."""
        _preload_torch_nccl()
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._ck(self.lib.vlc_comm_init_rank(self.h, world, rank, buf))

    def comm_info(self) -> dict:
        w, r, t = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.vlc_comm_info(self.h, C.byref(w), C.byref(r), C.byref(t)))
        return {"world": w.value, "rank": r.value, "transport": {0: "single", 1: "nccl", 2: "peer"}[t.value]}

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise VlcError(f"[{rc}] {self.lib.vlc_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.vlc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr: int | None, use_own: bool = False):
        """stream_ptr = torch.cuda.current_stream().cuda_stream (0/None = legacy default stream)."""
        self._ck(self.lib.vlc_set_stream(self.h, stream_ptr or None, int(use_own)))

    def sync(self):
        self._ck(self.lib.vlc_sync(self.h))

    def set_tuning(self, targets_per_thread: int = 0, nsplit: int = 0):
        self._ck(self.lib.vlc_set_tuning(self.h, targets_per_thread, nsplit))

    def set_precision(self, mode: int):
        """0 = full (third-order rsqrt refinement), 1 = fast (second order, pair error <= ~4e-14)."""
        self._ck(self.lib.vlc_set_precision(self.h, mode))

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._ck(self.lib.vlc_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}

    @property
    def launch_count(self) -> int:
        return int(self.lib.vlc_launch_count(self.h))

    def event_record(self, slot: int):
        """CUDA event on the library's stream (every member's for a multi-GPU handle)."""
        self._ck(self.lib.vlc_event_record(self.h, slot))

    def event_elapsed_ms(self, slot_a: int, slot_b: int) -> float:
        """Device time between two recorded events; the largest over the members of a multi-GPU handle."""
        ms = C.c_double()
        self._ck(self.lib.vlc_event_elapsed_ms(self.h, slot_a, slot_b, C.byref(ms)))
        return ms.value

    def sweep_stats(self, reset: int = 0) -> dict:
        """Per-launch device times of the dominant kernels since the last reset (reset=1 starts collecting)."""
        n, ms, pr, ins = (C.c_int64 * 2)(), (C.c_double * 4)(), (C.c_double * 4)(), (C.c_double * 4)()
        self._ck(self.lib.vlc_sweep_stats(self.h, reset, n, ms, pr, ins))
        return {k: {"launches": int(n[i]), "ms": ms[i], "pairs": pr[i], "fp64_instr": ins[i],
                    "sweep_ms": ms[2 + i], "sweep_pairs": pr[2 + i], "sweep_fp64_instr": ins[2 + i]}
                for i, k in enumerate(("bs_lattice_kernel", "bs_sweep_kernel"))}

    def l2_flush(self):
        self._ck(self.lib.vlc_l2_flush(self.h))

    def measure_fp64_peak(self, iters: int = 20000) -> tuple[float, float]:
        f, ms = C.c_double(), C.c_double()
        self._ck(self.lib.vlc_measure_fp64_peak(self.h, iters, C.byref(f), C.byref(ms)))
        return f.value, ms.value

    def measure_fp64_rate(self, pattern: int, iters: int = 20000) -> tuple[float, float]:
        """pattern 0 = vlc_measure_fp64_peak's chains; 1 = three distinct changing register operands per DFMA."""
        f, ms = C.c_double(), C.c_double()
        self._ck(self.lib.vlc_measure_fp64_rate(self.h, iters, pattern, C.byref(f), C.byref(ms)))
        return f.value, ms.value

    def probe_rsqrt(self, x):
        x = _f64(x).ravel()
        seed, full, fast = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self._ck(self.lib.vlc_probe_rsqrt(self.h, x.size, _ptr(x), _ptr(seed), _ptr(full), _ptr(fast)))
        return seed, full, fast

    # -- tier 1 -----------------------------------------------------------------------------
    def set_sources(self, set_: int, p1, p2, rvc, gam, wake_flag=None):
        """Host arrays: p1, p2 (n,3) [= Fortran (3,n)], rvc (n), gam (n), wake_flag (n) uint8 or None."""
        p1, p2, rvc, gam = _f64(p1), _f64(p2), _f64(rvc), _f64(gam)
        n = rvc.size
        assert p1.size == 3 * n and p2.size == 3 * n and gam.size == n
        fl = None if wake_flag is None else np.ascontiguousarray(wake_flag, dtype=np.uint8)
        self._ck(self.lib.vlc_set_sources(self.h, set_, n, _ptr(p1), _ptr(p2), _ptr(rvc), _ptr(gam), _ptr(fl)))

    def set_sources_dev(self, set_: int, n: int, p1, p2, rvc, gam, wake_flag=None):
        self._ck(self.lib.vlc_set_sources_dev(self.h, set_, n, _ptr(p1), _ptr(p2), _ptr(rvc), _ptr(gam),
                                              _ptr(wake_flag)))

    def num_sources(self, set_: int) -> int:
        return int(self.lib.vlc_num_sources(self.h, set_))

    def vind(self, set_: int, P) -> np.ndarray:
        """V (m,3) at host points P (m,3)."""
        P = _f64(P)
        m = P.size // 3
        V = np.empty((m, 3), dtype=np.float64)
        self._ck(self.lib.vlc_vind(self.h, set_, m, _ptr(P), _ptr(V)))
        return V

    def vind_into(self, set_: int, m: int, P, V):
        """Host buffers (numpy or pinned torch tensors) without allocation."""
        self._ck(self.lib.vlc_vind(self.h, set_, m, _ptr(P), _ptr(V)))

    def vind_dev(self, set_: int, m: int, dP, dV):
        self._ck(self.lib.vlc_vind_dev(self.h, set_, m, _ptr(dP), _ptr(dV)))

    def vind_range_dev(self, set_: int, first: int, count: int, m: int, dP, dV):
        self._ck(self.lib.vlc_vind_range_dev(self.h, set_, first, count, m, _ptr(dP), _ptr(dV)))

    # -- tier 2 -----------------------------------------------------------------------------
    def rotor_define(self, ir, nb, nc, ns, nNwake, nFwake, surfaceType=1):
        self._ck(self.lib.vlc_rotor_define(self.h, ir, nb, nc, ns, nNwake, nFwake, surfaceType))

    def rotors_clear(self):
        """Forget every declared rotor (the resident sweeps sum over all of them)."""
        self._ck(self.lib.vlc_rotors_clear(self.h))

    def rotor_set_rows(self, ir, rowNear, rowFar):
        self._ck(self.lib.vlc_rotor_set_rows(self.h, ir, rowNear, rowFar))

    def rotor_put_wing(self, ir, ib, wiP):
        a = _f64(wiP)  # bound to a local: the converted array must outlive the foreign call
        self._ck(self.lib.vlc_rotor_put_wing(self.h, ir, ib, _ptr(a)))

    def rotor_put_wing_gam(self, ir, ib, gam):
        a = _f64(gam)
        self._ck(self.lib.vlc_rotor_put_wing_gam(self.h, ir, ib, _ptr(a)))

    def rotor_put_nwake(self, ir, ib, waN, predicted=False):
        a = _f64(waN)
        self._ck(self.lib.vlc_rotor_put_nwake(self.h, ir, ib, int(predicted), _ptr(a)))

    def rotor_put_fwake(self, ir, ib, waF, predicted=False):
        a = _f64(waF)
        self._ck(self.lib.vlc_rotor_put_fwake(self.h, ir, ib, int(predicted), _ptr(a)))

    def rotor_put_pfwake(self, ir, ib, wapF, predicted=False):
        a = _f64(wapF)
        self._ck(self.lib.vlc_rotor_put_pfwake(self.h, ir, ib, int(predicted), _ptr(a)))

    def _points(self, fn, P, *pre):
        P = _f64(P)
        m = P.size // 3
        V = np.empty((m, 3), dtype=np.float64)
        self._ck(fn(self.h, *pre, m, _ptr(P), _ptr(V)))
        return V

    def rotor_vind_bywing(self, ir, P):
        return self._points(self.lib.vlc_rotor_vind_bywing, P, ir)

    def rotor_vind_bywake(self, ir, P, predicted=False):
        return self._points(self.lib.vlc_rotor_vind_bywake, P, ir, int(predicted))

    def rotor_vind_bywing_boundVortices(self, ir, P):
        return self._points(self.lib.vlc_rotor_vind_bywing_boundVortices, P, ir)

    def rotor_vind_bywing_chordwiseVortices(self, ir, P):
        return self._points(self.lib.vlc_rotor_vind_bywing_chordwiseVortices, P, ir)

    def rotor_vind(self, ir, P, predicted=False):
        return self._points(self.lib.vlc_rotor_vind, P, ir, int(predicted))

    def vind_onNwake_byRotor(self, ir, Nwake, rows, cols, ld, predicted=False, offset_records=0, out=None):
        """Nwake: the parent record array (ld*cols*50 doubles); slice starts `offset_records` records in.
        Returns (cols+1, rows, 3) = Fortran (3, rows, cols+1); `out`: a C-contiguous float64 array of that shape to fill."""
        Nwake = _f64(Nwake)
        if out is None:
            out = np.empty((cols + 1, rows, 3), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == 3 * rows * (cols + 1)
        self._ck(self.lib.vlc_vind_onNwake_byRotor(self.h, ir, Nwake.ctypes.data + 8 * VR_DOUBLES * offset_records,
                                                   rows, cols, ld, int(predicted), _ptr(out)))
        return out

    def vind_onFwake_byRotor(self, ir, Fwake, rows, predicted=False, offset_records=0, out=None):
        Fwake = _f64(Fwake)
        if out is None:
            out = np.empty((rows, 3), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == 3 * rows
        self._ck(self.lib.vlc_vind_onFwake_byRotor(self.h, ir, Fwake.ctypes.data + 8 * FWAKE_DOUBLES * offset_records,
                                                   rows, int(predicted), _ptr(out)))
        return out

    def vind_onNwake_byRotor_ptr(self, ir, Nwake_ptr: int, rows, cols, ld, predicted=False):
        """Same, from the raw address of element (1,1) of the slice (what the Fortran shim passes)."""
        out = np.empty((cols + 1, rows, 3), dtype=np.float64)
        self._ck(self.lib.vlc_vind_onNwake_byRotor(self.h, ir, Nwake_ptr, rows, cols, ld, int(predicted), _ptr(out)))
        return out

    def vind_onFwake_byRotor_ptr(self, ir, Fwake_ptr: int, rows, predicted=False):
        out = np.empty((rows, 3), dtype=np.float64)
        self._ck(self.lib.vlc_vind_onFwake_byRotor(self.h, ir, Fwake_ptr, rows, int(predicted), _ptr(out)))
        return out

    def rotor_calcAIC(self, ir, N, want_matrix=True):
        A = np.empty((N, N), dtype=np.float64, order="F") if want_matrix else None
        self._ck(self.lib.vlc_rotor_calcAIC(self.h, ir, _ptr(A)))
        return A

    def rotor_solve(self, ir, RHS):
        RHS = _f64(RHS)
        g = np.empty_like(RHS)
        self._ck(self.lib.vlc_rotor_solve(self.h, ir, _ptr(RHS), _ptr(g)))
        return g

    def rotor_get_AIC_inv(self, ir, N):
        A = np.empty((N, N), dtype=np.float64, order="F")
        self._ck(self.lib.vlc_rotor_get_AIC_inv(self.h, ir, _ptr(A)))
        return A

    # ---- tier 2b: the reference's wake mutators on the device copies (device-resident stepping) ----
    VEL_FIRST_STEP, VEL_AB2, VEL_AM2, VEL_SHIFT_HISTORY, VEL_ORDER2, VEL_COPY_TO_STEP = range(6)

    def rotor_set_wake_params(self, ir, nbConvect, axisymmetrySwitch, ductSwitch, suppressFwakeSwitch, rollupStart,
                              rollupEnd, rollupSign, apparentViscCoeff, decayCoeff, initWakeVel=0.0):
        self._ck(self.lib.vlc_rotor_set_wake_params(self.h, ir, nbConvect, axisymmetrySwitch, ductSwitch,
                                                    suppressFwakeSwitch, rollupStart, rollupEnd, rollupSign,
                                                    apparentViscCoeff, decayCoeff, initWakeVel))

    def rotor_set_frame(self, ir, shaftAxis, hubCoords):
        sa, hc = _f64(shaftAxis, (3,)), _f64(hubCoords, (3,))  # both alive across the call (lists are converted copies)
        self._ck(self.lib.vlc_rotor_set_frame(self.h, ir, _ptr(sa), _ptr(hc)))

    def rotor_assignshed(self, ir, edge: str):
        self._ck(self.lib.vlc_rotor_assignshed(self.h, ir, {"LE": 0, "TE": 1}[edge]))

    def rotor_age_wake(self, ir, dt, omegaSlow):
        self._ck(self.lib.vlc_rotor_age_wake(self.h, ir, dt, omegaSlow))

    def rotor_dissipate_wake(self, ir, dt, kinematicVisc):
        self._ck(self.lib.vlc_rotor_dissipate_wake(self.h, ir, dt, kinematicVisc))

    def rotor_strain_wake(self, ir):
        self._ck(self.lib.vlc_rotor_strain_wake(self.h, ir))

    def rotor_wake_to_predicted(self, ir):
        self._ck(self.lib.vlc_rotor_wake_to_predicted(self.h, ir))

    def rotor_convectwake(self, ir, dt, wakeType: str = "C"):
        self._ck(self.lib.vlc_rotor_convectwake(self.h, ir, dt, {"C": 0, "P": 1}[wakeType]))

    def rotor_calc_skew(self, ir):
        """rotor%calc_skew() (classdef.f90:4919-4936): vr%skew of the active near-wake rings (record member 49)."""
        self._ck(self.lib.vlc_rotor_calc_skew(self.h, ir))

    def rotor_burst_wake(self, ir, skewLimit, largeCoreRadius):
        """rotor%burst_wake() (classdef.f90:4911-4917) on the device's far wake."""
        self._ck(self.lib.vlc_rotor_burst_wake(self.h, ir, skewLimit, largeCoreRadius))

    def rotor_updatePrescribedWake(self, ir, deltaPsi, prescWakeGenNt=0, wakeType: str = "C"):
        """rotor%updatePrescribedWake(dt, wakeType) (classdef.f90:5170-5218) on the device records; deltaPsi = omegaSlow*dt."""
        self._ck(self.lib.vlc_rotor_updatePrescribedWake(self.h, ir, deltaPsi, prescWakeGenNt, {"C": 0, "P": 1}[wakeType]))

    def rotor_put_pfwake_helix(self, ir, ib, helix, predicted=False):
        hx = _f64(helix, (2,))
        self._ck(self.lib.vlc_rotor_put_pfwake_helix(self.h, ir, ib, int(predicted), _ptr(hx)))

    def rotor_get_pfwake(self, ir, ib, predicted=False):
        """-> (wapF (240, 13), (helixPitch, helixRadius)) of one blade."""
        w, hx = np.empty((240, 13)), np.empty(2)
        self._ck(self.lib.vlc_rotor_get_pfwake(self.h, ir, ib, int(predicted), _ptr(w), _ptr(hx)))
        return w, hx

    def rotor_rollup(self, ir):
        self._ck(self.lib.vlc_rotor_rollup(self.h, ir))

    def wake_sweep(self, predicted=False, add_init_wake_vel=False):
        self._ck(self.lib.vlc_wake_sweep(self.h, int(predicted), int(add_init_wake_vel)))

    def wake_sweep_count(self) -> int:
        m = C.c_int64()
        self._ck(self.lib.vlc_wake_sweep_count(self.h, C.byref(m)))
        return m.value

    def wake_sweep_slice(self, predicted, first, count, d_vel):
        self._ck(self.lib.vlc_wake_sweep_slice(self.h, int(predicted), first, count, _ptr(d_vel)))

    def wake_sweep_scatter(self, predicted, add_init_wake_vel, d_vel):
        self._ck(self.lib.vlc_wake_sweep_scatter(self.h, int(predicted), int(add_init_wake_vel), _ptr(d_vel)))

    def rotor_wakevel_op(self, ir, op: int):
        self._ck(self.lib.vlc_rotor_wakevel_op(self.h, ir, op))

    VEL_ARRAY, VEL_ARRAY_1, VEL_ARRAY_PREDICTED, VEL_ARRAY_STEP, VEL_ARRAY_2, VEL_ARRAY_3 = range(6)

    def rotor_wakevel_copy(self, ir, dst: int, src: int):
        self._ck(self.lib.vlc_rotor_wakevel_copy(self.h, ir, dst, src))

    def rotor_wakevel_lincomb(self, ir, dst: int, src, coef, divisor: float):
        """dst = (coef[0]*src[0] + ...)/divisor on the velocity arrays (ids VEL_ARRAY_*), terms added left to right."""
        n = len(src)
        self._ck(self.lib.vlc_rotor_wakevel_lincomb(self.h, ir, dst, n, (C.c_int * n)(*[int(x) for x in src]),
                                                    (C.c_double * n)(*[float(x) for x in coef]), float(divisor)))

    def rotor_get_nwake(self, ir, ib, nNwake, ns, predicted=False):
        a = np.empty((ns, nNwake, 50), dtype=np.float64)
        self._ck(self.lib.vlc_rotor_get_nwake(self.h, ir, ib, int(predicted), _ptr(a)))
        return a

    def rotor_get_fwake(self, ir, ib, nFwake, predicted=False):
        a = np.empty((nFwake, 13), dtype=np.float64)
        self._ck(self.lib.vlc_rotor_get_fwake(self.h, ir, ib, int(predicted), _ptr(a)))
        return a

    def rotor_put_wakevel(self, ir, ib, which, velN=None, velF=None):
        vn = _f64(velN) if velN is not None else None   # keep the (possibly copied) arrays alive across the call
        vf = _f64(velF) if velF is not None else None
        self._ck(self.lib.vlc_rotor_put_wakevel(self.h, ir, ib, which, _ptr(vn), _ptr(vf)))

    def rotor_get_wakevel(self, ir, ib, which, nNwake, ns, nFwake):
        vn = np.empty((ns + 1, nNwake, 3), dtype=np.float64)
        vf = np.empty((nFwake, 3), dtype=np.float64)
        self._ck(self.lib.vlc_rotor_get_wakevel(self.h, ir, ib, which, _ptr(vn) if vn.size else None,
                                                _ptr(vf) if vf.size else None))
        return vn, vf

    # ---- tier 2c: the collocation-point stage on the device copies of the wing records ----
    def rotor_calc_RHS(self, ir, m, N, want_velCP=True, want_RHS=True):
        """main.f90:548-603 for rotor ir; m = nbConvect*nc*ns, N = nc*ns*nb.  Returns (velCP (m, 3), RHS (N,))."""
        v = np.empty((m, 3), dtype=np.float64) if want_velCP else None
        r = np.empty(N, dtype=np.float64) if want_RHS else None
        self._ck(self.lib.vlc_rotor_calc_RHS(self.h, ir, _ptr(v), _ptr(r)))
        return v, r

    def rotor_reset_velCP(self, ir):
        """velCP <- velCPm (its kinematic part) on the device records: head of a sub-iteration pass (main.f90:528-547)."""
        self._ck(self.lib.vlc_rotor_reset_velCP(self.h, ir))

    def rotor_solve_map_gam(self, ir, N, want_gamVec=True):
        g = np.empty(N, dtype=np.float64) if want_gamVec else None
        self._ck(self.lib.vlc_rotor_solve_map_gam(self.h, ir, _ptr(g)))
        return g

    @staticmethod
    def pack_sections(secTauCapChord, secNormalVec, secCP, secArea, yAxisAziFlap, zAxisAziFlap) -> np.ndarray:
        """The 10*ns + 6 block of vlc_rotor_put_sections from the blade_class arrays ((ns, 3) = Fortran (3, ns))."""
        return np.concatenate([_f64(a).ravel() for a in (secTauCapChord, secNormalVec, secCP, secArea, yAxisAziFlap,
                                                         zAxisAziFlap)])

    def rotor_put_sections(self, ir, ib, sec):
        a = _f64(sec)
        self._ck(self.lib.vlc_rotor_put_sections(self.h, ir, ib, _ptr(a)))

    def rotor_calc_velCPTotal(self, ir):
        self._ck(self.lib.vlc_rotor_calc_velCPTotal(self.h, ir))

    def rotor_calc_force(self, ir, density, dt, Omega, spanwiseLiftSwitch=0):
        self._ck(self.lib.vlc_rotor_calc_force(self.h, ir, float(density), float(dt), float(Omega), int(spanwiseLiftSwitch)))

    LOADS_SEC3 = ("secChordwiseResVel", "secDragDir", "secLiftDir", "secForceInertial", "secLift", "secDrag", "secLiftUnsteady")
    LOADS_SEC1 = ("secAlpha", "secCL", "secCD", "secCLu")

    def rotor_get_loads(self, ir, ib, ns) -> dict:
        """Loads block of blade ib (12 + 25*ns doubles) as a dict of named arrays."""
        a = np.empty(12 + 25 * ns, dtype=np.float64)
        self._ck(self.lib.vlc_rotor_get_loads(self.h, ir, ib, _ptr(a)))
        out = {"forceInertial": a[0:3], "lift": a[3:6], "drag": a[6:9], "liftUnsteady": a[9:12]}
        for k, name in enumerate(self.LOADS_SEC3):
            out[name] = a[12 + 3 * ns * k: 12 + 3 * ns * (k + 1)].reshape(ns, 3)
        for k, name in enumerate(self.LOADS_SEC1):
            out[name] = a[12 + 21 * ns + ns * k: 12 + 21 * ns + ns * (k + 1)]
        return out

    def rotor_get_wing(self, ir, ib, nc, ns):
        a = np.empty((ns, nc, 104), dtype=np.float64)
        self._ck(self.lib.vlc_rotor_get_wing(self.h, ir, ib, _ptr(a)))
        return a

    def gridgen(self, nx, ny, nz, xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake):
        """program gridgen (src/gridgen.f90): (gridCentre, velCentre), each (nz-1, ny-1, nx-1, 3)."""
        a = [_f64(x) for x in (xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake)]
        shape = (nz - 1, ny - 1, nx - 1, 3)
        gc, vc = np.empty(shape), np.empty(shape)
        self._ck(self.lib.vlc_gridgen(self.h, nx, ny, nz, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), a[3].size // VR_DOUBLES,
                                      _ptr(a[3]), a[4].size // VR_DOUBLES, _ptr(a[4]), a[5].size // 12, _ptr(a[5]),
                                      _ptr(a[6]), a[7].size // 12, _ptr(a[7]), _ptr(a[8]), _ptr(gc), _ptr(vc)))
        return gc, vc

    def gridgen_slice(self, nx, ny, nz, xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake,
                      first, count):
        """Cells [first, first+count) of program gridgen's cell list: (gridCentre, velCentre), each (count, 3)."""
        a = [_f64(x) for x in (xyzMin, xyzMax, vel, vrWing, vrNwake, vfNwakeTE, gamNwakeTE, vfFwake, gamFwake)]
        gc, vc = np.empty((count, 3)), np.empty((count, 3))
        self._ck(self.lib.vlc_gridgen_slice(self.h, nx, ny, nz, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), a[3].size // VR_DOUBLES,
                                            _ptr(a[3]), a[4].size // VR_DOUBLES, _ptr(a[4]), a[5].size // 12, _ptr(a[5]),
                                            _ptr(a[6]), a[7].size // 12, _ptr(a[7]), _ptr(a[8]), int(first), int(count),
                                            _ptr(gc) if count else None, _ptr(vc) if count else None))
        return gc, vc

    # -- tier 3 (device pointers: torch tensors or ints) --------------------------------------
    def convect_dev(self, n, x, v, dt):
        self._ck(self.lib.vlc_convect_dev(self.h, n, _ptr(x), _ptr(v), dt))

    def ab2_dev(self, n, v, v1, out):
        self._ck(self.lib.vlc_ab2_dev(self.h, n, _ptr(v), _ptr(v1), _ptr(out)))

    def am2_dev(self, n, vp, vs, out):
        self._ck(self.lib.vlc_am2_dev(self.h, n, _ptr(vp), _ptr(vs), _ptr(out)))

    def vel_order2_dev(self, rows, cols, vn, vnp1, out):
        self._ck(self.lib.vlc_vel_order2_dev(self.h, rows, cols, _ptr(vn), _ptr(vnp1), _ptr(out)))

    def dissipate_dev(self, n_rvc, rvc, n_gam, gam, apparentViscCoeff, kinematicVisc, decayCoeff, dt):
        self._ck(self.lib.vlc_dissipate_dev(self.h, n_rvc, _ptr(rvc), n_gam, _ptr(gam), apparentViscCoeff,
                                            kinematicVisc, decayCoeff, dt))

    def dissipate_lattice_dev(self, nrows, ns, rvc4, gam, apparentViscCoeff, kinematicVisc, decayCoeff, dt):
        self._ck(self.lib.vlc_dissipate_lattice_dev(self.h, nrows, ns, _ptr(rvc4), _ptr(gam), apparentViscCoeff,
                                                    kinematicVisc, decayCoeff, dt))

    def strain_dev(self, n, p1, p2, l0, rvc0, rvc):
        self._ck(self.lib.vlc_strain_dev(self.h, n, _ptr(p1), _ptr(p2), _ptr(l0), _ptr(rvc0), _ptr(rvc)))

    def pack_lattice_dev(self, set_, append, nrows, ns, nodes, gam, rvc4, nfar=0, far_nodes=None, gamF=None,
                         rvcF=None):
        self._ck(self.lib.vlc_pack_lattice_dev(self.h, set_, int(append), nrows, ns, _ptr(nodes), _ptr(gam),
                                               _ptr(rvc4), nfar, _ptr(far_nodes), _ptr(gamF), _ptr(rvcF)))

    def pack_lattice(self, set_, append, nrows, ns, nodes, gam, rvc4, nfar=0, far_nodes=None, gamF=None, rvcF=None):
        """Host arrays (numpy / pinned torch): nodes (ns+1, nrows+1, 3), gam (ns, nrows), rvc4 (ns, nrows, 4), far chain."""
        self._ck(self.lib.vlc_pack_lattice(self.h, set_, int(append), nrows, ns, _ptr(nodes), _ptr(gam), _ptr(rvc4),
                                           nfar, _ptr(far_nodes), _ptr(gamF), _ptr(rvcF)))

    def set_shared_nodes(self, on: bool):
        """Lattice sets: shared-node kernel (default) or the flat reference enumeration."""
        self._ck(self.lib.vlc_set_shared_nodes(self.h, int(on)))

    def set_lattice_tuning(self, strip_width: int = 0, targets_per_thread: int = 0):
        self._ck(self.lib.vlc_set_lattice_tuning(self.h, strip_width, targets_per_thread))

    def set_info(self, set_: int) -> dict:
        out = (C.c_int64 * 6)()
        self._ck(self.lib.vlc_set_info(self.h, set_, out))
        return {"filaments": out[0], "lattice_records": out[1], "remainder_filaments": out[2], "shared_active": out[3],
                "strip_width": out[4], "tail_strip_width": out[5]}

    def rotor_info(self, ir: int, predicted: bool = False) -> dict:
        out = (C.c_int64 * 6)()
        self._ck(self.lib.vlc_rotor_info(self.h, ir, int(predicted), out))
        return {"filaments": out[0], "lattice_records": out[1], "remainder_filaments": out[2], "shared_active": out[3],
                "strip_width": out[4], "tail_strip_width": out[5]}

    def last_sweep_ms(self) -> tuple[float, float]:
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.vlc_last_sweep_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def lattice_targets_dev(self, nrows, ns, nodes, P):
        self._ck(self.lib.vlc_lattice_targets_dev(self.h, nrows, ns, _ptr(nodes), _ptr(P)))

    def lattice_scatter_dev(self, nrows, ns, nodes, P):
        self._ck(self.lib.vlc_lattice_scatter_dev(self.h, nrows, ns, _ptr(nodes), _ptr(P)))
