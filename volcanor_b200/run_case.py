"""Run an UNMODIFIED reference case directory on the GPU and write the reference's force history files.

  python -m volcanor_b200.run_case /path/to/some.case [--nt N] [--out DIR] [--gpus N]

`config.nml`, `geomNN.nml` and PLOT3D geometry files are read as they are (volcanor_b200/casefile.py); the program flow of
`src/main.f90` -- rotor%init, rigid motion, the order of the stages, the sum over blades -- is the product-side driver
volcanor_b200/csrc/case_driver.cpp; every stage of the hot path (AIC + LU, right-hand side, solve, loads, wake mutators,
wake sweeps) runs in libvolcanor_b200.so through its C ABI.  `--gpus N` puts N GPUs behind the one handle
(vlc_create_multi).  Output: `<out>/rNNForceNonDim.csv` in the format of force2file (libPostprocess.f90:824-838).
There is no CPU path: without a CUDA device the run fails at vlc_create.
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

from . import api, casefile


class CaseDriver:
    """ctypes view of libvolcanor_case.so (csrc/case_driver.cpp)."""

    def __init__(self, case: dict, ctx: "api.Context"):
        api.build_case_driver()
        self.lib = C.CDLL(str(api.case_lib_path()))
        L = self.lib
        L.vcase_new.restype = C.c_void_p
        L.vcase_new.argtypes = [C.c_int]
        L.vcase_error.restype = C.c_char_p
        for name, args in {"vcase_free": [C.c_void_p], "vcase_error": [C.c_void_p], "vcase_set_config": [C.c_void_p, C.c_char_p, C.c_double],
                           "vcase_set_geom": [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p],
                           "vcase_attach": [C.c_void_p, C.c_void_p], "vcase_init": [C.c_void_p], "vcase_step": [C.c_void_p],
                           "vcase_iter": [C.c_void_p], "vcase_info": [C.c_void_p, C.c_int, C.c_void_p],
                           "vcase_force_nondim": [C.c_void_p, C.c_int, C.c_void_p], "vcase_get_gamvec": [C.c_void_p, C.c_int, C.c_void_p],
                           "vcase_get_loads": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
                           "vcase_stage_seconds": [C.c_void_p, C.c_void_p]}.items():
            getattr(L, name).argtypes = args
        self.nr = int(case["config"].get("nr", len(case["geom"])))
        self.h = L.vcase_new(self.nr)
        if not self.h:
            raise api.VlcError("vcase_new failed")
        self.ctx = ctx
        self.ignored = []
        for k, v in case["config"].items():
            if isinstance(v, (int, float)) and L.vcase_set_config(self.h, k.encode(), float(v)) == 1:
                self.ignored.append(k)
        for ir, g in enumerate(case["geom"]):
            for k, v in g.items():
                if isinstance(v, str):
                    continue
                a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.float64)))
                if L.vcase_set_geom(self.h, ir, k.encode(), a.size, a.ctypes.data) == 1:
                    self.ignored.append(f"geom{ir + 1:02d}.{k}")
        self._ck(L.vcase_attach(self.h, ctx.h if ctx is not None else None))

    def _ck(self, rc):
        if rc != 0:
            raise api.VlcError(f"[{rc}] {self.lib.vcase_error(self.h).decode()}")

    def init(self):
        self._ck(self.lib.vcase_init(self.h))

    def step(self):
        self._ck(self.lib.vcase_step(self.h))

    @property
    def iter(self) -> int:
        return int(self.lib.vcase_iter(self.h))

    def info(self, ir: int = 0) -> dict:
        a = np.zeros(12)
        self.lib.vcase_info(self.h, ir, a.ctypes.data)
        keys = ("nt", "dt", "nb", "nc", "ns", "nNwake", "nFwake", "rowNear", "rowFar", "nonDimforceDenominator", "nbConvect", "wing_uploads")
        return {k: (float(v) if k in ("dt", "nonDimforceDenominator") else int(v)) for k, v in zip(keys, a)}

    def force_nondim(self, ir: int = 0) -> np.ndarray:
        a = np.zeros(9)
        self.lib.vcase_force_nondim(self.h, ir, a.ctypes.data)
        return a

    def gamvec(self, ir: int = 0) -> np.ndarray:
        i = self.info(ir)
        a = np.zeros(i["nb"] * i["nc"] * i["ns"])
        self.lib.vcase_get_gamvec(self.h, ir, a.ctypes.data)
        return a

    def loads(self, ir: int, ib: int) -> np.ndarray:
        a = np.zeros(12 + 25 * self.info(ir)["ns"])
        self.lib.vcase_get_loads(self.h, ir, ib, a.ctypes.data)
        return a

    def stage_seconds(self) -> dict:
        a = np.zeros(5)
        self.lib.vcase_stage_seconds(self.h, a.ctypes.data)
        return dict(zip(("motion", "upload+prestep", "rhs+solve", "forces", "wake"), a.tolist()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.vcase_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(case_dir, nt: int = 0, out=None, gpus: int = 1, quiet: bool = False, nsplit: int = 0, stats: bool = False) -> dict:
    if isinstance(case_dir, dict):          # an already parsed case ({"name", "config", "geom"}: what read_case returns)
        case, case_dir = case_dir, Path(".")
    elif str(case_dir).endswith(".json"):   # the same, stored as JSON (tests/golden/*.json: the shipped cases, parsed once)
        import json
        case, case_dir = json.loads(Path(case_dir).read_text()), Path(case_dir).parent
    else:
        case = casefile.read_case(case_dir)
    ctx = api.Context(devices=list(range(gpus))) if gpus > 1 else api.Context(0)
    if nsplit:
        ctx.set_tuning(0, nsplit)   # a fixed source split: a target's summation order no longer depends on the launch it is in
    drv = CaseDriver(case, ctx)
    if stats:
        ctx.sweep_stats(1)
    t0 = time.perf_counter()
    drv.init()
    n = drv.info()["nt"] if nt <= 0 else min(nt, drv.info()["nt"])
    lines = [[casefile.HEADER, casefile.force_nondim_line(0, drv.force_nondim(ir))] for ir in range(drv.nr)]
    t1 = time.perf_counter()
    for it in range(1, n + 1):
        drv.step()
        for ir in range(drv.nr):
            lines[ir].append(casefile.force_nondim_line(it, drv.force_nondim(ir)))
    ctx.sync()
    t2 = time.perf_counter()
    outdir = Path(out) if out else Path(case_dir) / "Results"
    outdir.mkdir(parents=True, exist_ok=True)
    for ir in range(drv.nr):
        (outdir / f"r{ir + 1:02d}ForceNonDim.csv").write_text("\n".join(lines[ir]) + "\n")
    res = {"case": case["name"], "steps": n, "init_s": t1 - t0, "loop_s": t2 - t1, "timesteps_per_s": n / (t2 - t1) if t2 > t1 else 0.0,
           "launches": ctx.launch_count, "gpus": gpus, "out": str(outdir), "ignored_keys": drv.ignored,
           "stage_seconds": drv.stage_seconds(),
           "final": [drv.force_nondim(ir).tolist() for ir in range(drv.nr)]}
    if stats:   # device time and FP64-pipe use of the sweeps (vlc_sweep_stats: one event pair per launch)
        st, (peak, _) = ctx.sweep_stats(-1), ctx.measure_fp64_peak()
        res["sweeps"] = {k: {"launches": v["launches"], "ms": v["sweep_ms"], "pairs": v["sweep_pairs"],
                             "pipe_frac": (v["sweep_fp64_instr"] * 2 / (v["sweep_ms"] * 1e-3) / peak) if v["sweep_ms"] > 0 else 0.0}
                         for k, v in st.items()}
    if not quiet:
        print(f"{case['name']}: {n} steps in {t2 - t1:.2f} s ({res['timesteps_per_s']:.1f} timesteps/s, {ctx.launch_count / max(n, 1):.0f} "
              f"kernel launches per step) on {gpus} GPU(s); wrote {outdir}/rNNForceNonDim.csv")
        print("  host time per stage [s]: " + ", ".join(f"{k} {v:.3f}" for k, v in res["stage_seconds"].items()))
        for k, v in res.get("sweeps", {}).items():
            print(f"  sweeps led by {k}: {v['launches']} launches, {v['ms'] * 1e-3:.3f} s on the device, {v['pairs']:.3e} pair "
                  f"interactions, FP64 pipe {100 * v['pipe_frac']:.1f} %")
    drv.close()
    ctx.close()
    return res


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("case_dir", help="a .case directory (config.nml, geomNN.nml[, PLOT3D grid]) or a JSON dump of read_case()")
    ap.add_argument("--nt", type=int, default=0, help="stop after this many steps (default: the case's nt)")
    ap.add_argument("--out", default=None, help="results directory (default: <case_dir>/Results)")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs behind the one library handle (vlc_create_multi)")
    ap.add_argument("--nsplit", type=int, default=0,
                    help="fixed number of source splits per sweep (default: chosen per launch): makes the history independent of "
                         "--gpus bit for bit")
    ap.add_argument("--stats", action="store_true", help="time every sweep on the device (vlc_sweep_stats) and print the totals")
    args = ap.parse_args(argv)
    run(args.case_dir, args.nt, args.out, args.gpus, nsplit=args.nsplit, stats=args.stats)
    return 0


if __name__ == "__main__":
    sys.exit(main())
