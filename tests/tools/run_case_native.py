#!/usr/bin/env python
"""Runs one of the reference's cases (tests/golden/<name>.json) through the reference driver restatement with its hot
path on the GPU (tests/native/case_gpu_hooks.c), prints CT history checkpoints and timings.  Test infrastructure
(uses oracle/): a profiling aid for profiles/, not a benchmark arm.

  python tests/tools/run_case_native.py caradonna [nsteps] [--resident] [--cp]
      --resident: tier 2b, the wake stays on the device; --cp: tier 2c as well (RHS, solve, map_gam, loads on the device)
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    resident = "--resident" in sys.argv
    if resident:
        sys.argv.remove("--resident")
    cp = "--cp" in sys.argv
    if cp:
        sys.argv.remove("--cp")
    name = sys.argv[1] if len(sys.argv) > 1 else "caradonna"
    import volcanor_b200 as vb
    from oracle import pyoracle
    from tests.test_gpu_case import _native_hooks
    from tests.test_gpu_resident import _resident_hooks
    pyoracle.build()
    fx = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    c = pyoracle.Case(fx)
    ctx = vb.Context(0)
    lib, h = (_resident_hooks if resident else _native_hooks)(c, ctx)
    if cp:
        import ctypes as C
        lib.case_hooks_enable_cp.argtypes = [C.c_void_p]
        assert lib.case_hooks_enable_cp(h) == 0, ctx.lib.vlc_last_error(ctx.h)
    t0 = time.perf_counter()
    c.init()
    nt = c.config.nt if len(sys.argv) < 3 else min(int(sys.argv[2]), c.config.nt)
    d = c.rotor(0).dims()
    print(f"{name}{' (wake resident on the device)' if resident else ''}{' (collocation-point stage on the device)' if cp else ''}: nt {c.config.nt} dt {c.config.dt:.6e} {d}; init {time.perf_counter() - t0:.2f} s")
    t1 = time.perf_counter()
    pairs, tlast = 0.0, t1
    per_step = []
    for it in range(1, nt + 1):
        ts = time.perf_counter()
        c.step()
        per_step.append(time.perf_counter() - ts)
        pairs += c.pairs_last_step
        if it % max(nt // 12, 1) == 0 or it == nt:
            now = time.perf_counter()
            f = c.force_nondim(0)
            info = ctx.rotor_info(0, resident)   # resident: the 'P' set is the one the last sweep used
            print(f"iter {it:5d}  CT {f[0]: .7e}  step pairs {c.pairs_last_step:.3e}  wall {now - t1:7.2f} s  "
                  f"last block {(now - tlast) / max(nt // 12, 1) * 1e3:7.2f} ms/step  shared={info['shared_active']} W={info['strip_width']}")
            tlast = now
    t2 = time.perf_counter()
    ps = np.array(per_step)
    med = np.array([np.median(ps[max(0, i - 10):i + 11]) for i in range(len(ps))])
    out = np.nonzero(ps > 3 * med)[0]
    print(f"steps slower than 3x their neighbourhood median: {len(out)}, {ps[out].sum():.2f} s in total; "
          + ", ".join(f"{i + 1}:{ps[i] * 1e3:.0f}ms" for i in out[:30]))
    print(f"kernel launches: {ctx.launch_count} = {ctx.launch_count / nt:.0f} per step")
    print(f"total {nt} steps in {t2 - t1:.2f} s = {nt / (t2 - t1):.2f} timesteps/s, {pairs:.3e} pair interactions "
          f"({pairs / (t2 - t1):.3e}/s incl. all host work), {lib.case_gpu_hooks_uploads(h)} uploads")


if __name__ == "__main__":
    main()
