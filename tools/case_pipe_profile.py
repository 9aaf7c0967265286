"""FP64-pipe use of a case's sweeps over the run: one line per block of steps (vlc_sweep_stats reset per block).
Usage: PYTHONPATH=. python tools/case_pipe_profile.py tests/golden/caradonna.json [block]"""
import json
import sys
from pathlib import Path

from volcanor_b200 import api
from volcanor_b200.run_case import CaseDriver


def main():
    case = json.loads(Path(sys.argv[1]).read_text())
    block = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ctx = api.Context(0)
    peak, _ = ctx.measure_fp64_peak()
    drv = CaseDriver(case, ctx)
    drv.init()
    nt = drv.info()["nt"]
    it = 0
    while it < nt:
        ctx.sweep_stats(1)
        n = min(block, nt - it)
        for _ in range(n):
            drv.step()
        it += n
        st = ctx.sweep_stats(-1)["bs_lattice_kernel"]
        if st["sweep_ms"] > 0:
            print(f"steps {it - n + 1:4d}-{it:4d}: {st['launches']:4d} lattice-led sweeps, {st['sweep_ms']:8.2f} ms on the device "
                  f"({st['sweep_ms'] / st['launches'] * 1e3:7.1f} us each), {st['sweep_pairs']:.3e} pairs, FP64 pipe "
                  f"{100 * st['sweep_fp64_instr'] * 2 / (st['sweep_ms'] * 1e-3) / peak:5.1f} %", flush=True)
    drv.close()
    ctx.close()


if __name__ == "__main__":
    main()
