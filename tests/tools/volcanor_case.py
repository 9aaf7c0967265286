#!/usr/bin/env python
"""Runs an unmodified reference case directory (config.nml, geomNN.nml[, *.xyz]) through the restatement of the
reference's driver and writes Results/r01ForceNonDim.csv in the reference's own format.  Test infrastructure /
demonstration of the caller side of the path; `--gpu` forwards the hot path to libvolcanor_b200.so
through tests/native/case_gpu_hooks.c, without it the CPU oracle computes everything.

  python tests/tools/volcanor_case.py /path/to/some.case [--gpu] [--nt N] [--out DIR]
"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case_dir")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--resident", action="store_true", help="with --gpu: keep the wake on the device (C ABI tier 2b)")
    ap.add_argument("--nt", type=int, default=0, help="stop after this many steps (default: the case's nt)")
    ap.add_argument("--out", default=None, help="results directory (default: <case_dir>/Results)")
    args = ap.parse_args()
    from oracle import casefile, pyoracle
    pyoracle.build()
    fx = casefile.read_case(args.case_dir)
    c = pyoracle.Case(fx)
    ctx = None
    if args.gpu:
        import volcanor_b200 as vb
        from tests.test_gpu_case import _native_hooks
        from tests.test_gpu_resident import _resident_hooks
        ctx = vb.Context(0)
        (_resident_hooks if args.resident else _native_hooks)(c, ctx)
    out = Path(args.out) if args.out else Path(args.case_dir) / "Results"
    out.mkdir(parents=True, exist_ok=True)
    t0 = time.perf_counter()
    c.init()
    nt = c.config.nt if args.nt <= 0 else min(args.nt, c.config.nt)
    lines = [[casefile.HEADER] for _ in range(c.nr)]
    for ir in range(c.nr):
        lines[ir].append(casefile.force_nondim_line(0, c.force_nondim(ir)))
    for it in range(1, nt + 1):
        c.step()
        for ir in range(c.nr):
            lines[ir].append(casefile.force_nondim_line(it, c.force_nondim(ir)))
    for ir in range(c.nr):
        (out / f"r{ir + 1:02d}ForceNonDim.csv").write_text("\n".join(lines[ir]) + "\n")
    print(f"{fx['name']}: {nt} steps in {time.perf_counter() - t0:.2f} s ({'GPU' if args.gpu else 'CPU oracle'}), "
          f"wrote {out}/rNNForceNonDim.csv")


if __name__ == "__main__":
    main()
