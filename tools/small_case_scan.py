"""Wall time of a shipped case against a FIXED source split (vlc_set_tuning), next to the planner's own choice (nsplit 0).
Usage: python tools/small_case_scan.py tests/golden/katzNplotkin_AR04.json 0 4 8 13 20 26 40 52 80"""
import sys
import tempfile

from volcanor_b200 import run_case


def main():
    case, splits = sys.argv[1], [int(a) for a in sys.argv[2:]] or [0]
    with tempfile.TemporaryDirectory() as out:
        run_case.run(case, out=out, quiet=True)  # warm-up: module load, cuSOLVER handle, allocations
        for ns in splits:
            runs = [run_case.run(case, out=out, quiet=True, nsplit=ns, stats=True) for _ in range(3)]
            r = min(runs, key=lambda x: x["loop_s"])
            sw = r["sweeps"]["bs_lattice_kernel"]
            print(f"nsplit {ns:3d}: {r['loop_s'] * 1e3:8.1f} ms for {r['steps']} steps; lattice-led sweeps {sw['ms']:8.2f} ms on the device "
                  f"({100 * sw['pipe_frac']:.1f} % FP64 pipe)", flush=True)


if __name__ == "__main__":
    main()
