#!/usr/bin/env python
"""Turn the ncu artefacts that tools/gpu_capture.sh leaves in gpurun_out/ (scratch) into the small tracked
summaries under profiles/ (what DESIGN.md and bench.py's roofline numbers cite).

  python tools/summarize_ncu.py r01a            # reads gpurun_out/*_r01a.*, writes profiles/r01a_*.{md,csv}
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.min.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = name.replace("void ", "")
    for pre in ("vlc::", "at::native::", "at::"):
        name = name.replace(pre, "")
    return name.split("(")[0][:90]


def launches(tag: str) -> str | None:
    f = OUT / f"launches_{tag}.csv"
    if not f.exists():
        return None
    lines = [l for l in f.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    agg: "OrderedDict[str, list[float]]" = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        agg.setdefault(short(r["Kernel Name"]), []).append(v)
    total = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list `{tag}` -- per-kernel device time (cold-cache, serialised: compare SHARES)", "",
           f"command: `ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3 "
           f"--no-e2e --no-cpu-baseline` ({len(rows)} launches, {total:.1f} ms total)", "",
           "| kernel | launches | total ms | mean ms | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {100 * sum(v) / total:.3f} % |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    # compact csv: one line per launch
    with open(PROF / f"{tag}_launches.csv", "w") as fh:
        fh.write("id,kernel,grid,block,ms\n")
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
            fh.write(f'{r["ID"]},"{short(r["Kernel Name"])}","{r["Grid Size"]}","{r["Block Size"]}",{v:.6f}\n')
    return "\n".join(out)


def full(tag: str) -> str | None:
    rep = OUT / f"prof_{tag}.ncu-rep"
    if not rep.exists():
        return None
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full capture `{tag}` of the dominant kernel", "",
           "command: `ncu --set full --clock-control none --import-source on -k regex:bs_lattice -s 4 -c 1 python bench.py "
           "--steps 1 --warmup 3 --no-e2e --no-cpu-baseline` (1 000 192 filaments x 258 176 targets)", ""]
    for d in data:
        out.append(f"## `{short(d[idx['Kernel Name']])}`  grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}")
        out += ["", "| metric | value | unit |", "|---|---:|---|"]
        for k in RAW_KEYS:
            if k in idx:
                out.append(f"| {k} | {d[idx[k]]} | {units[idx[k]]} |")
        out.append("")
    # source page: top stall lines + instruction mix by pipe for the first launch
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    if src.strip():
        (OUT / f"prof_{tag}_source.csv").write_text(src)
        try:
            raw = list(csv.reader(io.StringIO(src)))
            h = next(i for i, r in enumerate(raw) if "Source" in r or "SASS" in r)    # line 1 is the "Kernel Name" row
            srows = [dict(zip(raw[h], r)) for r in raw[h + 1:] if len(r) == len(raw[h])]
            stall = {c: sum(float(r[c] or 0) for r in srows) for c in raw[h] if c.startswith("stall_") and "Not Issued" not in c}
            if sum(stall.values()) > 0:
                tot_s = sum(stall.values())
                out += ["## Warp-state samples by stall reason (source page, all samples)", "", "| reason | samples | share |", "|---|---:|---:|"]
                out += [f"| {k} | {int(v)} | {100 * v / tot_s:.1f} % |" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]] + [""]
            col_inst = next((c for c in srows[0] if c.startswith("# Instructions Executed") or c == "Instructions Executed"), None)
            col_src = next((c for c in srows[0] if c in ("Source", "SASS", "Instruction")), None)
            if col_inst and col_src:
                mix: dict[str, float] = {}
                for r in srows:
                    try:
                        n = float(r[col_inst])
                    except (ValueError, TypeError):
                        continue
                    op = r[col_src].strip().split()[0] if r[col_src].strip() else "?"
                    if op.startswith("@"):
                        op = r[col_src].strip().split()[1]
                    mix[op.split(".")[0]] = mix.get(op.split(".")[0], 0.0) + n
                tot = sum(mix.values())
                out += ["## SASS instruction mix (warp-level instructions executed, first captured launch)", "",
                        "| opcode | executed | share |", "|---|---:|---:|"]
                for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:14]:
                    out.append(f"| {op} | {n:.3e} | {100 * n / tot:.2f} % |")
                out.append("")
        except Exception as e:  # the source page layout differs between ncu versions; the raw page is the record
            out.append(f"(source page not summarised: {e})")
    (PROF / f"{tag}_bs_sweep_full.md").write_text("\n".join(out) + "\n")
    return "\n".join(out)


def bench(tag: str):
    for name in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
        f = OUT / name
        if f.exists() and f.read_text().strip():
            try:
                j = json.loads(f.read_text().strip().splitlines()[-1])
                (PROF / name).write_text(json.dumps(j, indent=1) + "\n")
            except json.JSONDecodeError:
                pass
    for name in (f"pytest_gpu_{tag}.log", f"smoke_{tag}.log", f"smi_{tag}.txt"):
        f = OUT / name
        if f.exists():
            (PROF / name).write_text(f.read_text())


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01a"
    PROF.mkdir(exist_ok=True)
    a = launches(tag)
    b = full(tag)
    bench(tag)
    print(a or "no launch list")
    print(b or "no full capture")
