#!/usr/bin/env python
"""Static issue-cost model of the FP64 instructions in a kernel's main loop, from `cuobjdump -sass`.

Measured on B200 (tools/micro/fp64_operands.cu, profiles/r01h_fp64_operands.md): an FP64 instruction occupies its
scheduler's FP64 issue for 2 cycles, except a DFMA whose three source operands are three DIFFERENT registers, which takes
3 (immediates, constants and a register repeated in two slots are free).  This script counts, inside the largest loop of
the named kernel, the FP64 instructions by number of distinct source registers and prints the modelled cycles per
iteration and the pipe utilisation bound 2*N/cycles.

  cuobjdump -sass volcanor_b200/libvolcanor_b200.so > /tmp/lib.sass
  python tools/sass_fp64_cost.py /tmp/lib.sass bs_lattice_kernelILi4ELi1E
"""
import collections
import re
import sys


def kernel_instructions(path, name):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "Function :" in l and name in l)
    end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
    ins = []
    for l in lines[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def analyse(path, name):
    """-> dict(n_fp64, n_other, by={(op, distinct regs): count}, cycles, bound, other=Counter, reuse, loop=(lo, hi))"""
    ins = kernel_instructions(path, name)
    loops = []
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    lo, hi = max(loops, key=lambda x: x[1] - x[0])
    inner = [l for l in loops if l != (lo, hi) and lo <= l[0] and l[1] <= hi]
    if inner:                                   # the ring loop sits inside the tile loop
        lo, hi = max(inner, key=lambda x: x[1] - x[0])
    by = collections.Counter()
    other = collections.Counter()
    reuse = 0
    for a, t in ins:
        if not (lo <= a <= hi):
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0].split(".")[0]
        if op in ("DFMA", "DMUL", "DADD"):
            srcs = [s.strip() for s in t.split(None, 1)[1].split(",")][1:]
            regs = set()
            for s in srcs:
                if ".reuse" in s:
                    reuse += 1
                s = s.replace(".reuse", "").lstrip("-|").rstrip("|")
                if re.fullmatch(r"R\d+", s):
                    regs.add(s)
            by[(op, len(regs))] += 1
        else:
            other[op] += 1
    n = sum(by.values())
    cycles = sum(v * (3 if k[1] == 3 else 2) for k, v in by.items())
    return dict(n_fp64=n, n_other=sum(other.values()), by=dict(by), cycles=cycles, bound=2 * n / cycles, other=other,
                reuse=reuse, loop=(lo, hi))


def main():
    path, name = sys.argv[1], sys.argv[2]
    a = analyse(path, name)
    n, cycles, by, other, reuse, (lo, hi) = a["n_fp64"], a["cycles"], a["by"], a["other"], a["reuse"], a["loop"]
    print(f"{name}: loop 0x{lo:x}..0x{hi:x}, {n} FP64 instructions, {sum(other.values())} others, {reuse} .reuse operands")
    for k in sorted(by):
        print(f"  {k[0]} with {k[1]} distinct source register(s): {by[k]}")
    print(f"  modelled FP64 issue cycles {cycles} = {cycles / n:.3f} per instruction -> pipe utilisation bound {2 * n / cycles:.3f}")
    print("  others: " + ", ".join(f"{k} {v}" for k, v in other.most_common(8)))


if __name__ == "__main__":
    main()
