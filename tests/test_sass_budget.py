"""Static performance guard (no GPU): the SASS of the two sweep kernels keeps the instruction budget DESIGN.md §4 and
profiles/r01h_fp64_operands.md state -- FP64 instructions per ring / per pair, how many of them are DFMAs with three
different source registers (3 issue cycles each instead of 2 on B200), no spills, TMA bulk copies present.  A compiler
flag or a source change that silently costs throughput fails here, before any GPU time is spent."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not Path(CUOBJDUMP).exists(), reason="cuobjdump not available")


@pytest.fixture(scope="module")
def sass(tmp_path_factory):
    import volcanor_b200 as vb
    lib = vb.build_library()
    out = tmp_path_factory.mktemp("sass") / "lib.sass"
    out.write_text(subprocess.run([CUOBJDUMP, "-sass", str(lib)], capture_output=True, text=True, check=True).stdout)
    res = subprocess.run([CUOBJDUMP, "-res-usage", str(lib)], capture_output=True, text=True, check=True).stdout
    return out, res


def _usage(res, name):
    m = re.search(r"Function \S*" + re.escape(name) + r"\S*:\s*\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m, name
    return int(m.group(1)), int(m.group(2))


@pytest.mark.parametrize("kernel,rings,regs_max", [("bs_lattice_kernelILi4ELi1ELi128ELi3ELi2ELb0E", 8, 168),
                                                   ("bs_lattice_kernelILi4ELi2ELi128ELi3ELi2ELb0E", 16, 255),
                                                   ("bs_lattice_kernelILi2ELi2ELi128ELi3ELi2ELb0E", 8, 168),
                                                   ("bs_lattice_kernelILi1ELi3ELi128ELi3ELi2ELb0E", 6, 168)])
def test_lattice_kernel_instruction_budget(sass, kernel, rings, regs_max):
    import sass_fp64_cost as sc
    path, res = sass
    a = sc.analyse(str(path), kernel)
    W = int(re.search(r"ILi(\d)E", kernel).group(1))
    per_ring = (11 * (W + 1) + 50 * W) / W                       # nodes + 2 edges per ring (bs_lattice.cuh)
    assert a["n_fp64"] == round(per_ring * rings), (a["n_fp64"], per_ring * rings)
    three = sum(v for (op, k), v in a["by"].items() if op == "DFMA" and k == 3)
    assert three <= (18 + 1) * rings, (three, rings)             # 9 per edge, 2 edges per ring (+1 of slack)
    assert a["bound"] >= 0.87, a["bound"]
    assert a["other"]["MUFU"] == round((3 * W + 1) / W * rings)  # one seed per node and per edge
    regs, stack = _usage(res, kernel)
    assert stack == 0 and regs <= regs_max, (regs, stack)


def test_dual_lattice_kernel_instruction_budget(sass):
    """The dual form (streamwise edges whose two copies carry different core radii): 33 instead of 25 FP64 instructions on
    the streamwise edge, nodes and spanwise edges unchanged -> (11 (W+1) + 58 W) / W = 71.75 per ring at W = 4 (the flat
    enumeration such a set used to fall back to: 172)."""
    import sass_fp64_cost as sc
    path, res = sass
    kernel = "bs_lattice_kernelILi4ELi1ELi128ELi3ELi2ELb1E"
    a = sc.analyse(str(path), kernel)
    assert a["n_fp64"] == (11 * 5 + 58 * 4) * 2, a["n_fp64"]        # 2 records unrolled, 1 target per thread
    assert a["other"]["MUFU"] == (5 + 4 + 8) * 2                    # nodes + spanwise + two per streamwise edge
    assert a["bound"] >= 0.86, a["bound"]
    regs, stack = _usage(res, kernel)
    assert stack == 0 and regs <= 168, (regs, stack)


def test_flat_kernel_instruction_budget(sass):
    import sass_fp64_cost as sc
    path, res = sass
    a = sc.analyse(str(path), "bs_sweep_kernelILi4ELi128ELi128ELi3ELi3ELb0E")
    assert a["n_fp64"] == 43 * 8, a["n_fp64"]                    # 43 FP64 instructions per pair, 4 targets x 2 sources unrolled
    assert a["bound"] >= 0.90, a["bound"]
    regs, stack = _usage(res, "bs_sweep_kernelILi4ELi128ELi128ELi3ELi3ELb0E")
    assert stack == 0 and regs <= 168, (regs, stack)


def test_tma_bulk_copy_and_no_local_memory_in_hot_kernels(sass):
    path, res = sass
    text = path.read_text()
    for kernel in ("bs_lattice_kernelILi4ELi2ELi128ELi3ELi2ELb0E", "bs_lattice_kernelILi4ELi1ELi128ELi3ELi2ELb1E", "bs_sweep_kernelILi4E"):
        start = text.index(kernel)
        body = text[start:text.index("Function :", start + 10)] if "Function :" in text[start + 10:] else text[start:]
        assert "UBLKCP" in body, kernel                          # cp.async.bulk (1-D TMA) staging of the source tiles
        assert "SYNCS" in body, kernel                           # mbarrier
        assert not re.search(r"\b(LDL|STL)\b", body), kernel     # no local-memory traffic


def test_bit_exact_record_kernels_have_no_contracted_fma(sass):
    """The O(N) kernels whose results must equal the CPU restatement bit for bit (tier 2b wake mutators without a
    division or a square root, the tier 2c accumulate / RHS kernels) contain no DFMA at all: every product and sum is a
    separately rounded DMUL / DADD, as written (`__dmul_rn`, `__dadd_rn`).  Division and square root expand into
    Newton iterations that use DFMA internally and round correctly; kernels containing them are checked on data."""
    path, _ = sass
    text = path.read_text()
    for kernel in ("cp_rhs_kernel", "cp_accumulate_kernel", "cp_copy_field_kernel", "cp_map_gam_kernel", "rec_age_kernel",
                   "rec_accumulate_kernel"):
        start = text.index(kernel)
        nxt = text.find("Function :", start + 10)
        body = text[start:nxt] if nxt > 0 else text[start:]
        assert "DFMA" not in body, kernel
    start = text.index("cp_rhs_kernel")
    body = text[start:text.find("Function :", start + 10)]
    assert len(re.findall(r"\bDMUL\b", body)) == 4 and len(re.findall(r"\bDADD\b", body)) == 2, body   # 3 products + (-1)*, 2 sums
