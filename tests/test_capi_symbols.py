"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/volcanor_b200.h declares (no compute calls without a GPU)."""
import ctypes

import pytest

import volcanor_b200 as vb


def test_header_declares_expected_entry_points():
    names = set(vb.DECLARED_SYMBOLS)
    for must in ["vlc_create", "vlc_destroy", "vlc_set_sources", "vlc_vind", "vlc_vind_dev",
                 "vlc_vind_onNwake_byRotor", "vlc_vind_onFwake_byRotor", "vlc_rotor_vind_bywing",
                 "vlc_rotor_vind_bywake", "vlc_rotor_vind_bywing_boundVortices", "vlc_rotor_calcAIC",
                 "vlc_rotor_solve", "vlc_convect_dev", "vlc_dissipate_dev", "vlc_strain_dev",
                 "vlc_pack_lattice_dev", "vlc_measure_fp64_peak"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    vb.build_library()
    lib = ctypes.CDLL(str(vb.lib_path()))
    missing = [s for s in vb.DECLARED_SYMBOLS if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_every_declared_symbol():
    lib = vb.load_library()
    assert sorted(lib._vlc_signatures) == vb.DECLARED_SYMBOLS


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VlcError):
        vb.Context(0)


def test_product_does_not_reference_oracle():
    """The product package must not import, link or call anything under oracle/."""
    import pathlib
    import subprocess
    root = pathlib.Path(vb.__file__).resolve().parent
    for p in list(root.rglob("*.py")) + list(root.rglob("*.cu*")):
        txt = p.read_text()
        assert "pyoracle" not in txt and "vlc_oracle" not in txt and "orc_" not in txt, p
    out = subprocess.run(["ldd", str(vb.lib_path())], capture_output=True, text=True).stdout
    assert "oracle" not in out
    # outside tests/ only bench.py (cpu_baseline / --impl reference) and __graft_entry__ (build, smoke) may use the oracle
    repo = root.parent
    for d in ("tools", "include", "fortran"):
        for p in (repo / d).rglob("*"):
            if p.is_file() and p.suffix in (".py", ".sh", ".h", ".hpp", ".f90", ".cu", ".c", ".cpp"):
                txt = p.read_text()
                assert "pyoracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, p


class _RecordingLib:
    """Stand-in for the CDLL: every vlc_* call forces a garbage collection, churns the allocator, and then reads the
    doubles its pointer arguments address -- what the C side would see.  A binding that passes the address of a temporary
    (round 1: `_ptr(_f64(list))`) hands over freed, reused memory and is caught here without a GPU."""

    def __init__(self, sizes):
        self.sizes, self.seen = sizes, {}

    def __getattr__(self, name):
        if not name.startswith("vlc_"):
            raise AttributeError(name)

        def call(h, *args):
            import gc

            import numpy as np
            gc.collect()
            junk = [np.full(n, -7.0) for n in (2, 3, 3, 3, 16, 104, 650) for _ in range(8)]   # reuse freed blocks
            ptrs = [a for a in args if isinstance(a, int) and a > 4096]
            self.seen[name] = [np.ctypeslib.as_array((ctypes.c_double * n).from_address(p)).copy()
                               for p, n in zip(ptrs, self.sizes[name])]
            del junk
            return 0
        return call


def test_binding_keeps_converted_arrays_alive_across_the_call():
    """list / non-contiguous / non-float64 arguments are converted to temporaries; every one of them
    must outlive the foreign call (vlc_rotor_set_frame got shaftAxis == hubCoords; classdef.f90:1010-1012 reads both)."""
    import numpy as np
    sizes = {"vlc_rotor_set_frame": [3, 3], "vlc_rotor_put_wing": [208], "vlc_rotor_put_wing_gam": [6],
             "vlc_rotor_put_nwake": [100], "vlc_rotor_put_fwake": [26], "vlc_rotor_put_pfwake": [39],
             "vlc_rotor_put_pfwake_helix": [2], "vlc_rotor_put_sections": [16], "vlc_rotor_put_wakevel": [9, 6]}
    c = object.__new__(vb.Context)
    c.lib, c.h = _RecordingLib(sizes), None
    c.rotor_set_frame(0, [0.0, 0.6, 0.8], [1.0, 2.0, 3.0])
    sa, hc = c.lib.seen["vlc_rotor_set_frame"]
    assert sa.tolist() == [0.0, 0.6, 0.8] and hc.tolist() == [1.0, 2.0, 3.0]
    c.rotor_set_frame(0, (0, 0, 1), np.array([4, 5, 6], dtype=np.int32))       # tuple + wrong dtype
    sa, hc = c.lib.seen["vlc_rotor_set_frame"]
    assert sa.tolist() == [0.0, 0.0, 1.0] and hc.tolist() == [4.0, 5.0, 6.0]
    rng = np.random.default_rng(0)
    for meth, name, n, pre in [("rotor_put_wing", "vlc_rotor_put_wing", 208, (0, 0)),
                               ("rotor_put_wing_gam", "vlc_rotor_put_wing_gam", 6, (0, 0)),
                               ("rotor_put_nwake", "vlc_rotor_put_nwake", 100, (0, 0)),
                               ("rotor_put_fwake", "vlc_rotor_put_fwake", 26, (0, 0)),
                               ("rotor_put_pfwake", "vlc_rotor_put_pfwake", 39, (0, 0)),
                               ("rotor_put_pfwake_helix", "vlc_rotor_put_pfwake_helix", 2, (0, 0)),
                               ("rotor_put_sections", "vlc_rotor_put_sections", 16, (0, 0))]:
        want = rng.uniform(-1, 1, n)
        for arg in (want.tolist(), want.astype(np.float32).astype(np.float64)[::-1][::-1], np.asfortranarray(want)):
            getattr(c, meth)(*pre, arg)
            assert np.array_equal(c.lib.seen[name][0], np.asarray(arg, dtype=np.float64)), (meth, type(arg))
    vn, vf = rng.uniform(size=9), rng.uniform(size=6)
    c.rotor_put_wakevel(0, 0, 0, vn.tolist(), vf.tolist())
    assert np.array_equal(c.lib.seen["vlc_rotor_put_wakevel"][0], vn)
    assert np.array_equal(c.lib.seen["vlc_rotor_put_wakevel"][1], vf)
