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
