"""The C++ mirror of the reference's procedure interface (include/volcanor_b200.hpp) over the C ABI, exercised by a
small native program (tests/native/cpp_aic.cpp) the way tests/wing1x3_test.f90 exercises rotor%calcAIC."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import refgeom

NATIVE = Path(__file__).resolve().parent / "native"


def _build():
    r = subprocess.run(["make", "-C", str(NATIVE)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return NATIVE / "cpp_aic"


def test_cpp_mirror_compiles_links_and_fails_loudly_without_a_gpu():
    exe = _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(exe), "--solve-before-calcAIC"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_aic_kat_and_error_behaviour(tmp_path):
    exe = _build()
    rec = refgeom.wing1x3()
    f = tmp_path / "wing1x3.bin"
    np.ascontiguousarray(rec).tofile(f)
    r = subprocess.run([str(exe), str(f), "1", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [np.array(l.split(), dtype=float) for l in r.stdout.strip().splitlines()]
    A, g = np.array(rows[:3]), rows[3]
    assert np.max(np.abs(A / refgeom.AIC_WING1X3 - 1.0)) < 5e-13          # tests/wing1x3_test.f90:83-85, all 15 digits
    assert np.max(np.abs(A @ g - np.arange(1.0, 4.0))) < 1e-12
    r = subprocess.run([str(exe), "--solve-before-calcAIC"], capture_output=True, text=True)
    assert r.returncode == 3 and "before vlc_rotor_calcAIC" in r.stdout   # the reference would `error stop`
