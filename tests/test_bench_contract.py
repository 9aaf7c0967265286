"""bench.py's contract, as far as it can run without a GPU: the reference arm (`--impl reference`, pure CPU) prints exactly
one JSON line on stdout with the keys the driver reads; our own arm fails loudly without a CUDA device (no CPU path)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, env=None):
    e = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys(oracle):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--filaments", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout                       # stdout carries the JSON line alone
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["metric"] == "biot_savart_pair_interactions_per_s"
    assert out["unit"] == "pair-interactions/s" and out["higher_is_better"] is True and out["value"] > 0
    assert out["n_gpus"] == 1 and out["steps"] == 1 and out["warmup"] == 0 and out["ms_per_step"] > 0
    assert out["dtype"] == "f64" and out["data"] == "synthetic" and "workload" in out["config"]
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == out["value"] and cb["sample"]
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly(oracle):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--filaments", "20000", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout, r.stderr[-500:])


def test_own_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", "--filaments", "20000", "--no-cpu-baseline", "--no-e2e")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # no number without the CUDA path


def test_reference_arm_under_torchrun_uses_all_host_threads(oracle):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still run on the host's cores (round 1: the
    N >= 2 reference lines ran on one core and the driver's ratios at N = 2, 4, 8 were void)."""
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--filaments", "20000", "--gpus", "2",
             env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    want = len(os.sched_getaffinity(0))
    assert out["cpu_baseline"]["cores"] == want, (out["cpu_baseline"]["cores"], want)
    assert out["n_gpus"] == 2 and set(out["config"]) >= {"workload", "filaments", "targets", "step", "l2"}
