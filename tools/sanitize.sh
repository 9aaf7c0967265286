#!/bin/bash
# compute-sanitizer passes over the GPU tests that exercise every kernel family of the library (run on one B200 through gpurun;
# logs under gpurun_out/).  Each pass is bounded by its own timeout.
mkdir -p gpurun_out
export PYTHONPATH=.
run() {  # name, tool, timeout, pytest args...
  local name=$1 tool=$2 tmo=$3; shift 3
  timeout "$tmo" compute-sanitizer --tool "$tool" --error-exitcode 99 --log-file "gpurun_out/sanitizer_${name}.log" \
    python -m pytest -m gpu -q -x "$@" > "gpurun_out/sanitizer_${name}.pytest.log" 2>&1
  echo "$name ($tool): rc=$? | $(tail -n 1 gpurun_out/sanitizer_${name}.pytest.log) | $(grep -c 'ERROR SUMMARY: 0 errors\|RACECHECK SUMMARY: 0 hazards' gpurun_out/sanitizer_${name}.log) clean summaries | $(grep -h 'SUMMARY' gpurun_out/sanitizer_${name}.log | sort | uniq -c | tr '\n' ';')"
}
run mem_parity   memcheck  900 tests/test_gpu_parity.py -k "not full_size"
run race_parity  racecheck 900 tests/test_gpu_parity.py -k "flat_random or empty or skipped"
run mem_lattice  memcheck  900 tests/test_gpu_lattice.py tests/test_gpu_wake_sweep.py -k "not full_size"
run mem_resident memcheck  900 tests/test_gpu_resident.py tests/test_zz_gpu_cp_stage.py tests/test_gpu_group.py
run mem_case     memcheck  900 tests/test_case_driver.py -k "golden_history or sub_iterations or dual_form"
run race_lattice racecheck 900 tests/test_gpu_lattice.py -k "dual or tail or degenerate or far_field"
run sync_lattice synccheck 600 tests/test_gpu_lattice.py -k "dual or tail or degenerate"
