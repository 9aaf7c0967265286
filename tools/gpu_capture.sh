#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, bench lines, ncu launch list + one full capture of the
# dominant kernel.  Everything lands in gpurun_out/ (scratch); tools/summarize_ncu.py turns the ncu files
# into the tracked summaries under profiles/.
#   usage: tools/gpu_capture.sh <tag> [skip-tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/smi_$TAG.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rf --durations=15 > $OUT/pytest_gpu_$TAG.log 2>&1
  echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
  tail -25 $OUT/pytest_gpu_$TAG.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -3 $OUT/bench_$TAG.err
# launch list of the same command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
  > $OUT/bench_under_ncu_$TAG.log 2>&1
# one full capture of the dominant kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bs_lattice -s 4 -c 1 \
  -f -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > $OUT/ncu_full_$TAG.log 2>&1
# the reference cases through the C ABI: wake resident (tier 2b) and with the collocation-point stage on the device (tier 2c)
for CASE in katzNplotkin_AR04 elevateTest caradonna; do
  timeout 300 python tests/tools/run_case_native.py $CASE --resident --cp > $OUT/${CASE}_resident_cp_$TAG.log 2>&1
  tail -1 $OUT/${CASE}_resident_cp_$TAG.log
done
ls -la $OUT | tail -20
grep -h -o '"value": [0-9.e+]*' $OUT/bench_$TAG.json | head -3
