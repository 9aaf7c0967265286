python -m pytest tests/test_gpu_lattice.py tests/test_gpu_case.py -m gpu -x -q > gpurun_out/t3.log 2>&1; tail -4 gpurun_out/t3.log
for WT in "1 3" "2 1" "2 2" "2 3" "3 1" "3 2" "4 1" "4 2"; do set -- $WT
python bench.py --steps 3 --no-cpu-baseline --no-e2e --lat-w $1 --lat-t $2 > gpurun_out/bench_W$1T$2.json 2>gpurun_out/bench_W$1T$2.err
python -c "
import json;j=json.load(open('gpurun_out/bench_W$1T$2.json'));print($1,$2, '%.4e'%j['value'], '%.2f'%j['roofline']['kernel_ms'], '%.3f'%j['roofline']['pipe_frac'])"; done
