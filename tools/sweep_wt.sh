# W x T sweep of the shared-node kernel (profiles/r01e_wt_sweep.md); usage: bash tools/sweep_wt.sh "W T" "W T" ...
for WT in "$@"; do set -- $WT
python bench.py --steps 2 --no-cpu-baseline --no-e2e --lat-w $1 --lat-t $2 > gpurun_out/bench_W$1T$2.json 2>gpurun_out/bench_W$1T$2.err
python -c "
import json;j=json.load(open('gpurun_out/bench_W$1T$2.json'));print($1,$2, '%.4e'%j['value'], '%.2f'%j['roofline']['kernel_ms'], '%.3f'%j['roofline']['pipe_frac'])"; done
