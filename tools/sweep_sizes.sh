# pairs/s and timesteps/s of bench.py's step against wake size (BASELINE.json's named sizes) on one GPU:
#   bash tools/sweep_sizes.sh [--graph] 10000 100000 1000000
EXTRA=""; if [ "$1" = "--graph" ]; then EXTRA="--graph"; shift; fi
for N in "$@"; do
python bench.py --filaments $N --steps 5 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/bench_size_$N$EXTRA.json 2>gpurun_out/bench_size_$N$EXTRA.err || tail -5 gpurun_out/bench_size_$N$EXTRA.err
python -c "
import json;j=json.load(open('gpurun_out/bench_size_$N$EXTRA.json'));r=j['roofline'];c=j['config']
print(c['filaments'], c['targets'], '%.4e'%j['value'], '%.3f'%j['ms_per_step'], '%.2f'%j['timesteps_per_s'], '%.3f'%r['kernel_ms'], '%.3f'%r['pipe_frac'], c['sources'].get('strip_width'), j['gpu_launches'], c['launch'][:5])"; done
