# A/B of experimental builds of the same sources (VOLCANOR_B200_LIB): usage: bash tools/ab_libs.sh tag:path ...
for TP in "$@"; do TAG=${TP%%:*}; LIBP=${TP#*:}
[ "$LIBP" = "default" ] && unset VOLCANOR_B200_LIB || export VOLCANOR_B200_LIB=$PWD/$LIBP
python bench.py --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab_$TAG.json 2>gpurun_out/bench_ab_$TAG.err
python -c "
import json;j=json.load(open('gpurun_out/bench_ab_$TAG.json'));print('$TAG', '%.4e'%j['value'], '%.2f'%j['roofline']['kernel_ms'], '%.3f'%j['roofline']['pipe_frac'], j['ms_per_step'])"; done
