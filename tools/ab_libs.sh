# A/B of experimental builds of the same sources (VOLCANOR_B200_LIB): usage: bash tools/ab_libs.sh "tag:path[:bench flags]" ...
# path = default for the in-tree library
for TP in "$@"; do IFS=: read -r TAG LIBP EXTRA <<< "$TP"
[ "$LIBP" = "default" ] && unset VOLCANOR_B200_LIB || export VOLCANOR_B200_LIB=$PWD/$LIBP
python bench.py --steps 3 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/bench_ab_$TAG.json 2>gpurun_out/bench_ab_$TAG.err
python -c "
import json;j=json.load(open('gpurun_out/bench_ab_$TAG.json'));print('$TAG', '%.4e'%j['value'], '%.2f'%j['roofline']['kernel_ms'], '%.3f'%j['roofline']['pipe_frac'], j['ms_per_step'])"; done
