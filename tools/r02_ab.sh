# A/B with repeats on ONE box: early dW+Adam launch off / on  x  register cap of the dW+Adam kernel (80 / uncapped)
for rep in 1 2; do
for v in "" _mb1; do
for g in 0 92; do
  DRVAE_B200_LIB=$PWD/drvae_b200/lib/libdrvae_b200$v.so DRVAE_B200_DWA_EARLY_SMS=$g python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lib \"$v\" early $g: ms/step %.4f e2e %.4g dwadam %.4f' % (d['ms_per_step'], d['e2e']['value'], [r['ms_per_launch'] for r in d['breakdown'] if 'dw_adam' in r['kernel']][0]))"
done
done
done
