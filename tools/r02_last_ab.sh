for rep in 1 2; do
for cfg in "13 92" "29 92" "13 84" "13 100"; do
  set -- $cfg
  DRVAE_B200_SCHED=$1 DRVAE_B200_DWA_EARLY_SMS=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('sched $1 early $2: ms/step %.4f e2e %.4g' % (d['ms_per_step'], d['e2e']['value']))"
done
done
