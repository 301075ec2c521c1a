# schedule variants (DRVAE_B200_SCHED bits: 1 noise on side stream, 4 classifier weight gradient off the chain, 8 dW GEMMs of the
# gradient path on their own stream, 16 classifier forward split from T_post, 32 classifier input gradient inside T_back)
for s in 13 29 45 61; do
  echo "== sched $s"
  DRVAE_B200_SCHED=$s python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f e2e %.4g' % (d['ms_per_step'], d['e2e']['value']))"
done
DRVAE_B200_SCHED=29 python tools/trace_step.py 2>/dev/null | grep -v "^# g"
