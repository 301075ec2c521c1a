# row-kernel launch-bound variants (profiles/r02_experiments.md): timeline of the row kernels + step time per variant
for v in "" _lb5 _lb4 _lbm; do
  export DRVAE_B200_LIB=$PWD/drvae_b200/lib/libdrvae_b200$v.so
  echo "== variant '$v'"
  python tools/trace_step.py 2>/dev/null | grep -E "sample_q1|T_post|z3_post|pz1_post|z3_back|clf_back|T_back|q_back|dw_adam_all|# drvae"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f e2e %.4g' % (d['ms_per_step'], d['e2e']['value']))"
done
