# A/B with repeats on ONE box: L2 evict-first policy on the optimizer-state loads / stores (DRVAE_B200_DWA_DEBUG=4: default policy)
for rep in 1 2 3; do
for dbg in 4 0; do
  DRVAE_B200_DWA_DEBUG=$dbg python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('debug $dbg: ms/step %.4f e2e %.4g dwadam %.4f' % (d['ms_per_step'], d['e2e']['value'], [r['ms_per_launch'] for r in d['breakdown'] if 'dw_adam' in r['kernel']][0]))"
done
done
timeout 300 python -m pytest tests/test_step_gpu.py -m gpu -x -q -k "fused_adam or graph_replay or train_steps" 2>&1 | tail -2
