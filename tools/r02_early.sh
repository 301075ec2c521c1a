# early part of the grouped dW+Adam launch on a subset of the SMs (DRVAE_B200_DWA_EARLY_SMS; 0 = one launch at the end)
timeout 600 python -m pytest tests/test_step_gpu.py tests/test_shapes_gpu.py -m gpu -x -q -k "fused_adam or graph_replay or ensemble or train_steps or odd_dim or step_kernel" 2>&1 | tail -2
for g in 0 72 84 100 116 132; do
  echo "== early SMs $g"
  DRVAE_B200_DWA_EARLY_SMS=$g python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f e2e %.4g launches %d' % (d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
done
DRVAE_B200_DWA_EARLY_SMS=100 python tools/trace_step.py 2>/dev/null | grep -v "^# g" | tail -22
