# grouped dW+Adam: epilogue-warp variants (profiles/r02_experiments.md)
for v in "" _ew8 _ew8r _ew4r; do
  export DRVAE_B200_LIB=$PWD/drvae_b200/lib/libdrvae_b200$v.so
  echo "== variant '$v'"
  timeout 600 python -m pytest tests/test_step_gpu.py tests/test_shapes_gpu.py -m gpu -x -q -k "fused_adam or graph_replay or ensemble or train_steps or odd_dim" 2>&1 | tail -2
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f e2e %.4g' % (d['ms_per_step'], d['e2e']['value']), [(r['kernel'], round(r['ms_per_launch'],4)) for r in d['breakdown'][:3]])"
done
