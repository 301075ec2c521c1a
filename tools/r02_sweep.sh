for dbg in 2 3; do
  DRVAE_B200_DWA_DEBUG=$dbg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_sweep_q.json 2> gpurun_out/r02_sweep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_sweep_q.json").read().strip().splitlines()[-1])
print("dwa_debug ${dbg}: ms/step %.4f" % (d["ms_per_step"]), [(r["kernel"], r["ms_per_launch"]) for r in d["breakdown"][:1]])
PY
DRVAE_B200_DWA_DEBUG=$dbg python tools/trace_step.py 2>>gpurun_out/r02_sweep.err | head -1
done
DRVAE_B200_DWA_DEBUG=2 timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -x -q -k "fused_adam or graph_replay" 2>&1 | tail -3
