timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; tail -8 gpurun_out/r02_pytest5.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_sweep_x.json 2> gpurun_out/r02_sweep.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_sweep_x.json").read().strip().splitlines()[-1])
print("ms/step %.4f e2e %.3g launches %d" % (d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), [ (r["kernel"], r["ms_per_launch"]) for r in d["breakdown"][:2]])
PY
python tools/trace_step.py 2>>gpurun_out/r02_sweep.err | head -1
