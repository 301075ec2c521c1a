# 1/2/4/8-GPU runs of both multi-GPU workloads (one box): dp8192 (strong scaling, gradient exchange over NVLink peer
# memory) and the default ensemble shard (weak scaling).  Raw JSON lines -> gpurun_out/r02_scale_*.json
NMAX=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $NMAX ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --workload dp8192 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_dp_${n}gpu.json 2>>gpurun_out/r02_scale.err
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_scale_ens_${n}gpu.json 2>>gpurun_out/r02_scale.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload dp8192 --steps 40 --warmup 5 > gpurun_out/r02_scale_dp_${n}gpu.json 2>>gpurun_out/r02_scale.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_scale_ens_${n}gpu.json 2>>gpurun_out/r02_scale.err
  fi
done
python - <<PY
import json, glob
for kind in ("dp", "ens"):
    for n in (1, 2, 4, 8):
        try:
            d = json.loads([l for l in open("gpurun_out/r02_scale_%s_%dgpu.json" % (kind, n)) if l.startswith("{")][-1])
        except Exception as e:
            continue
        if kind == "dp":
            x = d["dp8192"]
            print("dp8192 %d GPU: %.4f ms/step  %.4g samples/s  launches %s  %s  ELBO %.4f" % (n, x["ms_per_step"], x["value"], x.get("gpu_launches"), x.get("exchange"), x["losses"]["ELBO"]))
        else:
            print("ensemble %d GPU: %.4f ms/step  value %.4g  e2e %.4g" % (n, d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
tail -3 gpurun_out/r02_scale.err
