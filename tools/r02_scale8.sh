# final-code refresh of the 4- and 8-GPU points (one 8-GPU box)
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload dp8192 --steps 40 --warmup 5 > gpurun_out/r02_scale_dp_${n}gpu.json 2>>gpurun_out/r02_scale.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_ens_8gpu.json 2>>gpurun_out/r02_scale.err
python - <<PY
import json
for kind, n in (("dp", 4), ("dp", 8), ("ens", 8)):
    d = json.loads([l for l in open("gpurun_out/r02_scale_%s_%dgpu.json" % (kind, n)) if l.startswith("{")][-1])
    if kind == "dp":
        x = d["dp8192"]; print("dp8192 %d GPU: %.4f ms/step  %.4g samples/s  ELBO %.4f" % (n, x["ms_per_step"], x["value"], x["losses"]["ELBO"]))
    else:
        print("ensemble %d GPU: %.4f ms/step  value %.4g  e2e %.4g" % (n, d["ms_per_step"], d["value"], d["e2e"]["value"]), "dp8192 key:", {k: v for k, v in d.get("dp8192", {}).items() if k in ("ms_per_step", "value")})
PY
tail -2 gpurun_out/r02_scale.err
