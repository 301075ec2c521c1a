N=${1:-8}
for n in 4 $N; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload dp8192 --steps 40 --warmup 5 2>>gpurun_out/r02_dp.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['dp8192']; print(d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.3g'%d['value'], d['exchange'], 'ELBO %.3f'%d['losses']['ELBO'])
"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_ens_${N}gpu.json 2>>gpurun_out/r02_dp.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_ens_${N}gpu.json").read().strip().splitlines()[-1])
print("ensemble", d["n_gpus"], "ms/step", d["ms_per_step"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["windows_ms_per_step"])
print("dp8192 key:", {k:v for k,v in d.get("dp8192",{}).items() if k in ("ms_per_step","value","exchange")})
PY
tail -2 gpurun_out/r02_dp.err
