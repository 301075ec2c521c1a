python tools/trace_step.py --models 1 --batch 1024 > gpurun_out/r02_trace_dp1024.txt 2>>gpurun_out/r02_dp.err; cat gpurun_out/r02_trace_dp1024.txt | head -60
for mc in 1; do
DRVAE_B200_DP_MULTICAST=$mc python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload dp8192 --steps 40 --warmup 5 2>>gpurun_out/r02_dp.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['dp8192']; print(d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.3g'%d['value'], d['exchange'], 'ELBO %.3f'%d['losses']['ELBO'])
"
done
DRVAE_B200_DP_MULTICAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tests/dp_worker.py peer readme 8192 6 2>>gpurun_out/r02_dp.err | grep DP
