# dp8192 at N ranks: peer-memory backend (plain loads / NVLS multicast) and the NCCL backend
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload dp8192 --steps 40 --warmup 5 2>>gpurun_out/r02_dp.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['dp8192']; print('$1', d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.3g'%d['value'], d['exchange'], 'launches', d['gpu_launches'], 'ELBO %.3f'%d['losses']['ELBO'], 'tensor frac %.3f'%d['frac_of_bf16_sustained_per_gpu'])
"; }
python bench.py --gpus 1 --workload dp8192 --steps 40 --warmup 5 2>>gpurun_out/r02_dp.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['dp8192']; print('1gpu', 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.3g'%d['value'], d['exchange'], 'ELBO %.3f'%d['losses']['ELBO'], 'tensor frac %.3f'%d['frac_of_bf16_sustained_per_gpu'])
"
DRVAE_B200_DP_BACKEND=peer run peer
DRVAE_B200_DP_BACKEND=peer DRVAE_B200_DP_MULTICAST=1 run peer_mc
DRVAE_B200_DP_BACKEND=nccl run nccl
tail -3 gpurun_out/r02_dp.err
