# dp8192 at 8 and 4 ranks: gradient sum by plain peer loads (every rank reads every rank's gradient) vs one switch-reduced
# load through the NVLS multicast mapping (multimem.ld_reduce)
for n in 8 4; do
for mc in 0 1; do
DRVAE_B200_DP_MULTICAST=$mc python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload dp8192 --steps 40 --warmup 5 2>>gpurun_out/r02_dp_mc.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['dp8192']; print('$n ranks multicast $mc:', 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.4g'%d['value'], d['exchange'], 'ELBO %.4f'%d['losses']['ELBO'])
"
done
done
tail -2 gpurun_out/r02_dp_mc.err
