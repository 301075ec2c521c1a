# Round-2 profile pass on ONE GPU (numbers taken under ncu are never bench values):
#  1. ncu launch list of the default bench command (graphs off: every kernel of the step is a launch)
#  2. ncu --set full of the three kernels that carry the step: grouped dW+Adam, decoder-loss GEMM, decoder dX GEMM
#  3. per-kernel event timing + timeline of one step (tools/trace_step.py), default schedule and the persistent step kernel
set -x
export DRVAE_B200_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwadam_kernel|gemm_tc_kernelILi(4|2)E" --launch-skip 18 -c 6 -f \
  -o gpurun_out/r02_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_top.ncu-rep --page raw --csv > gpurun_out/r02_top_raw.csv 2>/dev/null
unset DRVAE_B200_GRAPH
python tools/trace_step.py > gpurun_out/r02_trace_ens32.txt 2>&1
python tools/trace_step.py --models 1 > gpurun_out/r02_trace_drvae150.txt 2>&1
DRVAE_B200_STEPK=1 python tools/trace_step.py > gpurun_out/r02_trace_ens32_stepk.txt 2>&1
DRVAE_B200_STEPK=1 python tools/trace_step.py --models 1 > gpurun_out/r02_trace_drvae150_stepk.txt 2>&1
for k in pvae vfae; do python tools/trace_step.py --kind $k > gpurun_out/r02_trace_${k}32.txt 2>&1; done
DRVAE_B200_STEPK=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_stepk.json 2> gpurun_out/r02_bench_stepk.err
ls -la gpurun_out | tail -15
