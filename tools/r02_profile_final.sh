# Final round-2 profile pass on ONE GPU (numbers taken under ncu are never bench values)
set -x
export DRVAE_B200_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
# graphs off + per-launch serialisation under ncu: 19 matching launches per step (17 GEMMs + the two parts of the grouped dW+Adam launch);
# the decoder-loss GEMM is the 13th, the early dW+Adam part follows the decoder dX GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwadam_kernel|gemm_tc_kernel" --launch-skip 69 -c 7 -f \
  -o gpurun_out/r02_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_top.ncu-rep --page raw --csv > gpurun_out/r02_top_raw.csv 2>/dev/null
unset DRVAE_B200_GRAPH
python tools/trace_step.py > gpurun_out/r02_trace_ens32.txt 2>&1
python tools/trace_step.py --models 1 > gpurun_out/r02_trace_drvae150.txt 2>&1
DRVAE_B200_STEPK=1 python tools/trace_step.py > gpurun_out/r02_trace_ens32_stepk.txt 2>&1
DRVAE_B200_STEPK=1 python tools/trace_step.py --models 1 > gpurun_out/r02_trace_drvae150_stepk.txt 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -c 400 gpurun_out/r02_bench_reference.json
