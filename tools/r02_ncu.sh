# one ncu --set full capture of the grouped dW+Adam kernel (graphs off: every kernel is a launch)
DRVAE_B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwadam -s 5 -c 1 -f -o gpurun_out/r02_dwadam python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
ncu -i gpurun_out/r02_dwadam.ncu-rep --page raw --csv > gpurun_out/r02_dwadam_raw.csv 2>/dev/null
ls -la gpurun_out/
