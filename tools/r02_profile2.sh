# ncu --set full of the decoder-loss GEMM and the decoder dX GEMM of the 4th step (19 GEMM launches per step in enqueue order:
# enc h0, enc head, z3 h0, z3 head, dz1 h0, dz1 head, T head, dz1 dx.head, dz1 dx.h0, z3 dx.head, z3 dx.h0, dec h0,
# decloss, dec dx.head, dec dx.h0, T dx, enc dx.head  [the per-step count is printed by the launch list])
export DRVAE_B200_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip ${1:-63} -c ${2:-6} -f \
  -o gpurun_out/r02_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_gemm.log 2>&1
ncu -i gpurun_out/r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_gemm_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_gemm_raw.csv')))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60], r[ix['gpu__time_duration.sum']], r[ix['launch__grid_size']], r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])
PY
