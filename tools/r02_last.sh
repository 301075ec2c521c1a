# last GPU call of the round: PVAE with the early dW+Adam part (parity subset + timeline), ncu --set full of the persistent step kernel
timeout 300 python -m pytest tests/test_step_gpu.py -m gpu -x -q -k "pvae" 2>&1 | tail -2
for g in 0 92; do DRVAE_B200_DWA_EARLY_SMS=$g python tools/trace_step.py --kind pvae 2>/dev/null | grep "^# pvae"; done
DRVAE_B200_STEPK=1 DRVAE_B200_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_stepk \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_stepk.log 2>&1
ncu -i gpurun_out/r02_stepk.ncu-rep --page raw --csv > gpurun_out/r02_stepk_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_stepk_raw.csv')))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[2:]:
    print(r[ix['Kernel Name']][:50], r[ix['gpu__time_duration.sum']], r[ix['launch__grid_size']], r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']], r[ix['smsp__issue_active.avg.pct_of_peak_sustained_active']])
PY
