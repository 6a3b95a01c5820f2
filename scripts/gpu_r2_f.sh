#!/bin/bash
# round 2, call F (1 GPU): profiler evidence.  (a) every launch of the bench command with its device time, (b) ncu --set full of
# the step kernels of LUBM-2560 Q1 (expand, fused filter chain) and Q7, and of the bulk-copy seed.  The resident light-query server
# is switched off under the profiler (ncu serialises kernels; a resident kernel would be replayed until it idles out).
mkdir -p gpurun_out
export WK_RESIDENT=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|seed_bulk" -s 4 -c 8 -o gpurun_out/r2f_prof_q1 python scripts/expand_bench.py --scale 2560 --reps 3 --query 1 > gpurun_out/r2f_ncu_q1.log 2>&1
echo "ncu q1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel" -s 4 -c 8 -o gpurun_out/r2f_prof_q7 python scripts/expand_bench.py --scale 2560 --reps 3 --query 7 > gpurun_out/r2f_ncu_q7.log 2>&1
echo "ncu q7 rc=$?"
ls -la gpurun_out/r2f_*
