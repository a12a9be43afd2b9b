#!/bin/bash
# GPU session Q (1 GPU): where do the no_instruction stalls of the CasADi-formulation kernel sit?  Full capture at batch 8192 and 1024, the
# SOURCE page (per-instruction stall-reason samples) kept as csv.
mkdir -p gpurun_out
for b in 8192 1024; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/q_prof_b$b python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch $b > gpurun_out/q_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
  ncu -i gpurun_out/q_prof_b$b.ncu-rep --page source --csv > gpurun_out/q_source_b$b.csv 2>/dev/null
  rm -f gpurun_out/q_prof_b$b.ncu-rep
done
ls -la gpurun_out/q_*
