#!/bin/bash
# GPU session T (1 GPU): final-build ncu artefacts of the CasADi-formulation kernel (after the final-phase extrapolation): launch list,
# full captures at batch 1024 / 8192 with raw + per-phase pages.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/t_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 6 --warmup 3 > gpurun_out/t_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
LIB=motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
for b in 1024 8192; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/t_prof_b$b python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch $b > gpurun_out/t_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
  ncu -i gpurun_out/t_prof_b$b.ncu-rep --page raw --csv > gpurun_out/t_ncu_full_raw_b$b.csv 2>/dev/null
  python tools/ncu_by_phase.py gpurun_out/t_prof_b$b.ncu-rep $LIB mpc_warp_solve_kernelIfLi2ELi0ELi0 > gpurun_out/t_by_phase_b$b.txt 2>/dev/null
  rm -f gpurun_out/t_prof_b$b.ncu-rep
done
head -12 gpurun_out/t_by_phase_b1024.txt
