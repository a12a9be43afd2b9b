#!/bin/bash
# GPU session P (1 GPU): FORCESPRO-formulation kernel -- throughput of the final build, ncu launch list + full capture (batch 8192, 1024).
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/p_forces_launches.csv python tools/bench_forces.py profile 8192 > gpurun_out/p_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
LIB=motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
for b in 8192 1024; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_forces_solve -s 3 -c 1 -o gpurun_out/p_prof_b$b python tools/bench_forces.py profile $b > gpurun_out/p_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
  ncu -i gpurun_out/p_prof_b$b.ncu-rep --page raw --csv > gpurun_out/p_forces_ncu_full_raw_b$b.csv 2>/dev/null
  python tools/ncu_by_phase.py gpurun_out/p_prof_b$b.ncu-rep $LIB mpc_forces_solve_kernelIfLi1ELb0E forces_core.cuh > gpurun_out/p_forces_by_phase_b$b.txt 2>gpurun_out/p_by_phase_$b.err
  rm -f gpurun_out/p_prof_b$b.ncu-rep
done
head -c 2500 gpurun_out/p_forces_by_phase_b8192.txt; cat gpurun_out/p_by_phase_8192.err | tail -3
