#!/bin/bash
# GPU session X (1 GPU): ncu artefacts of the LAST build (after the r^1.5 rate-exit rule): launch list of the bench command and one
# full capture at batch 1024 with raw + per-phase pages.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/x_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 6 --warmup 3 > gpurun_out/x_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
LIB=motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
b=1024
timeout 100 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/x_prof_b$b python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch $b > gpurun_out/x_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
ncu -i gpurun_out/x_prof_b$b.ncu-rep --page raw --csv > gpurun_out/x_ncu_full_raw_b$b.csv 2>/dev/null
python tools/ncu_by_phase.py gpurun_out/x_prof_b$b.ncu-rep $LIB mpc_warp_solve_kernelIfLi2ELi0ELi0 > gpurun_out/x_by_phase_b$b.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/x_prof_b$b.ncu-rep | head -40
rm -f gpurun_out/x_prof_b$b.ncu-rep
