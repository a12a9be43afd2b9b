#!/bin/bash
# GPU session Y (1 GPU): full ncu capture of the LAST build at batch 8192 (raw + per-phase pages).
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
LIB=motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
b=8192
timeout 100 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/y_prof_b$b python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch $b > gpurun_out/y_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
ncu -i gpurun_out/y_prof_b$b.ncu-rep --page raw --csv > gpurun_out/y_ncu_full_raw_b$b.csv 2>/dev/null
python tools/ncu_by_phase.py gpurun_out/y_prof_b$b.ncu-rep $LIB mpc_warp_solve_kernelIfLi2ELi0ELi0 > gpurun_out/y_by_phase_b$b.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/y_prof_b$b.ncu-rep | head -40
ncu -i gpurun_out/y_prof_b$b.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,v=r[0],r[2]
for k in ('l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):
    print(k, v[h.index(k)] if k in h else 'n/a')
"
rm -f gpurun_out/y_prof_b$b.ncu-rep
