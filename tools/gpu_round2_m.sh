#!/bin/bash
# GPU session M (1 GPU): final-build artefacts only -- bench line + reference arm, ncu launch list, ncu full captures (batch 1024 / 8192).
# The .ncu-rep files are summarised ON the box (raw page csv + per-phase attribution) and removed: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/m_bench.json 2>> gpurun_out/m_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/m_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 6 --warmup 3 > gpurun_out/m_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
LIB=motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
for b in 1024 8192; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/m_prof_b$b python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch $b > gpurun_out/m_ncu_full_$b.out 2>&1; echo "ncu full $b rc=$?"
  ncu -i gpurun_out/m_prof_b$b.ncu-rep --page raw --csv > gpurun_out/m_ncu_full_raw_b$b.csv 2>/dev/null
  python tools/ncu_by_phase.py gpurun_out/m_prof_b$b.ncu-rep $LIB mpc_warp_solve_kernelIfLi2ELi0ELi0 > gpurun_out/m_by_phase_b$b.txt 2>/dev/null
  rm -f gpurun_out/m_prof_b$b.ncu-rep
done
python - <<'PY'
import json
for f in ("gpurun_out/m_bench.json","gpurun_out/m_bench_ref.json"):
    d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e"%d["e2e"]["value"], d.get("roofline",{}).get("issue_frac"))
PY
ls -la gpurun_out
