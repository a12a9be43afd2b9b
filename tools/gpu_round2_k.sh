#!/bin/bash
# GPU session K (1 GPU): A/B of the warp sync at the end of trial_merit, the gpu tests that changed, racecheck re-run.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
for i in 1 2; do
for b in 1024 8192; do
  timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --batch $b > gpurun_out/k_sync_b${b}_$i.json 2>> gpurun_out/k_err.txt
  MPCB200_LIB=$PWD/build_variants/libmpcb200_nosync.so timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --batch $b > gpurun_out/k_nosync_b${b}_$i.json 2>> gpurun_out/k_err.txt
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/k_*.json")):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f p50 %.4f"%(d["value"],d["ms_per_step"],d["p50_ms_per_step"]))
    except Exception as e: print(f,"ERR",e)
PY
timeout 900 python -m pytest tests -m gpu -x -q -k "per_problem_scenarios or dual_block or closed_loop_float32 or noised or n128 or misaligned" > gpurun_out/k_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/k_pytest.txt
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/k_racecheck.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise or misaligned or per_problem_scenarios or dual_block" > gpurun_out/k_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -n 2 gpurun_out/k_racecheck.txt; tail -n 2 gpurun_out/k_racecheck.out
