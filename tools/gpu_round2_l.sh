#!/bin/bash
# GPU session L (1 GPU): sweep epilogue variant -- bench at three batches, gpu tests, racecheck.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
for b in 1024 8192 32768; do
  timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --batch $b > gpurun_out/l_b${b}.json 2>> gpurun_out/l_err.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_*.json")):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f p50 %.4f e2e %.4e"%(d["value"],d["ms_per_step"],d["p50_ms_per_step"],d["e2e"]["value"]))
    except Exception as e: print(f,"ERR",e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/l_pytest.txt
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/l_racecheck.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise or misaligned or long_horizon_and_small" > gpurun_out/l_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -n 2 gpurun_out/l_racecheck.txt; tail -n 2 gpurun_out/l_racecheck.out
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/l_memcheck.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise or misaligned or long_horizon_and_small or infeasible" > gpurun_out/l_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -n 2 gpurun_out/l_memcheck.txt
