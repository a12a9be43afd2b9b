#!/bin/bash
# GPU session F: layout refactor (aligned KKT record, 128-bit loads, shuffle-free forward sweep) -- tests + bench at three batches.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/f_pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 > gpurun_out/f_b1024.json 2> gpurun_out/f_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 100 --batch 8192 > gpurun_out/f_b8192.json 2>> gpurun_out/f_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 50 --batch 32768 > gpurun_out/f_b32768.json 2>> gpurun_out/f_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/f_prof_b8192 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch 8192 > gpurun_out/f_ncu_full8192.out 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/f_b1024.json","gpurun_out/f_b8192.json","gpurun_out/f_b32768.json"):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e (%.4f ms)"%(d["e2e"]["value"],d["e2e"]["ms_per_step"]), "frac %.3f"%d["roofline"]["frac"], d["config"]["converged"], d["config"]["mean_sqp_iters"])
    except Exception as e: print(f,"ERR",e)
PY
