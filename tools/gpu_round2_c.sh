#!/bin/bash
# GPU session C: PDL refinement launch A/B, compute-sanitizer memcheck + racecheck on smoke() and the ragged / misaligned tests.
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --opt refine_f64=0 > gpurun_out/c_norefine_b1024_$i.json 2> gpurun_out/c_err.txt
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 > gpurun_out/c_refine_b1024_$i.json 2>> gpurun_out/c_err.txt
done
timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 50 --batch 8192 > gpurun_out/c_refine_b8192.json 2>> gpurun_out/c_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c_*.json")):
    try:
        d=json.load(open(f))
        print(f, "value %.3e ms %.4f p50 %.4f e2e %.3e (%.4f ms)"%(d["value"],d["ms_per_step"],d["p50_ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/c_memcheck_smoke.txt python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/c_racecheck_smoke.txt python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_racecheck_smoke.out 2>&1; echo "racecheck smoke rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/c_memcheck_tests.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or misaligned or stepwise or infeasible" > gpurun_out/c_memcheck_tests.out 2>&1; echo "memcheck tests rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --log-file gpurun_out/c_racecheck_tests.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise" > gpurun_out/c_racecheck_tests.out 2>&1; echo "racecheck tests rc=$?"
for f in gpurun_out/c_memcheck_smoke.txt gpurun_out/c_racecheck_smoke.txt gpurun_out/c_memcheck_tests.txt gpurun_out/c_racecheck_tests.txt; do echo "== $f"; tail -n 4 $f; done
tail -n 3 gpurun_out/c_memcheck_tests.out gpurun_out/c_racecheck_tests.out
