#!/bin/bash
# GPU session J (1 GPU): final artefacts -- whole -m gpu suite, bench line (+ reference arm), ncu launch list + full captures (batch 1024 / 8192),
# compute-sanitizer memcheck + racecheck.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/j_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/j_bench.json 2>> gpurun_out/j_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/j_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 6 --warmup 3 > gpurun_out/j_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/j_prof_b1024 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 > gpurun_out/j_ncu_full.out 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/j_prof_b8192 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch 8192 > gpurun_out/j_ncu_full8192.out 2>&1; echo "ncu full 8192 rc=$?"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/j_racecheck.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise or misaligned or per_problem_scenarios or dual_block" > gpurun_out/j_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -n 2 gpurun_out/j_racecheck.txt; tail -n 2 gpurun_out/j_racecheck.out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/j_memcheck.txt python -m pytest tests/test_gpu_parity.py tests/test_forces_model.py -m gpu -x -q -k "ragged or stepwise or misaligned or infeasible or forces or refinement or per_problem_scenarios or dual_block or step0" > gpurun_out/j_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -n 2 gpurun_out/j_memcheck.txt; tail -n 2 gpurun_out/j_memcheck.out
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/j_memcheck_smoke.txt python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_memcheck_smoke.out 2>&1; tail -n 1 gpurun_out/j_memcheck_smoke.txt
python - <<'PY'
import json
for f in ("gpurun_out/j_bench.json","gpurun_out/j_bench_ref.json"):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e"%d["e2e"]["value"], d.get("parity"))
        for k,v in (d.get("extra") or {}).items(): print("   ",k,"%.3e"%v["solves_per_s"],v["converged"],v.get("parity"))
    except Exception as e: print(f,"ERR",e)
PY
