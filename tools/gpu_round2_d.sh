#!/bin/bash
# GPU session D: the whole -m gpu suite, the bench line, ncu launch list + full capture (batch 1024 and 8192), racecheck re-run,
# closed-loop status histogram of config 1b.  Every command under its own timeout.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/d_pytest.txt
timeout 600 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/d_bench_ref.json 2>> gpurun_out/d_bench.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/d_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 6 --warmup 3 > gpurun_out/d_ncu_launches.out 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/d_prof_b1024 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 > gpurun_out/d_ncu_full.out 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpc_warp_solve -s 3 -c 1 -o gpurun_out/d_prof_b8192 python bench.py --no-cpu-baseline --no-extra --steps 3 --warmup 3 --batch 8192 > gpurun_out/d_ncu_full8192.out 2>&1; echo "ncu full 8192 rc=$?"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/d_racecheck_tests.txt python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or stepwise or misaligned" > gpurun_out/d_racecheck_tests.out 2>&1; echo "racecheck rc=$?"; tail -n 3 gpurun_out/d_racecheck_tests.txt
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/d_memcheck_tests.txt python -m pytest tests/test_gpu_parity.py tests/test_forces_model.py -m gpu -x -q -k "ragged or stepwise or misaligned or infeasible or forces or refinement" > gpurun_out/d_memcheck_tests.out 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/d_memcheck_tests.txt
timeout 300 python - <<'PY' > gpurun_out/d_closed_loop.txt 2>&1
import numpy as np, time, torch
import mpc_b200
from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
for name, N, B in (("ZAM_Over-1_1_LF", 30, 1024), ("ZAM_Over-1_1_LF", 10, 1024), ("USA_Lanker-2_18_T-1_LF", 50, 1024), ("ZAM_Over-1_1_CA", 30, 1024)):
    sc = mpc_b200.load_scenario(name)
    x0 = mpc_b200.make_batch(name, B, N, 7)[1]
    for prec in ("f32", "f64"):
        for wd in (0, 1):
            opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision=prec, max_batch=B, warm_duals=wd, max_iter=200)
            opt.optimize_batch(x0[:64])
            torch.cuda.synchronize(); t0 = time.perf_counter()
            tr, ct, st, it = opt.optimize_batch(x0)
            dt = time.perf_counter() - t0
            T = st.shape[1]
            print(f"{name} N={N} {prec} warm_duals={wd}: {B*T/dt/1e6:.2f} M MPC-steps/s, status {dict(zip(*[a.tolist() for a in np.unique(st, return_counts=True)]))}, iters step0 {it[:,0].mean():.1f} warm {it[:,1:].mean():.2f} max {it.max()}", flush=True)
PY
cat gpurun_out/d_closed_loop.txt
python - <<'PY'
import json
for f in ("gpurun_out/d_bench.json","gpurun_out/d_bench_ref.json"):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e"%d["e2e"]["value"], d.get("parity"))
        for k,v in (d.get("extra") or {}).items(): print("   ",k,"%.3e"%v["solves_per_s"],v["converged"],v.get("parity"))
    except Exception as e: print(f,"ERR",e)
PY
