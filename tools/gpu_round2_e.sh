#!/bin/bash
# GPU session E: whole -m gpu suite on the exact scenario data, bench line, device-timed closed loops (warm_duals 0 / 1).
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/e_pytest.txt
timeout 600 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 100 --batch 8192 > gpurun_out/e_b8192.json 2>> gpurun_out/e_bench.err
timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 50 --batch 32768 > gpurun_out/e_b32768.json 2>> gpurun_out/e_bench.err
timeout 400 python - <<'PY' > gpurun_out/e_closed_loop.txt 2>&1
import numpy as np, torch
import mpc_b200
from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
for name, N, B in (("ZAM_Over-1_1_LF", 30, 1024), ("ZAM_Over-1_1_LF", 10, 1024), ("ZAM_Over-1_1_LF", 30, 8192), ("USA_Lanker-2_18_T-1_LF", 50, 1024), ("USA_Lanker-2_18_T-1_LF", 10, 1024), ("ZAM_Over-1_1_CA", 30, 1024)):
    sc = mpc_b200.load_scenario(name)
    x0 = mpc_b200.make_batch(name, B, N, 7)[1]
    for prec in ("f32", "f64"):
        if prec == "f64" and B > 1024: continue
        for wd in (0, 1):
            opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision=prec, max_batch=B, warm_duals=wd, max_iter=200)
            d_x0 = opt._dev(x0)
            for _ in range(2): opt.optimize_batch(d_x0, return_device=True)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s.record()
            for _ in range(3): tr, ct, st, it = opt.optimize_batch(d_x0, return_device=True)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 3
            st, it = st.cpu().numpy(), it.cpu().numpy(); T = st.shape[1]
            print(f"{name} N={N} B={B} {prec} warm_duals={wd}: {ms:.3f} ms per batch of closed loops, {B*T/ms/1e3:.2f} M MPC-steps/s, status {dict(zip(*[a.tolist() for a in np.unique(st, return_counts=True)]))}, iters step0 {it[:,0].mean():.1f} warm {it[:,1:].mean():.2f} max {it.max()}", flush=True)
PY
cat gpurun_out/e_closed_loop.txt
python - <<'PY'
import json
for f in ("gpurun_out/e_bench.json","gpurun_out/e_b8192.json","gpurun_out/e_b32768.json"):
    try:
        d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e (%.4f ms)"%(d["e2e"]["value"],d["e2e"]["ms_per_step"]), "frac %.3f"%d["roofline"]["frac"], d.get("parity"))
        for k,v in (d.get("extra") or {}).items(): print("   ",k,"%.3e"%v["solves_per_s"],v["converged"],"frac %.3f"%v["roofline_frac_model"],v.get("parity"))
    except Exception as e: print(f,"ERR",e)
PY
