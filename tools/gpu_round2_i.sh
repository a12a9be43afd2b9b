#!/bin/bash
# GPU session I (1 GPU): whole -m gpu suite (scenario ids, dual ABI, new tests), run_configs on 1 GPU (config 5 one-launch), smoke.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/i_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/i_smoke.txt
MPCB200_CFG_STEPS=10 timeout 900 python tools/run_configs.py > gpurun_out/i_configs_1gpu.jsonl 2> gpurun_out/i_configs_1gpu.err; echo "configs rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/i_configs_1gpu.jsonl"):
    try:
        d=json.loads(l)
        if d.get("config")==5: print(" config 5: sequential %.3e solves/s; one launch mixed"%d["solves_per_s"], d["one_launch_mixed"], [(r["scenario"],r["converged"],"%.2e"%r["solves_per_s"]) for r in d["per_scenario"]])
        else: print(" config",d.get("config"),{k:d[k] for k in d if k in ("solves_per_s","converged","mean_sqp_iters","mpc_steps_per_s","max_abs_err_traj_vs_oracle","max_abs_err_ctrl_vs_oracle","parity_sample","converged_steps","mean_sqp_iters_warm_steps","wall_s")})
    except Exception as e: print("ERR",e,l[:80])
PY
tail -n 3 gpurun_out/i_configs_1gpu.err
