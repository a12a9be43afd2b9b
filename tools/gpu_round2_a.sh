#!/bin/bash
# GPU session A of round 2: parity tests, the bench line, the 18-warp build variant, casadi install attempt.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
( python -m pip install casadi 2>&1 | tail -n 3; python -m pip download casadi 2>&1 | tail -n 1; ls /opt/wheelhouse | grep -i casadi ) > gpurun_out/a_casadi_install.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.txt
tail -n 15 gpurun_out/a_pytest.txt
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
for b in 1024 8192 32768; do
  timeout 300 python bench.py --no-cpu-baseline --no-extra --batch $b --steps 50 > gpurun_out/a_w16_b$b.json 2>> gpurun_out/a_bench.err
  MPCB200_LIB=$PWD/build_variants/libmpcb200_w18.so timeout 300 python bench.py --no-cpu-baseline --no-extra --batch $b --steps 50 > gpurun_out/a_w18_b$b.json 2>> gpurun_out/a_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/a_*.json")):
    try:
        d=json.load(open(f))
        print(f, "value %.3e ms %.4f e2e %.3e frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["frac"]), d["config"]["converged"], d["config"]["mean_sqp_iters"])
        if d.get("extra"):
            for k,v in d["extra"].items(): print("   ",k, "%.3e"%v["solves_per_s"], v["converged"], v["mean_sqp_iters"], v.get("parity"))
        print("    parity", d.get("parity"))
    except Exception as e: print(f, "ERR", e)
PY
