#!/bin/bash
# GPU session N (1 GPU): final check of the whole -m gpu suite + smoke + the default bench line, as the driver runs them.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/n_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/n_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/n_smoke.txt
timeout 300 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench.err; echo "ref rc=$?"
timeout 600 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/n_bench.json 2>> gpurun_out/n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n_bench.json","gpurun_out/n_bench_ref.json"):
    d=json.load(open(f)); print(f, "value %.4e ms %.4f"%(d["value"],d["ms_per_step"]), "e2e %.4e"%d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("issue_frac"), (d.get("extra") or {}).get("config4_lanker_n50",{}).get("parity"))
PY
