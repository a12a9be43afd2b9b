#!/bin/bash
# GPU session S (1 GPU): final-phase extrapolation + rate-based exit in the CasADi-formulation core: throughput at three batches, then the whole GPU suite.
mkdir -p gpurun_out
for b in 1024 8192 32768; do
  timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 60 --warmup 5 --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $b: %.4f ms  %.3f M solves/s  e2e %.3f M/s  iters mean %.3f max %d conv %s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['config']['mean_sqp_iters'], d['config']['max_sqp_iters'], d['config']['converged']))"
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/s_pytest.txt
