#!/bin/bash
# GPU session R (1 GPU): the phase-aligned kernel -- bitwise test, then device-timed throughput against the independent-warp kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k phase_aligned 2>&1 | grep -E "^E|passed|failed" | head
for b in 1024 8192 32768; do
  for w in 0 8 16; do
    timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 40 --warmup 5 --batch $b --opt warps_per_cta=$w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $b wpc $w: %.4f ms  %.3f M solves/s  e2e %.3f M/s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))"
  done
done
