#!/bin/bash
# GPU session V (1 GPU): compute-sanitizer over the kernels added this round (FORCESPRO-formulation solve / closed loop / road-boundary
# variant, phase-aligned solve kernel): memcheck and racecheck.
mkdir -p gpurun_out
OUT=gpurun_out/v_sanitizer.txt
: > $OUT
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool : tests/test_forces_solver.py -k 'warm_start_misaligned or road_boundary or closed_loop_on_device'" >> $OUT
  timeout 500 compute-sanitizer --tool $tool python -m pytest tests/test_forces_solver.py -x -q -m gpu -k "warm_start_misaligned or road_boundary or closed_loop_on_device" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|Invalid|Error" | tail -6 >> $OUT
  echo "=== compute-sanitizer --tool $tool : tests/test_gpu_parity.py -k 'phase_aligned'" >> $OUT
  timeout 500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "phase_aligned" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|Invalid|Error" | tail -6 >> $OUT
done
cat $OUT
