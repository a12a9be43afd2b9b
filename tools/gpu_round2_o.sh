#!/bin/bash
# GPU session O (1 GPU): FORCESPRO-formulation solver -- GPU parity tests, throughput, sanitizer on a small solve.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 1200 python -m pytest tests/test_forces_solver.py tests/test_forces_model.py -x -q -m gpu > gpurun_out/o_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/o_pytest.txt
timeout 600 python tools/bench_forces.py > gpurun_out/o_bench_forces.jsonl 2> gpurun_out/o_bench_forces.err; echo "bench rc=$?"; cat gpurun_out/o_bench_forces.jsonl; tail -n 5 gpurun_out/o_bench_forces.err
