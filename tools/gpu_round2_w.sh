#!/bin/bash
# GPU session W (2 GPUs): the final build on the multi-GPU path, as the driver's scaling run launches it -- the NCCL sharded test
# (world 2, real solver, gathered == single-GPU solve bit for bit) and bench.py --gpus 2 under torch.distributed.run.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 150 python -m pytest tests/test_sharded_nccl.py -m gpu -x -q > gpurun_out/w_pytest_nccl.txt 2>&1; echo "nccl test rc=$?"; tail -n 2 gpurun_out/w_pytest_nccl.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/w_bench_2gpu.json 2> gpurun_out/w_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/w_bench_2gpu.json")); print("value %.4e ms %.4f e2e %.4e"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), json.dumps(d.get("sharded"))[:600])
PY
