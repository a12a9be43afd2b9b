#!/bin/bash
# GPU session B (2 GPUs): sharded NCCL test, 2-GPU bench line, refine on/off A/B at batch 1024, the re-stated config-3 test.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_nccl.py -m gpu -x -q > gpurun_out/b_pytest_nccl.txt 2>&1; echo "nccl pytest rc=$?"; tail -n 5 gpurun_out/b_pytest_nccl.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or config4 or config5 or misaligned or large_batch" > gpurun_out/b_pytest_new.txt 2>&1; echo "new tests rc=$?"; tail -n 8 gpurun_out/b_pytest_new.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/b_bench_2gpu.json 2> gpurun_out/b_bench_2gpu.err; echo "bench2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --opt refine_f64=0 > gpurun_out/b_norefine_b1024.json 2> gpurun_out/b_err.txt
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 > gpurun_out/b_refine_b1024.json 2>> gpurun_out/b_err.txt
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 200 --opt warps_per_cta=1 --opt refine_f64=0 > gpurun_out/b_wpc1_b1024.json 2>> gpurun_out/b_err.txt
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 50 --batch 8192 --opt warps_per_cta=1 > gpurun_out/b_wpc1_b8192.json 2>> gpurun_out/b_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d=json.load(open(f))
        print(f, "value %.3e ms %.4f p50 %.4f e2e %.3e (%.4f ms) pageable %.4f"%(d["value"],d["ms_per_step"],d["p50_ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e"]["pageable_ms_per_step"]))
        if d.get("sharded"): print("   sharded", d["sharded"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 5 gpurun_out/b_bench_2gpu.err
