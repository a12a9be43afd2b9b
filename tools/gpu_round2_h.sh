#!/bin/bash
# GPU session H (8 GPUs): the collective-based sharded mode at 2 / 4 / 8 GPUs, world-2 NCCL test (both algorithms), configs 4 / 5 sharded.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
timeout 300 python -m pytest tests/test_sharded_nccl.py -m gpu -x -q > gpurun_out/h_pytest_nccl.txt 2>&1; echo "nccl test rc=$?"; tail -n 3 gpurun_out/h_pytest_nccl.txt
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/h_bench_${n}gpu.json 2> gpurun_out/h_bench_${n}gpu.err; echo "bench $n rc=$?"
done
MPCB200_CFG_STEPS=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/run_configs.py > gpurun_out/h_configs_8gpu.jsonl 2> gpurun_out/h_configs_8gpu.err; echo "configs8 rc=$?"
MPCB200_CFG_STEPS=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/run_configs.py > gpurun_out/h_configs_4gpu.jsonl 2> gpurun_out/h_configs_4gpu.err; echo "configs4 rc=$?"
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.load(open(f"gpurun_out/h_bench_{n}gpu.json")); s=d.get("sharded",{})
        print(n,"GPUs: value %.3e ms %.4f e2e %.3e | sharded %.3e ms %.4f coll_us %.1f bitwise %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],s.get("value",0),s.get("ms_per_step",0),s.get("collective_us_per_step",0),s.get("gathered_equals_single_gpu_solve_bitwise")))
    except Exception as e: print(n,"ERR",e)
for f in ("gpurun_out/h_configs_8gpu.jsonl","gpurun_out/h_configs_4gpu.jsonl"):
    print("==",f)
    for l in open(f):
        try:
            d=json.loads(l)
            if d.get("config")==5: print(" config 5: %.3e solves/s"%d["solves_per_s"], [(r["scenario"],r["converged"],"%.2e"%r["solves_per_s"],"sharded %.2e"%r["sharded_nccl"]["solves_per_s"]) for r in d["per_scenario"]])
            elif d.get("config") in (2,3,4): print(" config",d.get("config"),"%.3e"%d["solves_per_s"],d["converged"],"sharded %.3e (%.3f ms)"%(d["sharded_nccl"]["solves_per_s"],d["sharded_nccl"]["ms_per_batch"]))
        except Exception as e: pass
PY
