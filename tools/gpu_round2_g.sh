#!/bin/bash
# GPU session G (8 GPUs): scaling lines of bench.py at 2 / 4 / 8 GPUs (incl. the NCCL sharded mode), BASELINE configs 2-5 sharded
# over 8 GPUs and config 4 over 4 GPUs (tools/run_configs.py), the world-2 NCCL test.
mkdir -p gpurun_out
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/g_bench_${n}gpu.json 2> gpurun_out/g_bench_${n}gpu.err; echo "bench $n rc=$?"
done
MPCB200_CFG_STEPS=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_configs.py > gpurun_out/g_configs_8gpu.jsonl 2> gpurun_out/g_configs_8gpu.err; echo "configs8 rc=$?"
MPCB200_CFG_STEPS=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 tools/run_configs.py > gpurun_out/g_configs_4gpu.jsonl 2> gpurun_out/g_configs_4gpu.err; echo "configs4 rc=$?"
timeout 300 python -m pytest tests/test_sharded_nccl.py -m gpu -x -q > gpurun_out/g_pytest_nccl.txt 2>&1; echo "nccl test rc=$?"
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.load(open(f"gpurun_out/g_bench_{n}gpu.json")); s=d.get("sharded",{})
        print(n,"GPUs: value %.3e ms %.4f e2e %.3e (%.4f ms) per-rank %s | sharded %.3e ms %.4f coll_us %.1f bitwise %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],[round(x,4) for x in d["e2e"].get("per_rank_ms_per_step",[])],s.get("value",0),s.get("ms_per_step",0),s.get("collective_us_per_step",0),s.get("gathered_equals_single_gpu_solve_bitwise")))
    except Exception as e: print(n,"ERR",e)
for f in ("gpurun_out/g_configs_8gpu.jsonl","gpurun_out/g_configs_4gpu.jsonl"):
    print("==",f)
    for l in open(f):
        try:
            d=json.loads(l)
            if d.get("config")==5: print(" config 5: %.3e solves/s"%d["solves_per_s"], [(r["scenario"],r["converged"],"%.2e"%r["solves_per_s"],"sharded %.2e"%r["sharded_nccl"]["solves_per_s"]) for r in d["per_scenario"]])
            else: print(" config",d.get("config"),{k:d[k] for k in d if k in ("solves_per_s","converged","mean_sqp_iters","mpc_steps_per_s","max_abs_err_traj_vs_oracle","parity_sample","sharded_nccl")})
        except Exception as e: print("ERR",e,l[:100])
PY
tail -n 3 gpurun_out/g_configs_8gpu.err
