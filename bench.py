#!/usr/bin/env python3
"""bench.py -- MPC solves/sec on BASELINE.json's configs[1] (ZAM_Over-1_1 lane following, batch 1024 perturbed x0, N=30).

One "step" = one pass of the hot path over one batch: `mpcb200_solve` (ONE fused kernel launch that runs every SQP
iteration of its 1024 NLPs) on inputs already resident in HBM.  `value` = solves/s over all ranks (weak scaling: every
rank owns its own copy of the seeded 1024-instance batch, no data-path collective).  `e2e` = the same metric through the public host-buffer call
`B200Optimizer.solve_batch_host` (pinned host arrays, H2D + solve + D2H inside the timed region).

  python bench.py [--gpus N --steps K --warmup W]          product arm (N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference [...]                   reference arm: the CPU restatement of the reference path (oracle/)
                                                           on all host cores, bounded sample per step
"""
import argparse
import json
import os

# the CPU legs run one single-threaded oracle process per core; keep BLAS/OpenMP from oversubscribing them
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENARIO, N_HORIZON, BATCH, SEED = "ZAM_Over-1_1_LF", 30, int(os.environ.get("MPCB200_BENCH_BATCH", "1024")), 20261017
METRIC = "MPC solves/sec (N=30, 5-state kinematic bicycle) at batch 1024"


def bytes_iter(N):
    """Algorithmic bytes per problem per SQP iteration (SURVEY.md 8d / DESIGN.md): 4*(219N + 65)."""
    return 4 * (219 * N + 65)


# ------------------------------------------------------------------------------------------------ CPU (oracle) leg
def _oracle_one(args):
    name, N, xref, X, U = args
    import mpc_b200
    from oracle import nlp, ipm
    sc = mpc_b200.load_scenario(name)
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref, sc.static_obstacle)
    r = ipm.solve(d, nlp.pack(U, X))
    return r["status"], r["iters"], r["w"]


def cpu_oracle_rate(n_sample, cores, pool=None):
    """Times the float64 oracle on the first n_sample instances of the workload over `cores` processes."""
    import multiprocessing as mp
    import mpc_b200
    sc, x0, xref, X, U = mpc_b200.make_batch(SCENARIO, n_sample, N_HORIZON, SEED)
    jobs = [(SCENARIO, N_HORIZON, xref[b], X[b], U[b]) for b in range(n_sample)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(_oracle_one, jobs[:cores])          # warm the workers (imports)
    t0 = time.perf_counter()
    res = pool.map(_oracle_one, jobs, chunksize=max(1, n_sample // (4 * cores)))
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    ok = sum(1 for r in res if r[0] == 1)
    return n_sample / dt, dt, ok, res


def run_reference(args):
    """--impl reference: the reference's CPU path (restated: oracle/), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_sample = max(cores, min(BATCH, 16 * cores))          # bounded sample per step (~0.5 s on all cores)
    import mpc_b200
    sc, x0, xref, X, U = mpc_b200.make_batch(SCENARIO, n_sample, N_HORIZON, SEED)
    jobs = [(SCENARIO, N_HORIZON, xref[b], X[b], U[b]) for b in range(n_sample)]
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_oracle_one, jobs[:cores])
    for _ in range(args.warmup):
        pool.map(_oracle_one, jobs)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        pool.map(_oracle_one, jobs)
        times.append(time.perf_counter() - t0)
    pool.close()
    total = sum(times)
    value = n_sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{SCENARIO} batch={BATCH} perturbed x0 N={N_HORIZON} (BASELINE configs[1])",
                   "sample": f"first {n_sample} of the {BATCH} instances per step"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} instances per step x {args.steps} steps, float64 oracle (oracle/ipm.py), "
                                   f"multiprocessing.Pool({cores})"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  The region lasts only tens of milliseconds, far below the
    start-up time of an `nvidia-smi -lms` child, so the samples come from NVML in-process (pynvml, a thread polling every
    ~2 ms; NVML queries are host-side and do not enter the CUDA stream).  Falls back to one `nvidia-smi` query."""
    BAD = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.sm, self.mask, self.stop_flag, self.t, self.h, self.nv = index, [], 0, False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it holds plain ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            reasons = sorted(n for n, bit in self.BAD if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(self.sm), "source": "NVML in-process, 2 ms poll over the timed regions"}
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            reasons = sorted(n for (n, _), v in zip(self.BAD, f[2:6]) if v.lower().startswith("active"))
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi, one query right after the timed regions"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------ product arm
def run_product(args):
    import torch
    import mpc_b200
    from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # stdout carries exactly ONE line (the JSON): NCCL's debug output (the "NCCL version ..." banner of NCCL_DEBUG=VERSION,
        # which the box sets) goes to stderr instead.  NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION
        # becomes WARN (same banner, plus warnings if any).
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
        local = 0
    dev = torch.device("cuda", local)
    N, B = N_HORIZON, BATCH
    # weak scaling: every rank owns its own copy of the seeded config-2 batch (seed 20261017), so the per-GPU work
    # is EXACTLY the same at every N (with per-rank seeds the slowest draw -- the instance with the most SQP iterations --
    # would set the max-over-ranks time and read as a scaling loss); no data-path collective
    sc, x0, xref, X0, U0 = mpc_b200.make_batch(SCENARIO, B, N, SEED)
    solver_opts = {}
    for kv in args.opt:                                   # tuning knob: fields of mpcb200_config, e.g. --opt mu_min=1e-6
        k, v = kv.split("=")
        solver_opts[k] = int(v) if k in ("max_iter", "ls_max", "acc_iters", "stall_iters", "refine_f64", "init_rollout") else float(v)
    opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision=args.precision,
                        hessian=args.hessian, max_batch=B, device=local, **solver_opts)
    f64 = torch.float64
    d_xref = torch.as_tensor(xref, device=dev)
    assert np.array_equal(X0, np.repeat(xref[:, :1], N + 1, axis=1)) and not U0.any()      # the workload IS the cold start
    d_X = torch.empty(B, N + 1, 5, dtype=f64, device=dev); d_U = torch.empty(B, N, 2, dtype=f64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    d_iters = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)      # 256 MB > 126 MB L2
    h, stream = opt.handle, torch.cuda.current_stream(dev)

    def step():
        # cold start = the reference's step-0 initial guess (X_0 tiled, zero controls): X / U are outputs only
        h.check(h.lib.mpcb200_solve_cold(h.h, d_xref.data_ptr(), d_X.data_ptr(), d_U.data_ptr(), d_status.data_ptr(),
                                         d_iters.data_ptr(), B, stream.cuda_stream))

    def reset():
        flush.fill_(1.0)                     # evict L2 between timed iterations

    for _ in range(max(args.warmup, 3)):
        reset(); step()
    torch.cuda.synchronize(dev)
    status = d_status.cpu().numpy(); iters = d_iters.cpu().numpy()
    n_ok = int((status == 1).sum())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    n0 = h.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s_, e_ in ev:
        reset()
        s_.record(stream); step(); e_.record(stream)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    launches = h.launch_count - n0
    ms = np.array([s_.elapsed_time(e_) for s_, e_ in ev])
    total_ms = float(ms.sum())
    if dist:
        t = torch.tensor([total_ms], device=dev, dtype=f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    # ---- end-to-end through the public host-buffer API (pinned host memory, H2D + solve + D2H per step)
    hx = torch.as_tensor(xref).pin_memory()
    hX = torch.empty(B, N + 1, 5, dtype=f64).pin_memory(); hU = torch.empty(B, N, 2, dtype=f64).pin_memory()   # pinned result buffers
    for _ in range(3):
        opt.solve_batch_host(hx.numpy(), out=(hX.numpy(), hU.numpy()))
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    n1 = h.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Ue, Xe, ste, ite = opt.solve_batch_host(hx.numpy(), out=(hX.numpy(), hU.numpy()))
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    assert (ste == 1).all() and float(Ue[0, 0, 1]) == float(Ue[0, 0, 1])               # the result was read back
    launches += h.launch_count - n1
    if dist:
        t = torch.tensor([e2e_s], device=dev, dtype=f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant (only) kernel: algorithmic bytes of one launch / its mean duration
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs, burst)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    alg_bytes = float(iters.sum()) * bytes_iter(N)
    launch_ms = total_ms / args.steps
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 only at N=1
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = BATCH                                   # the whole workload: ~25 ms per solve per core, 10-40 core-seconds
        rate, dt, ok, ores = cpu_oracle_rate(n_sample, cores)
        cpu = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"first {n_sample} of the {BATCH} instances, float64 oracle (oracle/ipm.py), Pool({cores}), {dt:.1f} s, {ok} converged"}
        # the oracle solutions of that leg double as the checker of the timed GPU results (same instances, same cold start)
        from oracle import nlp as _nlp
        Xg, Ug = d_X.cpu().numpy(), d_U.cpu().numpy()
        dU = dX = 0.0
        n_cmp = 0
        for b, (st_o, _, w_o) in enumerate(ores):
            if st_o == 1 and status[b] == 1:
                Uo, Xo = _nlp.split(w_o, N)
                dU = max(dU, float(np.abs(Uo - Ug[b]).max())); dX = max(dX, float(np.abs(Xo - Xg[b]).max()))
                n_cmp += 1
        parity = {"checked": n_cmp, "of": B, "max_abs_dU": dU, "max_abs_dX": dX, "tolerance": 1e-3,
                  "against": "float64 oracle (restated reference NLP), every instance of the timed batch"}
    nx, nu = 5 * (N + 1), 2 * N
    line = {
        "metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "p50_ms_per_step": float(np.median(ms)), "p50_ms_per_solve": float(np.median(ms)) / B,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{SCENARIO} batch={B} perturbed x0 (seed {SEED}) N={N} cold start (BASELINE configs[1]); every rank solves its own copy",
                   "parallelism": f"batch-shard x{world} (independent NLPs, no data-path collective)",
                   "l2": "256 MB flush write between timed iterations", "hessian": args.hessian, **({"solver_opts": solver_opts} if solver_opts else {}),
                   "converged": f"{n_ok}/{B}", "mean_sqp_iters": float(iters.mean()), "max_sqp_iters": int(iters.max())},
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": B * nx * 8,
                "d2h_bytes_per_step": B * (nx + nu) * 8 + B * 8, "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "mpc_warp_solve_kernel",
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "algorithmic bytes = sum over problems of SQP iterations x 4(219N+65) B (KKT slab staged once per iteration); "
                             "the fused kernel keeps the slab in shared memory, so real DRAM traffic is far below this"},
        "cpu_baseline": cpu,
        "parity": parity,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--hessian", default="gn", choices=["exact", "gn"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="solver option override key=value (experiments; default: none)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
