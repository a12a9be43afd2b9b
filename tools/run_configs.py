#!/usr/bin/env python3
"""Run the five BASELINE.json configs on the GPUs of this box and print one JSON line per config.

  python tools/run_configs.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
         tools/run_configs.py                                   # G GPUs: every config's batch is sharded contiguously

Per config: converged fraction, SQP iteration statistics, device-timed solves/s over all ranks (CUDA events, max over
ranks), and a parity sample against the float64 oracle (checker only; oracle/ is never on the measured path).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SCEN6 = ["ZAM_Over-1_1_CA", "ZAM_Over-1_1_LFfile", "USA_Lanker-2_18_T-1_LF", "USA_Peach-2_1_T-1", "ZAM_Tutorial-1_2_T-1",
         "ZAM_Tutorial_Urban-3_2"]


def main():
    import torch
    import mpc_b200
    from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
    from mpc_b200.sharding import shard_range
    from oracle import nlp, ipm, closed_loop
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    steps = int(os.environ.get("MPCB200_CFG_STEPS", "10"))

    def timed_solve(name, B, N, seed, precision="f32", max_iter=200, n_check=4, **opts):
        sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, B, N, seed)
        lo, hi = shard_range(B, rank, world)
        opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision=precision,
                            max_batch=max(hi - lo, 1), device=local, max_iter=max_iter, **opts)
        d_xref, d_X0, d_U0 = opt._dev(xref[lo:hi]), opt._dev(X0[lo:hi]), opt._dev(U0[lo:hi])
        stream = torch.cuda.current_stream(dev)
        for _ in range(2):
            U, X, st, it = opt.solve_batch(d_xref, d_X0, d_U0)
        torch.cuda.synchronize(dev)
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            U, X, st, it = opt.solve_batch(d_xref, d_X0, d_U0)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        st_n, it_n = st.cpu().numpy(), it.cpu().numpy()
        stats = torch.tensor([ms, float((st_n == 1).sum()), float(it_n.sum()), float(it_n.max()), float(hi - lo)], device=dev, dtype=torch.float64)
        if dist:
            mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, n_ok, it_sum, it_max = float(mx[0]), float(sm[1]), float(sm[2]), float(mx[3])
        else:
            ms, n_ok, it_sum, it_max = [float(v) for v in stats[:4]]
        # the same batch through the NCCL data path: rank 0 owns it, shards go out and solutions come back inside the timed region
        sharded = None
        if dist:
            from mpc_b200 import sharding
            g_xref = opt._dev(xref) if rank == 0 else None
            for _ in range(2):
                sharding.solve_sharded_nccl(opt, g_xref, B, N, src=0)
            torch.cuda.synchronize(dev); dist.barrier()
            e0.record(stream)
            for _ in range(steps):
                res = sharding.solve_sharded_nccl(opt, g_xref, B, N, src=0)
            e1.record(stream); torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nxb, nub = 5 * (N + 1) * 8, 2 * N * 8
            sharded = dict(ms_per_batch=float(t), solves_per_s=B / (float(t) * 1e-3), algo="collective (broadcast + all-gather)",
                           scatter_bytes=(B - (hi - lo)) * nxb if rank == 0 else None,
                           gather_bytes=(B - (hi - lo)) * (nxb + nub + 8) if rank == 0 else None)
            if rank == 0:
                sharded["all_converged"] = bool((res[2] == 1).all().item())
        # parity sample on rank 0's shard (oracle = checker)
        worst, n_cmp, kkt = 0.0, 0, []
        if rank == 0:
            Un, Xn = U.cpu().numpy(), X.cpu().numpy()
            for b in range(min(n_check, hi - lo)):
                if st_n[b] != 1:
                    continue
                d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref[b], sc.static_obstacle)
                if sc.use_case == "collision_avoidance":
                    r = ipm.solve(d, nlp.pack(Un[b], Xn[b]))          # multi-modal: the oracle warm-started at our point must stay
                else:
                    r = ipm.solve(d, nlp.pack(U0[b], X0[b]))
                if r["status"] != 1:
                    continue
                Uo, Xo = nlp.split(r["w"], N)
                worst = max(worst, float(np.abs(Un[b] - Uo).max()), float(np.abs(Xn[b] - Xo).max())); n_cmp += 1
        return dict(scenario=name, B=B, N=N, gpus=world, precision=precision, ms_per_batch=ms, solves_per_s=B / (ms * 1e-3),
                    converged=f"{int(n_ok)}/{B}", mean_sqp_iters=it_sum / B, max_sqp_iters=int(it_max),
                    parity_sample=dict(n=n_cmp, max_abs_err_vs_oracle=worst), **({"sharded_nccl": sharded} if sharded else {}))

    out = []
    # ---- config 1: reference plumbing, single ego, closed loop T = 30, N = 30 (rank 0 only)
    if rank == 0:
        sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
        N = 30
        opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision="f64", max_batch=8, device=local)
        t0 = time.perf_counter(); ts, us, tv = opt.optimize(); t1 = time.perf_counter()
        traj_o, u_o = closed_loop.optimize(sc, N)
        out.append(dict(config=1, scenario="ZAM_Over-1_1_LF", B=1, N=N, closed_loop_steps=int(ts.shape[0]), wall_s=t1 - t0,
                        max_abs_err_traj_vs_oracle=float(np.abs(ts - traj_o).max()), max_abs_err_ctrl_vs_oracle=float(np.abs(us - u_o).max()),
                        end_speed=float(ts[-1, 3])))
        print(json.dumps(out[-1]), flush=True)
        # ---- config 1b: the same closed loop for 1024 perturbed egos, whole loop on the device (one launch)
        sc, x0b, *_ = mpc_b200.make_batch("ZAM_Over-1_1_LF", 1024, N, 20261017)
        opt32 = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, precision="f32", max_batch=1024, device=local)
        d_x0 = opt32._dev(x0b)
        for _ in range(2):
            tr, ct, stl, itl = opt32.optimize_batch(d_x0, return_device=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        for _ in range(5):
            tr, ct, stl, itl = opt32.optimize_batch(d_x0, return_device=True)
        e1.record(torch.cuda.current_stream(dev)); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 5
        stl, itl = stl.cpu().numpy(), itl.cpu().numpy()
        out.append(dict(config="1b", scenario="ZAM_Over-1_1_LF", B=1024, N=N, closed_loop_steps=int(stl.shape[1]), ms_per_batch=ms,
                        closed_loops_per_s=1024 / (ms * 1e-3), mpc_steps_per_s=1024 * stl.shape[1] / (ms * 1e-3),
                        converged_steps=f"{int(np.isin(stl, (1, 3)).sum())}/{stl.size}", mean_sqp_iters_per_step=float(itl.mean()),
                        mean_sqp_iters_warm_steps=float(itl[:, 1:].mean())))
        print(json.dumps(out[-1]), flush=True)
    if dist:
        dist.barrier()
    for cfg_id, (name, B, N, seed, kw) in {2: ("ZAM_Over-1_1_LF", 1024, 30, 20261017, {}),
                                           3: ("ZAM_Over-1_1_CA", 4096, 30, 20261018, dict(max_iter=300, refine_f64=1)),
                                           4: ("USA_Lanker-2_18_T-1_LF", 8192, 50, 20261019, {})}.items():
        r = timed_solve(name, B, N, seed, **kw); r["config"] = cfg_id
        if rank == 0:
            print(json.dumps(r), flush=True)
    tot_ms, rows = 0.0, []
    for i, name in enumerate(SCEN6):
        r = timed_solve(name, 4096, 30, 20261020 + i, max_iter=300, n_check=2, **(dict(refine_f64=1) if name.endswith("_CA") else {})); rows.append(r); tot_ms += r["ms_per_batch"]
    # the same 6 x 4096 instances as ONE mixed batch per GPU (per-problem scenario ids, one launch + its refinement pass);
    # instances are interleaved across scenarios so that every rank's contiguous shard holds the same mix
    scs, xs = [], []
    for i, name in enumerate(SCEN6):
        sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, 4096, 30, 20261020 + i)
        scs.append(sc); xs.append(xref)
    xref_all = np.stack(xs, axis=1).reshape(6 * 4096, 31, 5)
    sid_all = np.tile(np.arange(6, dtype=np.int32), 4096)
    Bm = 6 * 4096
    lo, hi = shard_range(Bm, rank, world)
    optm = B200Optimizer(make_configuration(scs[1], 30), init_values_from_state(scs[1].x0), 30, precision="f32", max_batch=hi - lo, device=local,
                         max_iter=300, refine_f64=1)
    optm.set_scenarios(scs)
    d_x, d_s = optm._dev(xref_all[lo:hi]), torch.as_tensor(sid_all[lo:hi], device=dev)
    for _ in range(2):
        Um, Xm, stm, itm = optm.solve_batch_scenarios(d_x, d_s)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        Um, Xm, stm, itm = optm.solve_batch_scenarios(d_x, d_s)
    e1.record(stream); torch.cuda.synchronize(dev)
    stats = torch.tensor([e0.elapsed_time(e1) / steps, float((stm == 1).sum().item())], device=dev, dtype=torch.float64)
    if dist:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_mixed, ok_mixed = float(mx[0]), float(sm[1])
    else:
        ms_mixed, ok_mixed = float(stats[0]), float(stats[1])
    if rank == 0:
        print(json.dumps(dict(config=5, gpus=world, total_instances=6 * 4096, ms_total=tot_ms, solves_per_s=6 * 4096 / (tot_ms * 1e-3),
                              one_launch_mixed=dict(ms=ms_mixed, solves_per_s=Bm / (ms_mixed * 1e-3), converged=f"{int(ok_mixed)}/{Bm}",
                                                    how="per-problem scenario ids (mpcb200_solve_scenarios), one float32 launch + its float64 refinement pass per GPU"),
                              per_scenario=rows)), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
