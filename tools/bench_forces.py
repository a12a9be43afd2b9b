"""Device-timed throughput of `mpcb200_forces_solve` (FORCESPRO formulation) on synthetic batches: CUDA events on the launching
stream, L2 flushed between timed launches.  Prints one JSON line per workload."""
import json
import sys
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpc_b200  # noqa: E402
from test_forces_solver import _problem, _gpu_opt  # noqa: E402


def run(name, N, B, precision="f32", steps=20, warmup=3, **kw):
    sc, d0, P, x0 = _problem(name, N, seed=20261022, B=B)
    opt = _gpu_opt(sc, N, precision, max_batch=B, **kw)
    xd = opt._dev(x0)
    pd = opt._dev(P).unsqueeze(0).expand(B, N, 10).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=opt.device)
    for _ in range(warmup):
        Z, st, it = opt.forces_solve_batch(xd, pd)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        Z, st, it = opt.forces_solve_batch(xd, pd)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    st, it = st.cpu().numpy(), it.cpu().numpy()
    print(json.dumps(dict(workload=f"{name} FORCESPRO formulation N={N} B={B} {precision}", ms_per_launch=round(ms, 4),
                          solves_per_s=round(B / ms * 1e3, 1), status_counts={int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                          mean_iters=round(float(it.mean()), 2), max_iters=int(it.max()), opts=kw)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "profile":          # the launches ncu captures: lane following, N = 30
        run("ZAM_Over-1_1_LF", 30, int(sys.argv[2]) if len(sys.argv) > 2 else 8192, steps=2, warmup=3, refine_f64=0)
        sys.exit(0)
    run("ZAM_Over-1_1_LF", 30, 1024)
    run("ZAM_Over-1_1_LF", 30, 8192)
    run("ZAM_Over-1_1_LF", 30, 8192, warps_per_cta=1)
    run("ZAM_Over-1_1_CA", 30, 4096)
    run("USA_Lanker-2_18_T-1_LF", 50, 8192)
    run("ZAM_Over-1_1_LF", 30, 1024, precision="f64")
