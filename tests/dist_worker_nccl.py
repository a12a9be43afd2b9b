"""Worker of tests/test_sharded_nccl.py: run under `python -m torch.distributed.run --nproc-per-node 2` on a box with >= 2 GPUs.
Rank 0 owns a global batch; `sharding.solve_sharded_nccl` scatters the parameter blocks over NCCL, every rank solves its
shard with the real CUDA solver, the solutions come back over NCCL; rank 0 compares with its own single-GPU solve bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_b200  # noqa: E402
from mpc_b200 import sharding  # noqa: E402
from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for name, N, B in (("ZAM_Over-1_1_LF", 30, 1001), ("ZAM_Over-1_1_CA", 30, 300), ("USA_Lanker-2_18_T-1_LF", 50, 257)):
        sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, B, N, 4242)
        opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, max_batch=B, device=local, max_iter=300)
        g = torch.as_tensor(xref, device=opt.device) if rank == 0 else None
        for algo in ("collective", "p2p"):
            res = sharding.solve_sharded_nccl(opt, g, B, N, src=0, algo=algo)
            torch.cuda.synchronize()
            if algo == "collective":                       # results on every rank: they must all agree with rank 0's
                chk = res[0].sum() + res[1].sum()
                ref = chk.clone(); dist.broadcast(ref, src=0)
                ok = ok and bool(torch.equal(chk, ref))
            if rank == 0:
                U, X, st, it = res
                U1, X1, st1, it1 = opt.solve_batch(g)
                same = bool(torch.equal(U, U1) and torch.equal(X, X1) and torch.equal(st, st1) and torch.equal(it, it1))
                print(f"{name} N={N} B={B} world={world} {algo}: gathered == single-GPU solve bitwise: {same}; converged {(st == 1).sum().item()}/{B}", flush=True)
                ok = ok and same and bool((st == 1).all().item())
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_NCCL_OK" if ok else "SHARDED_NCCL_FAIL", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
