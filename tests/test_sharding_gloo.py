"""N>1 path on CPU: world_size-2 gloo run of the shard / solve / all_gather plumbing (mpc_b200.sharding).  The local
'solver' is a deterministic stand-in so the test checks partitioning and gathering, not numerics."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from mpc_b200.sharding import shard_range, solve_sharded, broadcast_scenario
    dist.init_process_group("gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    B, N = 37, 6
    g = torch.Generator().manual_seed(5)
    xref = torch.randn(B, N + 1, 5, generator=g, dtype=torch.float64)
    X0 = torch.randn(B, N + 1, 5, generator=g, dtype=torch.float64)
    U0 = torch.randn(B, N, 2, generator=g, dtype=torch.float64)
    seen = []
    def fake_solve(xr, X, U):
        seen.append(xr.shape[0])
        return U + xr[:, :N, :2], X * 2 + xr, torch.full((xr.shape[0],), rank, dtype=torch.int32), torch.arange(xr.shape[0], dtype=torch.int32)
    U, X, status, iters = solve_sharded(fake_solve, xref, X0, U0)
    lo, hi = shard_range(B, rank, world)
    assert seen == [hi - lo]
    assert torch.equal(U, U0 + xref[:, :N, :2]) and torch.equal(X, X0 * 2 + xref)
    exp_status = torch.cat([torch.full((shard_range(B, r, world)[1] - shard_range(B, r, world)[0],), r, dtype=torch.int32) for r in range(world)])
    assert torch.equal(status, exp_status)
    arrs = {"path": np.arange(10.0).reshape(5, 2) * (1 if rank == 0 else -1), "w": np.ones(3) * (rank + 1)}
    out = broadcast_scenario(arrs)
    assert np.array_equal(out["path"], np.arange(10.0).reshape(5, 2)) and np.array_equal(out["w"], np.ones(3))
    # the hardware data path (solve_sharded_nccl, algo="collective": broadcast + all-gather) with a stand-in optimizer on CPU tensors
    from mpc_b200.sharding import solve_sharded_nccl
    class FakeOpt:
        device = torch.device("cpu")
        def solve_batch(self, xr, out=None):
            U, X, st, it = out
            U.copy_(xr[:, :N, :2] * 3); X.copy_(xr + 1); st.fill_(1); it.copy_(torch.arange(xr.shape[0], dtype=torch.int32) + 100 * rank)
            return out
    for Bt in (37, 40):
        xg = torch.randn(Bt, N + 1, 5, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
        U, X, st, it = solve_sharded_nccl(FakeOpt(), xg if rank == 0 else None, Bt, N, src=0, algo="collective")
        assert U.shape == (Bt, N, 2) and torch.equal(U, xg[:, :N, :2] * 3) and torch.equal(X, xg + 1) and bool((st == 1).all())
        exp_it = torch.cat([torch.arange(shard_range(Bt, r, world)[1] - shard_range(Bt, r, world)[0], dtype=torch.int32) + 100 * r for r in range(world)])
        assert torch.equal(it, exp_it)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo_shard_solve_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
