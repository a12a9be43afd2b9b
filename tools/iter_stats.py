#!/usr/bin/env python3
"""TEST TOOLING: SQP-iteration statistics of the solver core on the host warp emulator (tests/host_sim), all cores.
  python tools/iter_stats.py [key=value ...]      e.g.  mu_up_alpha=0.3 mu_factor=0.1   (fields of mpcb200_config)
Prints, per workload: converged count, iteration histogram / mean / max, max deviation from the default-config result."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")):
    sys.path.insert(0, p)

WORK = [("ZAM_Over-1_1_LF", 30, 1024, 20261017), ("ZAM_Over-1_1_CA", 30, 256, 20261018), ("USA_Lanker-2_18_T-1_LF", 50, 256, 20261019),
        ("USA_Peach-2_1_T-1", 30, 128, 20261022), ("ZAM_Tutorial-1_2_T-1", 30, 128, 20261023), ("ZAM_Tutorial_Urban-3_2", 30, 128, 20261024)]


def _job(a):
    name, N, lo, hi, seed, B, opts, so = a
    import ctypes
    import hostsim
    if so:
        hostsim._lib = ctypes.CDLL(so)
    import mpc_b200
    from test_host_logic import _cfg
    sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, B, N, seed)
    cfg = _cfg(sc, N, 0, **opts)
    X, U, st, it, _ = hostsim.solve(cfg, xref[lo:hi], X0[lo:hi], U0[lo:hi])
    return X, U, st, it


def run(opts, so=None, work=WORK, pool=None):
    out = {}
    if not so:
        import hostsim
        hostsim.lib()                       # (re)build once before the workers fork
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(os.cpu_count())
    jobs, spans = [], []
    for name, N, B, seed in work:
        step = max(4, B // 32)
        for lo in range(0, B, step):
            jobs.append((name, N, lo, min(B, lo + step), seed, B, opts, so))
            spans.append(name)
    res = pool.map(_job, jobs, chunksize=1)
    if own:
        pool.close()
    for name, N, B, seed in work:
        parts = [r for r, s in zip(res, spans) if s == name]
        out[name] = tuple(np.concatenate([p[i] for p in parts]) for i in range(4))
    return out


def report(out, base=None):
    for name, (X, U, st, it) in out.items():
        dev = ""
        if base is not None:
            Xb, Ub, stb, itb = base[name]
            ok = (st == 1) & (stb == 1)
            dev = f" dev_vs_base X {np.abs(X - Xb)[ok].max():.2e} U {np.abs(U - Ub)[ok].max():.2e}"
        print(f"{name:26s} ok {int((st == 1).sum())}/{len(st)} st3 {int((st == 3).sum())} mean {it.mean():6.2f} p99 {np.percentile(it, 99):5.1f} "
              f"max {it.max():3d} hist {np.bincount(it)[:16].tolist()}{dev}")


if __name__ == "__main__":
    opts, so = {}, None
    for a in sys.argv[1:]:
        k, v = a.split("=")
        if k == "so":
            so = v
        else:
            opts[k] = int(v) if k in ("max_iter", "ls_max", "acc_iters", "stall_iters", "hessian", "init_rollout") else float(v)
    base = None
    cache = "/tmp/exp/base.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        base = {n: tuple(z[f"{n}|{i}"] for i in range(4)) for n, *_ in WORK}
    out = run(opts, so)
    if base is None and not opts and not so:
        np.savez(cache, **{f"{n}|{i}": out[n][i] for n in out for i in range(4)})
    report(out, base)
