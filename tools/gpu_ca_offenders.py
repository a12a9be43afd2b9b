"""Diagnostic (GPU box): which config-3 instances end further than 5e-4 from the oracle's local optimum, and were they refined?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, multiprocessing as mp
import mpc_b200, bench
from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
name, N, B = "ZAM_Over-1_1_CA", 30, 4096
sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, B, N, 20261018)
def run(**kw):
    opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, max_batch=B, max_iter=300, **kw)
    return [t.cpu().numpy() for t in opt.solve_batch(xref)]
U0_, X0_, st0, it0 = run(refine_f64=0)
U, X, st, it = run()
print("f32 only status", dict(zip(*np.unique(st0, return_counts=True))), "default", dict(zip(*np.unique(st, return_counts=True))))
with mp.get_context("fork").Pool(os.cpu_count()) as pool:
    res = pool.map(bench._warm_one, [(name, N, xref[b], X[b], U[b]) for b in range(B)], chunksize=8)
sto = np.array([r[0] for r in res]); dw = np.array([r[1] for r in res])
off = np.where((sto == 1) & (dw > 3e-4))[0]
for b in off:
    print(b, "f32 status", st0[b], "iters f32", it0[b], "total", it[b], "dw %.2e" % dw[b])
print("oracle failures", int((sto != 1).sum()), "max dw", dw[sto == 1].max(), "n>1e-3", int((dw[sto == 1] > 1e-3).sum()), "n>3e-4", len(off))
