#!/usr/bin/env python3
"""Experiment: the end-to-end host path as ONE launch whose TMA bulk copies read / write pinned host memory directly
(zero copy over PCIe), against the chunked H2D + solve + D2H pipeline of mpcb200_solve_host.  Also prints the SQP
iteration histogram of the bench workload.  GPU box only."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_b200  # noqa: E402
from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state  # noqa: E402

B, N = int(os.environ.get("EXP_B", "1024")), 30
sc, x0, xref, X0, U0 = mpc_b200.make_batch("ZAM_Over-1_1_LF", B, N, 20261017)
opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, max_batch=B, device=0)
h = opt.handle
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
f64 = torch.float64
hx = torch.as_tensor(xref).pin_memory()
hX = torch.empty(B, N + 1, 5, dtype=f64).pin_memory()
hU = torch.empty(B, N, 2, dtype=f64).pin_memory()
hst = torch.empty(B, dtype=torch.int32).pin_memory()
hit = torch.empty(B, dtype=torch.int32).pin_memory()
d_xref = torch.as_tensor(xref, device=dev)
d_X = torch.empty(B, N + 1, 5, dtype=f64, device=dev)
d_U = torch.empty(B, N, 2, dtype=f64, device=dev)
d_st = torch.empty(B, dtype=torch.int32, device=dev)
d_it = torch.empty(B, dtype=torch.int32, device=dev)


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def host_path():
    opt.solve_batch_host(hx.numpy(), out=(hX.numpy(), hU.numpy()))


def dev_only():
    h.check(h.lib.mpcb200_solve_cold(h.h, d_xref.data_ptr(), d_X.data_ptr(), d_U.data_ptr(), d_st.data_ptr(), d_it.data_ptr(), B,
                                     stream.cuda_stream))
    torch.cuda.synchronize()


def zero_copy():
    h.check(h.lib.mpcb200_solve_cold(h.h, hx.data_ptr(), hX.data_ptr(), hU.data_ptr(), hst.data_ptr(), hit.data_ptr(), B,
                                     stream.cuda_stream))
    torch.cuda.synchronize()


def zero_copy_in():
    # inputs read from host memory by the kernel, outputs to HBM then copied
    h.check(h.lib.mpcb200_solve_cold(h.h, hx.data_ptr(), d_X.data_ptr(), d_U.data_ptr(), d_st.data_ptr(), d_it.data_ptr(), B,
                                     stream.cuda_stream))
    hX.copy_(d_X, non_blocking=True); hU.copy_(d_U, non_blocking=True); hst.copy_(d_st, non_blocking=True); hit.copy_(d_it, non_blocking=True)
    torch.cuda.synchronize()


def zero_copy_out():
    d_xref.copy_(hx, non_blocking=True)
    h.check(h.lib.mpcb200_solve_cold(h.h, d_xref.data_ptr(), hX.data_ptr(), hU.data_ptr(), hst.data_ptr(), hit.data_ptr(), B,
                                     stream.cuda_stream))
    torch.cuda.synchronize()


print("B", B)
print("device-only + sync      ms", timeit(dev_only))
it = d_it.cpu().numpy()
print("iters histogram", np.bincount(it), "mean", it.mean())
Xd, Ud = d_X.cpu().numpy(), d_U.cpu().numpy()
print("solve_host (chunked)    ms", timeit(host_path))
for name, fn in (("zero-copy in+out", zero_copy), ("zero-copy in only", zero_copy_in), ("zero-copy out only", zero_copy_out)):
    try:
        print(f"{name:22s}  ms", timeit(fn))
    except Exception as e:  # noqa: BLE001
        print(name, "FAILED", e)
zero_copy()
print("zero-copy results identical:", np.array_equal(hX.numpy(), Xd), np.array_equal(hU.numpy(), Ud),
      np.array_equal(hit.numpy(), it), int((hst.numpy() == 1).sum()))
for nch in (1, 2, 4):
    os.environ["MPCB200_HOST_CHUNKS"] = str(nch)
    print("solve_host chunks", nch, "ms", timeit(host_path))
