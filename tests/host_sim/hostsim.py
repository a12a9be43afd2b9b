"""TEST TOOLING ONLY: ctypes loader for the g++ build of the device solver core (tests/host_sim/host_sim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Config(C.Structure):
    """Mirror of mpcb200_config (include/mpcb200.h)."""
    _fields_ = ([(n, C.c_int32) for n in ("abi_version", "device", "N", "max_batch", "precision", "hessian",
                                          "max_iter", "ls_max")] +
                [("dt", C.c_double), ("l_wb", C.c_double), ("l_fric", C.c_double), ("Q", C.c_double * 5),
                 ("R", C.c_double * 2), ("deltav_min", C.c_double), ("deltav_max", C.c_double), ("a_max", C.c_double),
                 ("delta_min", C.c_double), ("delta_max", C.c_double), ("v_min", C.c_double), ("v_max", C.c_double),
                 ("r_sum", C.c_double), ("ego_offset", C.c_double), ("obstacle", C.c_double * 6),
                 ("mu0", C.c_double), ("mu_min", C.c_double), ("mu_factor", C.c_double), ("tol_step", C.c_double),
                 ("tol_feas", C.c_double), ("tau_min", C.c_double), ("bound_push", C.c_double), ("mu_min_alpha", C.c_double), ("mu_up_alpha", C.c_double), ("mu_up_factor", C.c_double), ("mu_max", C.c_double), ("mu_factor_full", C.c_double),
                 ("kappa_sigma", C.c_double), ("screen_inv_curv", C.c_double), ("trust_step", C.c_double), ("acc_factor", C.c_double),
                 ("acc_iters", C.c_int32), ("stall_iters", C.c_int32), ("refine_f64", C.c_int32), ("init_rollout", C.c_int32),
                 ("mu_warm", C.c_double), ("warm_push", C.c_double), ("kappa_warm", C.c_double), ("stiff_slack", C.c_double),
                 ("warm_duals", C.c_int32), ("warps_per_cta", C.c_int32), ("host_route", C.c_int32), ("host_chunks", C.c_int32)])


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libhostsim.so")
        src = os.path.join(HERE, "host_sim.cpp")
        csrc = os.path.join(HERE, "..", "..", "motion-planning-for-autonomous-driving-with-mpc_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("warp_core.cuh", "loop_core.cuh", "forces_model.cuh", "forces_core.cuh", "warp_ctx.cuh", "mpc_types.cuh", "config_params.h")]
        if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMPC_DIAG", "-shared", "-fPIC", "-o", so, src], cwd=HERE)
        _lib = C.CDLL(so)
    return _lib


def default_config(N, precision=0):
    c = Config()
    lib().hostsim_default_config(C.byref(c), N, precision)
    return c


def solve(cfg, xref, X, U, trace=0):
    xref = np.ascontiguousarray(xref, np.float64)
    X = np.ascontiguousarray(X, np.float64).copy()
    U = np.ascontiguousarray(U, np.float64).copy()
    B = xref.shape[0]
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    kkt = np.zeros(B, np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hostsim_solve(C.byref(cfg), p(xref), p(X), p(U), p(st), p(it), p(kkt), B, trace)
    return X, U, st, it, kkt


def closed_loop(cfg, sc, x0):
    """The device closed-loop body (csrc/loop_core.cuh) on the emulator: x0 [B,5] -> traj [B,T,5], ctrl [B,T,2], status, iters [B,T]."""
    x0 = np.ascontiguousarray(np.atleast_2d(x0), np.float64)
    B, T = x0.shape[0], int(sc.iter_length)
    path = np.ascontiguousarray(np.asarray(sc.reference_path, float)[:, :2])
    orient = np.ascontiguousarray(np.asarray(sc.orientation, float))
    traj = np.zeros((B, T, 5)); ctrl = np.zeros((B, T, 2))
    st = np.zeros((B, T), np.int32); it = np.zeros((B, T), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hostsim_closed_loop.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5 + [C.c_int]
    lib().hostsim_closed_loop(C.byref(cfg), T, p(path), p(orient), float(sc.desired_velocity), p(x0), p(traj), p(ctrl), p(st), p(it), B)
    return traj, ctrl, st, it


FORCES_OUT_WORDS = 136


def unpack_forces(out):
    """[n,136] -> dict(c[n,5], dc[n,5,7], h[n,10], dh[n,10,7], f[n], df[n,7], fN[n], dfN[n,7]) (csrc/forces_model.cuh layout)."""
    out = np.asarray(out)
    n = out.shape[0]
    return dict(c=out[:, 0:5], dc=out[:, 5:40].reshape(n, 5, 7), h=out[:, 40:50], dh=out[:, 50:120].reshape(n, 10, 7),
                f=out[:, 120], df=out[:, 121:128], fN=out[:, 128], dfN=out[:, 129:136])


def forces_eval(consts, z, p):
    z = np.ascontiguousarray(z, np.float64); p = np.ascontiguousarray(p, np.float64)
    consts = np.ascontiguousarray(consts, np.float64)
    n = z.shape[0]
    out = np.zeros((n, FORCES_OUT_WORDS))
    q = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hostsim_forces_eval(q(consts), q(z), q(p), q(out), n)
    return unpack_forces(out)


def solve_dual(cfg, xref, X, U, lam=None):
    """`mpcb200_solve_dual` on the emulator: returns (X, U, status, iters, lam [B, 14N+2])."""
    xref = np.ascontiguousarray(xref, np.float64)
    X = np.ascontiguousarray(X, np.float64).copy()
    U = np.ascontiguousarray(U, np.float64).copy()
    B, N = xref.shape[0], cfg.N
    lam = np.zeros((B, 14 * N + 2)) if lam is None else np.ascontiguousarray(lam, np.float64).copy()
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hostsim_solve_dual(C.byref(cfg), p(xref), p(X), p(U), p(lam), p(st), p(it), B)
    return X, U, st, it, lam


def forces_solve(cfg, weights_terminal, xinit, params, Zin=None, trace=0, road_boundaries=None, r_min=1.2):
    """The FORCESPRO-formulation solver core (csrc/forces_core.cuh) on the emulator: xinit [B,5], params [B,N,10], optional warm
    start Zin [B,N,7] -> (Z [B,N,7], status, iters)."""
    xinit = np.ascontiguousarray(np.atleast_2d(xinit), np.float64)
    B, N = xinit.shape[0], cfg.N
    params = np.ascontiguousarray(params, np.float64).reshape(B, N, 10)
    Pt = np.ascontiguousarray(weights_terminal, np.float64)
    Z = np.zeros((B, N, 7))
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    zin = None if Zin is None else np.ascontiguousarray(Zin, np.float64)
    if road_boundaries is None:
        lib().hostsim_forces_solve(C.byref(cfg), p(Pt), p(xinit), p(params), None if zin is None else p(zin), p(Z), p(st), p(it), B, trace)
    else:
        bl = np.ascontiguousarray(road_boundaries[0], np.float64); br = np.ascontiguousarray(road_boundaries[1], np.float64)
        f = lib().hostsim_forces_solve_rb
        f.argtypes = [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double]
        f(C.byref(cfg), p(Pt), p(xinit), p(params), None if zin is None else p(zin), p(Z), p(st), p(it), B, trace,
          p(bl), len(bl), p(br), len(br), float(r_min))
    return Z, st, it
