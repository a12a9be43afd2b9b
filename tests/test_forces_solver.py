"""The FORCESPRO-formulation solver (csrc/forces_core.cuh, `mpcb200_forces_solve`).

Oracle: oracle/forces_nlp.py (float64 restatement of /root/reference/MPC_Planner/optimizer.py:86-246 with complex-step
derivatives) solved by oracle/ipm.py (exact Hessian, sparse LU).  Its stage functions are pinned here against the reference's own
CasADi-generated C model (tests/golden/forces_model_kat.npz).  The closed-source FORCESPRO core itself (one BFGS QP per call) is not
reproducible: parity of the optimum is against the oracle, tolerance 1e-3 in float32 (stated in north_star), 1e-6 in float64.

CPU tests run the device code on the warp emulator (tests/host_sim); `-m gpu` tests call the C ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_sim"))
import hostsim  # noqa: E402
import mpc_b200  # noqa: E402
from oracle import forces_nlp as fn, ipm, nlp  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
TOL32, TOL64 = 1e-3, 1e-6


def _problem(name, N, seed=None, B=1):
    sc = mpc_b200.load_scenario(name)
    path, ori = np.asarray(sc.reference_path)[:, :2], np.asarray(sc.orientation)
    so = sc.static_obstacle
    oc = nlp.compute_centers_of_approximation_circles(so["position_x"], so["position_y"], so["length"], so["width"], so["orientation"])
    vel = fn.velocity_profile(sc.iter_length, N, sc.desired_velocity)
    P = fn.stage_parameters(0, N, path, ori, vel, oc)
    d0 = fn.make_nlp(N, sc.dt, sc.weights_setting, sc.x0, P, so)
    if seed is None:
        x0 = sc.x0[None, :]
    else:
        x0 = mpc_b200.perturbed_initial_states(sc, B, seed, r_clear=d0.r_sum + 0.05, obstacle_circles=oc, ego_offset=d0.ego_offset)
    return sc, d0, P, x0


def _emu_cfg(d, N, prec):
    cfg = hostsim.default_config(N, prec)
    cfg.Q[:] = list(d.Q); cfg.R[:] = list(d.R); cfg.r_sum = d.r_sum; cfg.dt = d.dt
    cfg.hessian = 1            # dynamics-curvature term on: B200ForcesproOptimizer's default
    return cfg


def _oracle(d0, x0):
    d = fn.ForcesData(**{**d0.__dict__, "xinit": np.asarray(x0, float)})
    r = ipm.solve(d, fn.initial_guess(d), model=fn)
    return d, r


def test_oracle_stage_functions_match_the_generated_c_model():
    """oracle/forces_nlp.py (complex-step derivatives) == FORCESNLPsolver_model.c known answers: pins the oracle's NLP functions."""
    k = np.load(os.path.join(G, "forces_model_kat.npz"))
    C = np.array([0.1, 2.5789128, 2.578, 0.75, 2, 2, 50, 0.1, 5, 2, 0.2, 4, 4, 100, 0.2, 10], float)      # weights baked into the generated model
    d = fn.ForcesData(N=2, dt=C[0], Q=C[4:9], R=C[9:11], Pt=C[11:16], xinit=np.zeros(5), params=np.zeros((2, 10)), r_sum=1.2,
                      ego_offset=C[3], l_wb=C[1], l_fric=C[2])
    r = fn.stage_eval(d, k["z"], k["p"])
    sc = lambda a: max(1.0, float(np.abs(a).max()))          # noqa: E731
    for a, b in (("c", "dynamics"), ("dc", "ddynamics"), ("h", "inequalities"), ("dh", "dinequalities")):
        assert np.abs(r[a] - k[b]).max() < 1e-10 * sc(k[b]), a
    assert np.abs(r["f"] - k["objective"][:, 0]).max() < 1e-10 * sc(k["objective"])
    assert np.abs(r["fN"] - k["objective"][:, 1]).max() < 1e-10 * sc(k["objective"])
    assert np.abs(r["df"] - k["dobjective"][:, 0]).max() < 1e-10 * sc(k["dobjective"])
    assert np.abs(r["dfN"] - k["dobjective"][:, 1]).max() < 1e-10 * sc(k["dobjective"])


def test_oracle_parameters_follow_the_reference_rules():
    # optimizer.py:291-311: velocity ramp over the last N steps, path points k+1.., replenished with the last point
    v = fn.velocity_profile(30, 10, 20.0)
    assert len(v) == 30 and (v[:20] == 20.0).all() and v[20] == 20.0 and v[-1] == 0.0 and np.allclose(np.diff(v[20:]), -20.0 / 9)
    path = np.stack([np.arange(30.0), np.zeros(30)], axis=1)
    P = fn.stage_parameters(25, 10, path, np.linspace(0, 1, 30), v, [[1, 2], [3, 4], [5, 6]])
    assert P.shape == (10, 10) and list(P[:, 0]) == [26, 27, 28, 29, 29, 29, 29, 29, 29, 29]
    assert P[4, 2] == v[29] and P[0, 2] == v[26] and list(P[3, 4:]) == [1, 2, 3, 4, 5, 6]


def test_oracle_ipm_agrees_with_scipy_trust_constr_on_the_forces_nlp():
    """Two independent solvers on the restated FORCESPRO-formulation NLP (N = 6, lane following)."""
    from scipy.optimize import minimize, NonlinearConstraint, Bounds
    _, d0, P, x0 = _problem("ZAM_Over-1_1_LF", 6)
    d, r = _oracle(d0, x0[0])
    assert r["status"] == 1 and r["kkt"] < 1e-8
    lbg, ubg, lbx, ubx = fn.g_bounds(d)
    con = NonlinearConstraint(lambda w: fn.g_fun(d, w), lbg, ubg, jac=lambda w: fn.g_jac(d, w).toarray())
    res = minimize(lambda w: fn.cost(d, w), fn.initial_guess(d), jac=lambda w: fn.cost_grad(d, w), method="trust-constr", constraints=[con],
                   bounds=Bounds(lbx, ubx), options=dict(gtol=1e-10, xtol=1e-12, maxiter=3000, barrier_tol=1e-12))
    assert np.abs(res.x - r["w"]).max() < 1e-5, np.abs(res.x - r["w"]).max()


@pytest.mark.parametrize("name,N", [("ZAM_Over-1_1_LF", 10), ("ZAM_Over-1_1_CA", 10), ("ZAM_Over-1_1_LF", 30)])
def test_emulated_device_core_matches_the_oracle(name, N):
    """The kernels' device code (forces_core.cuh) on the warp emulator == oracle: float64 to 1e-6, float32 to 1e-3; every point a KKT point."""
    _, d0, P, x0 = _problem(name, N)
    d, r = _oracle(d0, x0[0])
    assert r["status"] == 1
    Zo = fn.split(d, r["w"])
    for prec, tol in ((1, TOL64), (0, TOL32)):
        Z, st, it = hostsim.forces_solve(_emu_cfg(d, N, prec), d.Pt, x0, P)
        assert st[0] in (1, 3), st
        assert np.abs(Z[0] - Zo).max() < tol, (prec, np.abs(Z[0] - Zo).max())
        assert np.array_equal(Z[0, 0, 2:], x0[0])                         # the pinned stage comes back exactly
        assert Z[0, -1, 0] == pytest.approx(0.0, abs=tol) and Z[0, -1, 1] == pytest.approx(0.0, abs=tol)   # minimum-norm last inputs
    assert ipm.kkt_error(d, Z[0].reshape(-1), model=fn, act_tol=1e-4)[1]["primal"] < 1e-4


def test_emulated_core_collision_avoidance_n30_float64():
    """Active circle rows + active friction circle, N = 30 (the casadi path's config 3 geometry): KKT point, == oracle."""
    _, d0, P, x0 = _problem("ZAM_Over-1_1_CA", 30)
    d, r = _oracle(d0, x0[0])
    Z, st, it = hostsim.forces_solve(_emu_cfg(d, 30, 1), d.Pt, x0, P)
    assert st[0] == 1 and r["status"] == 1
    assert np.abs(Z[0] - fn.split(d, r["w"])).max() < TOL64
    h = fn.inequalities(d, Z[0], P)
    assert h[:, 1:].min() >= d.r_sum ** 2 - 1e-6 and h[:, 0].max() <= d.veh.a_max ** 2 + 1e-6
    assert h[:, 1:].min() < d.r_sum ** 2 + 1e-3                           # a circle row IS active on this instance


def test_emulated_core_warm_start_and_infeasible_start():
    _, d0, P, x0 = _problem("ZAM_Over-1_1_LF", 10)
    cfg = _emu_cfg(d0, 10, 1)
    Z, st, it = hostsim.forces_solve(cfg, d0.Pt, x0, P)
    Z2, st2, it2 = hostsim.forces_solve(cfg, d0.Pt, x0, P, Zin=Z)
    assert st2[0] == 1 and np.abs(Z2 - Z).max() < 1e-7          # (the barrier restarts at mu0: a warm start is not faster)
    bad = x0.copy(); bad[0, 2] = 0.5; bad[0, 3] = 30.0                  # v^2 tan(delta) / l = 190 > a_max: friction circle infeasible at xinit
    _, st3, _ = hostsim.forces_solve(cfg, d0.Pt, bad, P)
    assert st3[0] == -8


def _rb_problem(N, shift=-1.5):
    """Lane following with the reference path pushed 1.5 m towards the right road boundary: the boundary rows become active."""
    sc, d0, P, x0 = _problem("ZAM_Over-1_1_LF", N)
    P = P.copy(); P[:, 1] += shift
    rb = (sc.left_road_boundary, sc.right_road_boundary)
    d = fn.make_nlp(N, sc.dt, sc.weights_setting, x0[0], P, sc.static_obstacle, road_boundaries=rb)
    return sc, d, P, x0, rb


def test_road_boundaries_are_the_reference_lanelet_bounds():
    """configuration.py:432-433: right vertices of the network's 2nd / 1st lanelet (201 vertices each on ZAM_Over-1_1)."""
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    L, R = sc.left_road_boundary, sc.right_road_boundary
    assert L.shape == (201, 2) and R.shape == (201, 2)
    assert np.allclose(R[0], [0.0, -3.25]) and np.allclose(L[-1], [0.0, 3.25])       # a 6.5 m wide road around the x axis at its start
    assert mpc_b200.load_scenario("USA_Lanker-2_18_T-1_LF").left_road_boundary is None  # the hard-wired lanelet indices only mean something on ZAM_Over


def test_emulated_core_road_boundary_rows_match_the_oracle():
    """min-over-vertices distance rows (optimizer.py:18-30, 136-161), active on this instance; == oracle."""
    sc, d, P, x0, rb = _rb_problem(10)
    r = ipm.solve(d, fn.initial_guess(d), model=fn)
    assert r["status"] == 1
    Zo = fn.split(d, r["w"])
    assert abs(fn.inequalities(d, Zo, P)[:, 10:].min() - d.r_ego) < 1e-8            # a boundary row is active at the optimum
    for prec, tol in ((1, TOL64), (0, TOL32)):
        Z, st, it = hostsim.forces_solve(_emu_cfg(d, 10, prec), d.Pt, x0, P, road_boundaries=rb, r_min=d.r_ego)
        assert st[0] in (1, 3) and np.abs(Z[0] - Zo).max() < tol, (prec, st, np.abs(Z[0] - Zo).max())
        assert fn.inequalities(d, Z[0], P)[:, 10:].min() > d.r_ego - tol
    Zn, _, _ = hostsim.forces_solve(_emu_cfg(d, 10, 1), d.Pt, x0, P)                 # rows off: the ego cuts through the margin
    assert fn.inequalities(d, Zn[0], P)[:, 10:].min() < 0.5


# ----------------------------------------------------------------------------------------------------------------- GPU
def _gpu_opt(sc, N, precision, **kw):
    from mpc_b200.optimizer import make_configuration, init_values_from_state
    from mpc_b200.forces_optimizer import B200ForcesproOptimizer
    return B200ForcesproOptimizer(make_configuration(sc, N, framework_name="forcespro"), init_values_from_state(sc.x0), N,
                                  precision=precision, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,B", [("ZAM_Over-1_1_LF", 30, 256), ("ZAM_Over-1_1_CA", 30, 256), ("USA_Lanker-2_18_T-1_LF", 50, 128)])
def test_cuda_forces_solve_matches_oracle_and_emulator(name, N, B):
    sc, d0, P, x0 = _problem(name, N, seed=20261021, B=B)
    for precision, tol in (("f64", TOL64), ("f32", TOL32)):
        opt = _gpu_opt(sc, N, precision, max_batch=B)
        Z, st, it = opt.forces_solve_batch(x0, P)
        Z, st, it = Z.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
        conv = np.isin(st, (1, 3))
        # collision avoidance: a fraction of a percent of the instances crawl (barrier warm-up to mu_max next to an obstacle the
        # cold start drives through) and end at the iteration limit, status 0 -- reported, never silently wrong
        assert conv.mean() >= (0.99 if "CA" in name else 1.0), np.unique(st, return_counts=True)
        assert (st == 1).mean() > 0.95
        assert np.array_equal(Z[:, 0, 2:], x0)
        # converged instances: feasible to tolerance (defects, bounds, friction circle, circle distances)
        for b in range(0, B, max(1, B // 32)):
            if not conv[b]:
                continue
            d = fn.ForcesData(**{**d0.__dict__, "xinit": x0[b]})
            g = fn.g_fun(d, Z[b].reshape(-1))
            lbg, ubg, _, _ = fn.g_bounds(d)
            assert np.maximum(lbg - g, g - ubg).max() < (2e-3 if precision == "f32" else 1e-6)
        # sampled instances against the oracle
        worst = 0.0
        for b in (0, B // 3, B - 1):
            d, r = _oracle(d0, x0[b])
            if r["status"] != 1 or not conv[b]:
                continue
            worst = max(worst, np.abs(Z[b] - fn.split(d, r["w"])).max())
        assert worst < tol, (precision, worst)
        if precision == "f64":                                            # same source on the emulator: agreement to rounding of libm
            Ze, ste, ite = hostsim.forces_solve(_emu_cfg(d0, N, 1), d0.Pt, x0[:2], np.tile(P[None], (2, 1, 1)))
            assert np.abs(Ze - Z[:2]).max() < 1e-7 and list(ite) == list(it[:2])


@pytest.mark.gpu
def test_cuda_forces_solve_warm_start_misaligned_and_ragged():
    sc, d0, P, x0 = _problem("ZAM_Over-1_1_LF", 10, seed=5, B=67)          # ragged: 67 problems, 2 per CTA
    import torch
    opt = _gpu_opt(sc, 10, "f64", max_batch=128)
    Z, st, it = opt.forces_solve_batch(x0, P)
    assert (st == 1).all()
    Z2, st2, it2 = opt.forces_solve_batch(x0, P, Z_init=Z)
    assert (st2 == 1).all() and (Z2 - Z).abs().max() < 1e-7
    # parameter block at an address that is 8 (mod 16): the plain-load route instead of the TMA bulk copy, same numbers
    buf = torch.empty(67 * 10 * 10 + 1, dtype=torch.float64, device=opt.device)
    pm = buf[1:].view(67, 10, 10)
    pm.copy_(torch.as_tensor(P, device=opt.device).expand(67, 10, 10))
    assert pm.data_ptr() % 16 == 8
    Z3, st3, _ = opt.forces_solve_batch(x0, pm)
    assert torch.equal(Z3, Z) and torch.equal(st3, st)


@pytest.mark.gpu
def test_cuda_forcespro_optimizer_closed_loop_contract():
    """B200ForcesproOptimizer.optimize(): the reference's return contract (optimizer.py:368), RK4 plant transitions, tracking."""
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    opt = _gpu_opt(sc, 10, "f32", max_batch=8)
    x, u, tv = opt.optimize()
    T = sc.iter_length
    assert x.shape == (T, 5) and u.shape == (T, 2) and tv.shape == (T,)
    assert np.allclose(x[0], [sc.x0[0], sc.x0[1], 0.0, sc.x0[3], sc.x0[4]])
    for k in range(T - 1):
        assert np.abs(nlp.rk4_step(x[k], u[k], sc.dt) - x[k + 1]).max() < 1e-9
    assert np.abs(u[:, 0]).max() <= 0.4 + 1e-6 and np.abs(u[:, 1]).max() <= 11.5 + 1e-6
    path = np.asarray(sc.reference_path)[:, :2]
    dev = np.hypot(x[:, None, 0] - path[None, :, 0], x[:, None, 1] - path[None, :, 1]).min(axis=1)
    assert dev[5:].max() < 1.5                                            # the ego converges onto the reference path
    # the first solve of the loop against the oracle
    _, d0, P, x0 = _problem("ZAM_Over-1_1_LF", 10)
    d, r = _oracle(d0, x[0])
    assert np.abs(u[0] - fn.split(d, r["w"])[0, :2]).max() < TOL32


@pytest.mark.gpu
def test_cuda_road_boundary_rows_match_the_oracle_and_switch_off():
    """`mpcb200_forces_set_road_boundaries`: active boundary rows on the GPU == oracle; switching them off restores the plain solve."""
    N = 10
    sc, d, P, x0, rb = _rb_problem(N)
    r = ipm.solve(d, fn.initial_guess(d), model=fn)
    Zo = fn.split(d, r["w"])
    rng = np.random.default_rng(3)
    xb = x0 + rng.normal(size=(33, 5)) * np.array([0.3, 0.1, 0.005, 0.5, 0.01])      # a ragged batch around the nominal start; row 0 = nominal
    xb[0] = x0[0]
    for precision, tol in (("f64", TOL64), ("f32", TOL32)):
        opt = _gpu_opt(sc, N, precision, max_batch=64)
        Z0, st0, _ = opt.forces_solve_batch(xb, P)
        opt.set_road_boundaries()                                                    # the configuration's boundaries, r_min = radius_ego
        Z, st, it = opt.forces_solve_batch(xb, P)
        Zc, stc = Z.cpu().numpy(), st.cpu().numpy()
        assert np.isin(stc, (1, 3)).all(), stc
        assert np.abs(Zc[0] - Zo).max() < tol, (precision, np.abs(Zc[0] - Zo).max())
        for b in range(len(xb)):
            assert fn.inequalities(d, Zc[b], P)[1:, 10:].min() > d.r_ego - 2 * tol     # every instance keeps the margin (stage 0 is given)
        assert fn.inequalities(d, Z0.cpu().numpy()[0], P)[:, 10:].min() < 0.5         # ... which the plain solve does not
        opt.clear_road_boundaries()
        Z1, st1, _ = opt.forces_solve_batch(xb, P)
        assert (Z1 - Z0).abs().max() == 0.0                                           # rows off again: bit-identical to the first solve


@pytest.mark.gpu
def test_cuda_forces_closed_loop_on_device_equals_the_host_driven_loop():
    """`mpcb200_forces_closed_loop` (one launch, one ego per warp) == one `mpcb200_forces_solve` per step driven from Python."""
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    rng = np.random.default_rng(11)
    x0 = sc.x0[None, :] + rng.normal(size=(37, 5)) * np.array([0.3, 0.2, 0.005, 0.5, 0.02])
    x0[:, 2] = 0.0
    opt = _gpu_opt(sc, 10, "f64", max_batch=64)
    td, cd, sd, itd = opt.optimize_batch(x0, on_device=True)
    th, ch, sh, ith = opt.optimize_batch(x0, on_device=False)
    assert (sd == 1).all() and (sh == 1).all()
    assert np.abs(td - th).max() < 1e-6 and np.abs(cd - ch).max() < 1e-6
    assert np.array_equal(td[:, 0], x0)
    o32 = _gpu_opt(sc, 10, "f32", max_batch=64)
    t32, c32, s32, _ = o32.optimize_batch(x0)                       # float32 device loop (default route)
    assert np.isin(s32, (1, 3)).all() and np.abs(c32 - ch).max() < 5e-3 and np.abs(t32 - th).max() < 5e-3
