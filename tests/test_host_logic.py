"""CPU tests of the host side and of the solver core's ALGORITHM and LANE MAPPINGS (tests/host_sim = csrc/warp_core.cuh,
the device code of the warp-per-problem CUDA kernels, compiled with g++ on a 32-fiber lock-step warp emulator; test
tooling only -- the product has no CPU path)."""
import os

import numpy as np
import pytest

import hostsim
import mpc_b200
from mpc_b200 import optimizer as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _cfg(sc, N, prec, hess=0, **opts):
    circles, r_sum, off = O.obstacle_circles_and_radius(sc.static_obstacle)
    cfg = hostsim.default_config(N, prec)
    cfg.hessian = hess
    cfg.dt = sc.dt
    w = sc.weights_setting
    for i, k in enumerate(("weight_x", "weight_y", "weight_steering_angle", "weight_velocity", "weight_heading_angle")):
        cfg.Q[i] = w[k]
    cfg.R[0], cfg.R[1] = w["weight_velocity_steering_angle"], w["weight_long_acceleration"]
    cfg.r_sum, cfg.ego_offset = r_sum, off
    for j in range(3):
        cfg.obstacle[2 * j], cfg.obstacle[2 * j + 1] = circles[j]
    for k, v in opts.items():
        setattr(cfg, k, v)
    return cfg


@pytest.mark.parametrize("key,name,N", [("lf_zam_n30", "ZAM_Over-1_1_LF", 30), ("lf_lanker_n50", "USA_Lanker-2_18_T-1_LF", 50),
                                         ("lf_zam_n10", "ZAM_Over-1_1_LF", 10)])
@pytest.mark.parametrize("prec,tol", [(0, 1e-3), (1, 1e-6)])
@pytest.mark.parametrize("hess", [0, 1])
def test_solver_core_matches_oracle_golden(key, name, N, prec, tol, hess):
    g = np.load(os.path.join(G, "nlp_solutions.npz"))
    sc = mpc_b200.load_scenario(name)
    xref = g[key + "_xref"]
    B = xref.shape[0]
    X0 = np.repeat(xref[:, :1, :], N + 1, axis=1)
    X, U, st, it, _ = hostsim.solve(_cfg(sc, N, prec, hess), xref, X0, np.zeros((B, N, 2)))
    assert (st == 1).all(), st
    assert np.abs(U - g[key + "_U"]).max() < tol        # ||dU||inf  (rad/s, m/s^2)
    assert np.abs(X - g[key + "_X"]).max() < tol        # ||dX||inf  (m, rad, m/s)


def test_solver_core_step0_known_answer():
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    N = 10
    xref = np.tile(sc.x0, (1, N + 1, 1))
    for prec, tol in ((0, 1e-4), (1, 1e-7)):
        X, U, st, it, _ = hostsim.solve(_cfg(sc, N, prec), xref, xref, np.zeros((1, N, 2)))
        assert st[0] == 1
        assert abs(U[0, 0, 1] + np.sqrt(11.5)) < tol and abs(U[0, 0, 0]) < 1e-4


def test_solver_core_flags_infeasible_pinned_stage():
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    N = 10
    x0 = sc.x0.copy()
    x0[2] = 0.5          # v0^2 tan(delta0)/2.578 = 84 > a_max: friction row infeasible (quirk Q3)
    xref = np.tile(x0, (1, N + 1, 1))
    _, _, st, _, _ = hostsim.solve(_cfg(sc, N, 1), xref, xref, np.zeros((1, N, 2)))
    assert st[0] == -8


def test_collision_avoidance_solution_is_a_kkt_point_of_the_reference_nlp():
    from oracle import nlp, ipm
    N, B = 30, 3
    sc, x0, xref, X, U = mpc_b200.make_batch("ZAM_Over-1_1_CA", B, N, 20261018)
    Xs, Us, st, it, _ = hostsim.solve(_cfg(sc, N, 1, max_iter=200), xref, X, U)
    assert (st == 1).all()
    for b in range(B):
        d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref[b], sc.static_obstacle)
        w = nlp.pack(Us[b], Xs[b])
        assert ipm.kkt_error(d, w)[0] < 1e-6
        assert (nlp.g_fun(d, w)[1 + 5 * (N + 1):] >= d.r_sum - 1e-7).all()      # keeps >= 3.3 m from the obstacle
        r = ipm.solve(d, w)                                                      # oracle warm-started at that point
        assert r["status"] == 1 and np.abs(r["w"] - w).max() < 1e-5


def test_rollout_start_reaches_the_same_lane_following_optimum():
    """init_rollout = 1 replaces the caller's X warm start by the Euler rollout of U from the pinned state."""
    N, B = 30, 4
    sc, x0, xref, X, U = mpc_b200.make_batch("ZAM_Over-1_1_LF", B, N, 20261017)
    Xa, Ua, sta, _, _ = hostsim.solve(_cfg(sc, N, 1), xref, X, U)
    Xb, Ub, stb, _, _ = hostsim.solve(_cfg(sc, N, 1, init_rollout=1), xref, np.full_like(X, 1e3), U)     # X is ignored
    assert (sta == 1).all() and (stb == 1).all()
    assert np.abs(Ua - Ub).max() < 1e-6 and np.abs(Xa - Xb).max() < 1e-6


def test_exact_hessian_falls_back_and_agrees_with_gauss_newton():
    N, B = 30, 4
    sc, x0, xref, X, U = mpc_b200.make_batch("USA_Lanker-2_18_T-1_LF", B, 50, 20261019)
    Xa, Ua, sta, _, _ = hostsim.solve(_cfg(sc, 50, 1, hess=0), xref, X, U)
    Xb, Ub, stb, _, _ = hostsim.solve(_cfg(sc, 50, 1, hess=1), xref, X, U)
    assert (sta == 1).all() and (stb == 1).all()
    assert np.abs(Ua - Ub).max() < 1e-6 and np.abs(Xa - Xb).max() < 1e-6


def test_scenarios_and_batch_generator():
    names = mpc_b200.scenario_names()
    assert {"ZAM_Over-1_1_LF", "ZAM_Over-1_1_CA", "USA_Lanker-2_18_T-1_LF"} <= set(names)
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    assert sc.iter_length == 30 and abs(sc.desired_velocity - 19.9995) < 1e-9 and sc.dt == 0.1
    assert np.allclose(sc.reference_path[0], [29.9948, -1.1501]) and np.allclose(sc.reference_path[-1], [87.8, 3.3])
    assert mpc_b200.load_scenario("USA_Lanker-2_18_T-1_LF").iter_length == 70
    a = mpc_b200.make_batch("ZAM_Over-1_1_LF", 64, 30, 20261017)
    b = mpc_b200.make_batch("ZAM_Over-1_1_LF", 64, 30, 20261017)
    assert np.array_equal(a[1], b[1]) and a[2].shape == (64, 31, 5)
    x0 = a[1]
    assert (np.abs(x0[:, 2]) <= 0.05).all() and (x0[:, 3] >= 1).all() and (x0[:, 3] <= 25).all()
    assert np.array_equal(a[2][:, 0], x0) and np.allclose(a[2][:, 1:, :2], sc.reference_path[None])
    ca = mpc_b200.make_batch("ZAM_Over-1_1_CA", 256, 30, 20261018)
    circles, r_sum, off = O.obstacle_circles_and_radius(ca[0].static_obstacle)
    assert abs(r_sum - 3.3) < 1e-12 and off == 0.75
    with pytest.raises(ValueError):
        mpc_b200.make_batch("ZAM_Over-1_1_LF", 4, 31, 1)


def test_reference_window_matches_oracle_rule():
    from oracle import nlp
    sc = mpc_b200.load_scenario("USA_Lanker-2_18_T-1_LF")
    x = np.array([[1.0, 2.0, 0.1, 5.0, 0.3], [0.0, 0.0, 0.0, 6.0, -0.4]])
    for i in (0, 5, 19, 20, 21, 60, 69):
        for N in (10, 50):
            w = mpc_b200.reference_window(i, x, N, sc.iter_length, sc.reference_path, sc.orientation, sc.desired_velocity)
            for b in range(2):
                assert np.array_equal(w[b], nlp.reference_window(i, x[b], N, sc.iter_length, sc.reference_path, sc.orientation,
                                                                 sc.desired_velocity))


def test_optimizer_base_mirrors_reference_attributes():
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_CA")
    conf = O.make_configuration(sc, 30)
    base = O.OptimizerBase(conf, O.init_values_from_state(sc.x0), 30)
    assert (base.delta_min, base.delta_max, base.deltav_min, base.deltav_max) == (-1.066, 1.066, -0.4, 0.4)
    assert (base.v_min, base.v_max, base.a_max) == (0, 50.8, 11.5)
    assert base.iter_length == 30 and base.predict_horizon == 30 and base.delta_t == 0.1
    assert abs(base.radius_ego - 1.2) < 1e-12 and abs(base.radius_obstacle - 2.1) < 1e-12
    c = base.obstacle_circles_centers_tuple
    assert np.allclose(c[0], [59.948, 0.08323]) and np.allclose(c[1], [59.948 + np.cos(0.07759), 0.08323 + np.sin(0.07759)])
    for m in ("equal_constraints", "inequal_constraints", "cost_function", "solver", "optimize"):
        assert callable(getattr(base, m))
    assert issubclass(O.B200Optimizer, O.OptimizerBase) or O._Base is not O.OptimizerBase


def test_shard_ranges_cover_batch():
    from mpc_b200.sharding import shard_range, shard_sizes
    for B in (1, 7, 1024, 4097):
        for W in (1, 2, 4, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B and all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            assert sum(shard_sizes(B, W)) == B and max(shard_sizes(B, W)) - min(shard_sizes(B, W)) <= 1


@pytest.mark.parametrize("name,N", [("ZAM_Over-1_1_LF", 30), ("USA_Lanker-2_18_T-1_LF", 50), ("ZAM_Over-1_1_CA", 30)])
def test_dual_block_round_trip_warm_starts_the_solve(name, N):
    """`mpcb200_solve_dual` semantics on the emulator: the dual block written by a solve (multipliers, obstacle slacks, mu, valid
    flag) re-imported together with the primal solution restarts the barrier at mu_warm and converges to the same point in a
    few iterations; an invalid (zero) block means cold duals = the plain solve."""
    B = 6
    sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, B, N, 123)
    cfg = _cfg(sc, N, 1, max_iter=200)
    Xa, Ua, sta, ita, _ = hostsim.solve(cfg, xref, X0, U0)
    Xb, Ub, stb, itb, lam = hostsim.solve_dual(cfg, xref, X0, U0)
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub) and np.array_equal(ita, itb)          # zero block = cold duals
    ok = stb == 1
    assert ok.sum() >= B - 1 and (lam[ok, -1] == 1.0).all() and (lam[ok, :-2] >= 0).all() and (lam[ok, -2] <= cfg.mu_min * 1.01).all()
    Xc, Uc, stc, itc, lam2 = hostsim.solve_dual(cfg, xref, Xb, Ub, lam)
    assert (stc[ok] == 1).all()
    assert np.abs(Uc - Ub)[ok].max() < 1e-5 and np.abs(Xc - Xb)[ok].max() < 1e-5
    assert itc[ok].mean() <= 9.0 and (itc[ok] < itb[ok]).all()          # float64: mu_warm = 1e-4 -> 1e-9 alone takes ~5 reductions


def test_final_phase_extrapolation_cuts_the_halving_tail():
    """Problem 948 of the bench batch (BASELINE configs[1], seed 20261017): a weakly active stage-0 row makes Newton's step halve
    per iteration at the final barrier parameter (11 iterations before the rule).  The doubled step on the detected halving
    sequence + the rate-based exit bring it to <= 8; the solution still equals the float64 oracle within the stated tolerance, and
    over the first 128 instances nothing needs more than 10 iterations."""
    from oracle import nlp, ipm
    sc, x0, xref, X, U = mpc_b200.make_batch("ZAM_Over-1_1_LF", 1024, 30, 20261017)
    cfg = _cfg(sc, 30, 0)
    Xs, Us, st, it, _ = hostsim.solve(cfg, xref[948:949], X[948:949], U[948:949])
    assert st[0] == 1 and it[0] <= 8, (st, it)
    d = nlp.make_nlp(30, sc.dt, sc.weights_setting, xref[948], sc.static_obstacle)
    r = ipm.solve(d, nlp.pack(U[948], X[948]))
    Uo, Xo = nlp.split(r["w"], 30)
    assert r["status"] == 1 and np.abs(Us[0] - Uo).max() < 1e-3 and np.abs(Xs[0] - Xo).max() < 1e-3
    _, _, st, it, _ = hostsim.solve(cfg, xref[:128], X[:128], U[:128])
    assert (st == 1).all() and it.max() <= 10 and it.mean() < 5.4
