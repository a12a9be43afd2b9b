"""Pins the oracle (oracle/nlp.py, oracle/ipm.py) against everything the reference holds for this path:
recorded plant transitions, the CasADi-generated C model, the exact step-0 optimum, and an independent SLSQP solve."""
import os

import numpy as np
import pytest

from oracle import ipm, nlp

G = os.path.join(os.path.dirname(__file__), "golden")


def test_euler_transitions_recorded_casadi_runs_exact():
    d = np.load(os.path.join(G, "plant_transitions.npz"))
    xn = nlp.euler_step(d["casadi_x"], d["casadi_u"], 0.1)
    assert np.abs(xn - d["casadi_xnext"]).max() <= 1e-12     # 127 transitions, l = 2.5789128


def test_rk4_transitions_recorded_forcespro_runs_exact():
    d = np.load(os.path.join(G, "plant_transitions.npz"))
    xn = nlp.rk4_step(d["forcespro_x"], d["forcespro_u"], 0.1)
    assert np.abs(xn - d["forcespro_xnext"]).max() <= 1e-12


def test_recorded_initial_states():
    d = np.load(os.path.join(G, "plant_transitions.npz"))
    assert np.allclose(d["casadi_x"][0], [29.9948, -1.1501, 0.0, 20.0, 0.03495])
    assert np.allclose(d["casadi_x"][58], [0.0, 0.0, 0.0, 6.8062, -0.4268])


def test_forces_model_dynamics_and_jacobian():
    k = np.load(os.path.join(G, "forces_model_kat.npz"))
    z = k["z"]
    x, u = z[:, 2:7], z[:, 0:2]
    assert np.abs(nlp.rk4_step(x, u, 0.1) - k["dynamics"]).max() <= 1e-12
    # Jacobian of the RK4 step wrt z = [u; x] by central differences of the restated f
    eps = 1e-6
    for i in range(8):
        J = np.zeros((5, 7))
        for j in range(7):
            zp, zm = z[i].copy(), z[i].copy()
            zp[j] += eps
            zm[j] -= eps
            J[:, j] = (nlp.rk4_step(zp[2:], zp[:2], 0.1) - nlp.rk4_step(zm[2:], zm[:2], 0.1)) / (2 * eps)
        assert np.abs(J - k["ddynamics"][i]).max() <= 1e-6


def test_forces_model_friction_literal_and_circle_geometry():
    k = np.load(os.path.join(G, "forces_model_kat.npz"))
    z, p, h = k["z"], k["p"], k["inequalities"]
    a, de, v, psi = z[:, 1], z[:, 4], z[:, 5], z[:, 6]
    # friction term of the Forcespro formulation uses the same 2.578 literal as optimizer.py:378
    assert np.allclose(h[:, 0], a ** 2 + (v * (v * np.tan(de) / nlp.L_FRICTION)) ** 2, rtol=1e-13)
    # ego circle centres: compute_centers_of_approximation_circles with (4.508, 1.610) -> offset 0.75
    for i in range(len(z)):
        c, f, r = nlp.compute_centers_of_approximation_circles(z[i, 2], z[i, 3], 4.508, 1.610, psi[i])
        obst = p[i, 4:10].reshape(3, 2)
        d2 = [[(e[0] - o[0]) ** 2 + (e[1] - o[1]) ** 2 for o in obst] for e in (c, f, r)]
        assert np.allclose(h[i, 1:], np.array(d2).reshape(-1), rtol=1e-12)


def test_circle_radii_constants():
    assert nlp.compute_approximating_circle_radius(4.508, 1.610) == (pytest.approx(1.2), pytest.approx(3.0))
    assert nlp.compute_approximating_circle_radius(6.0, 3.5) == (pytest.approx(2.1), pytest.approx(4.0))
    assert nlp.compute_approximating_circle_radius(0.0, 0.0) == (0.0, 0.0)


def _zam(weights="LF", N=10):
    import mpc_b200
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF" if weights == "LF" else "ZAM_Over-1_1_CA")
    xref = np.tile(sc.x0, (N + 1, 1))
    return sc, nlp.make_nlp(N, sc.dt, sc.weights_setting, xref, sc.static_obstacle), xref


@pytest.mark.parametrize("weights", ["LF", "CA"])
def test_step0_known_answer_friction_row_active(weights):
    """Q3/Q4/Q7: first MPC step regulates to x0, wants to brake harder than allowed, is stopped at a0 = -sqrt(11.5)."""
    sc, d, xref = _zam(weights)
    r = ipm.solve(d, nlp.pack(np.zeros((10, 2)), xref))
    U, X = nlp.split(r["w"], 10)
    assert r["status"] == 1 and r["kkt"] < 1e-8
    assert abs(U[0, 1] + np.sqrt(11.5)) < 1e-8
    assert abs(U[0, 0]) < 1e-5
    if weights == "LF":   # profile found independently with SLSQP in the survey session (3 digits)
        assert np.allclose(U[:, 1], [-3.391, -30.63, -23.29, -16.72, -11.04, -6.38, -2.83, -0.51, 0.47, 0.0], atol=6e-3)
        assert np.allclose(X[:, 3], [20, 19.661, 16.598, 14.269, 12.598, 11.494, 10.856, 10.573, 10.522, 10.569, 10.569], atol=2e-3)


def test_verbatim_abs_friction_row_matches_smooth_statement_away_from_kink():
    sc, d, xref = _zam("LF")
    r1 = ipm.solve(d, nlp.pack(np.zeros((10, 2)), xref))
    d.friction_smooth = False
    r2 = ipm.solve(d, nlp.pack(np.zeros((10, 2)), xref))
    assert r2["status"] == 1 and np.abs(r1["w"] - r2["w"]).max() < 1e-7


def test_oracle_vs_scipy_slsqp_small():
    """Independent solver on the same NLP (single shooting over U, N=6)."""
    from scipy.optimize import minimize
    import mpc_b200
    N = 6
    sc, x0, xref, X, U = mpc_b200.make_batch("ZAM_Over-1_1_LF", 2, N, 7)
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref[0], sc.static_obstacle)
    r = ipm.solve(d, nlp.pack(U[0], X[0]))
    assert r["status"] == 1

    def rollout(u):
        u = u.reshape(N, 2)
        xs = [xref[0, 0]]
        for k in range(N):
            xs.append(nlp.euler_step(xs[-1], u[k], sc.dt))
        return np.array(xs)

    def f(u):
        return nlp.cost(d, nlp.pack(u.reshape(N, 2), rollout(u)))

    lbg, ubg, lbx, ubx = nlp.g_bounds(d)

    def ineq(u):
        xs = rollout(u)
        w = nlp.pack(u.reshape(N, 2), xs)
        g = nlp.g_fun(d, w)
        out = [ubg[0] - g[0], g[0] - lbg[0]]
        out += list(g[1 + 5 * (N + 1):] - d.r_sum)
        out += list(xs[:, 2] - d.veh.delta_min) + list(d.veh.delta_max - xs[:, 2]) + list(xs[:, 3]) + list(d.veh.v_max - xs[:, 3])
        return np.array(out)

    bnds = [(d.veh.deltav_min, d.veh.deltav_max), (None, d.veh.a_max)] * N
    s = minimize(f, np.zeros(2 * N), method="SLSQP", bounds=bnds, constraints=[{"type": "ineq", "fun": ineq}],
                 options={"ftol": 1e-14, "maxiter": 500})
    Uo, _ = nlp.split(r["w"], N)
    assert np.abs(s.x.reshape(N, 2) - Uo).max() < 2e-4
    assert abs(s.fun - r["obj"]) < 1e-6 * max(1.0, abs(r["obj"]))


def test_kkt_checker_rejects_perturbed_point():
    sc, d, xref = _zam("LF")
    r = ipm.solve(d, nlp.pack(np.zeros((10, 2)), xref))
    assert ipm.kkt_error(d, r["w"])[0] < 1e-8
    w = r["w"].copy()
    w[3] += 1e-3
    assert ipm.kkt_error(d, w)[0] > 1e-5


def test_golden_nlp_solutions_reproduce():
    g = np.load(os.path.join(G, "nlp_solutions.npz"))
    import mpc_b200
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    N = 30
    xref = g["lf_zam_n30_xref"][3]
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref, sc.static_obstacle)
    w = nlp.pack(g["lf_zam_n30_U"][3], g["lf_zam_n30_X"][3])
    assert ipm.kkt_error(d, w)[0] < 1e-7
    r = ipm.solve(d, nlp.pack(np.zeros((N, 2)), np.tile(xref[0], (N + 1, 1))))
    assert np.abs(r["w"] - w).max() < 1e-6


def test_reference_window_rule_q8():
    T, N = 70, 10
    path = np.stack([np.arange(T, dtype=float), np.arange(T, dtype=float) * 2], axis=1)
    orient = np.arange(T, dtype=float) * 0.01
    x = np.arange(5, dtype=float)
    w = nlp.reference_window(5, x, N, T, path, orient, 7.0)
    assert w.shape == (N + 1, 5) and np.all(w[0] == x)
    assert np.all(w[1:, 0] == np.arange(6, 16)) and np.all(w[1:, 3] == 7.0) and np.all(w[1:, 2] == 0.0)
    w = nlp.reference_window(65, x, N, T, path, orient, 7.0)     # i >= T - N: frozen at path[60:70]
    assert np.all(w[1:, 0] == np.arange(60, 70))
    w30 = nlp.reference_window(0, x, 30, 30, path[:30], orient[:30], 7.0)   # N == T: path[0:30] at every step
    assert np.all(w30[1:, 0] == np.arange(0, 30))


@pytest.mark.parametrize("key,name,sigma", [("casadi_zam_lf", "ZAM_Over-1_1_LF", 0.1), ("casadi_zam_ca", "ZAM_Over-1_1_CA", 0.05),
                                            ("casadi_lanker_lf", "USA_Lanker-2_18_T-1_LF", 0.1)])
def test_recorded_ipopt_controls_pin_the_oracle_optimum_statistically(key, name, sigma):
    """SURVEY 8c (6), extended to every recorded step.  The reference's recorded CasADi/IPOPT closed loops (N = 10,
    /root/reference/test/2D_plots_casadi_ZAM_Over-1_1_*/) applied u = u*_0 + N(0, sigma^2) (optimizer.py:611-617) at the
    recorded state x_k.  Re-solving the restated NLP at every recorded (x_k, window k) must therefore leave residuals
    u_rec - u*_0(oracle) that look like that noise: zero median within its standard error, robust spread = sigma.
    A wrong cost pairing (Q2), a terminal cost (Q1), a wrong window (Q8) or friction row (Q3) shifts u*_0 by far more.
    A few steps are outliers in the RECORDING (IPOPT's return status is never checked, optimizer.py:607-609), so the
    statistics are robust ones.  All three recorded CasADi runs are used: since the route planner's reference path is restated
    exactly (tests/test_results_format.py) the 70-step USA_Lanker run -- different weights, a left turn and two lane changes,
    both friction-limited phases -- pins the NLP as well as the two ZAM_Over runs (residual mean 0.015 / 0.000, spread 0.10 / 0.085
    for sigma = 0.1; with the approximate path of round 1 the acceleration residual had a mean of +0.23)."""
    import mpc_b200
    from oracle import ipm
    g = np.load(os.path.join(G, "recorded_runs.npz"))
    sc = mpc_b200.load_scenario(name)
    Xr, Ur = g[key + "_x"], g[key + "_u"]
    N, T = 10, sc.iter_length
    res = []
    for i, x in enumerate(Xr):
        xref = np.tile(x, (N + 1, 1)) if i == 0 else nlp.reference_window(i - 1, x, N, T, sc.reference_path, sc.orientation,
                                                                          sc.desired_velocity)
        d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref, sc.static_obstacle)
        r = ipm.solve(d, nlp.pack(np.zeros((N, 2)), np.tile(x, (N + 1, 1))))
        if r["status"] == 1:
            res.append(Ur[i] - nlp.split(r["w"], N)[0][0])
    res = np.array(res)
    assert len(res) >= len(Xr) - 2
    se = sigma / np.sqrt(len(res))
    for c in range(2):
        e = res[:, c]
        inl = np.abs(e) < 4 * sigma
        assert inl.mean() >= 0.85                                              # at most a few recording outliers
        assert abs(np.median(e)) < 4 * 1.2533 * se                             # median of n normals: se * sqrt(pi/2)
        assert abs(e[inl].mean()) < 4 * se
        mad_sigma = 1.4826 * np.median(np.abs(e - np.median(e)))
        assert 0.5 * sigma < mad_sigma < 1.5 * sigma


def _trust_constr(d, w0):
    """scipy's trust-region interior-point method on the restated NLP (analytic Jacobian / Hessian from oracle.nlp): an
    implementation that shares no code with oracle/ipm.py.  The 3x duplicated obstacle rows (Q6) are passed once (LICQ)."""
    import scipy.sparse as sp
    from scipy.optimize import Bounds, NonlinearConstraint, minimize
    N = d.N
    lbg, ubg, lbx, ubx = nlp.g_bounds(d)
    keep = np.concatenate([np.arange(1 + 5 * (N + 1)), 1 + 5 * (N + 1) + np.arange(0, 9 * (N + 1), 3)])
    con = NonlinearConstraint(lambda w: nlp.g_fun(d, w)[keep], lbg[keep], ubg[keep], jac=lambda w: nlp.g_jac(d, w)[keep],
                              hess=lambda w, v: nlp.lag_hess(d, w, np.bincount(keep, weights=v, minlength=d.m), 0.0))
    return minimize(lambda w: nlp.cost(d, w), w0, jac=lambda w: nlp.cost_grad(d, w), hess=lambda w: sp.diags(nlp.cost_hess_diag(d)),
                    method="trust-constr", constraints=[con], bounds=Bounds(lbx, ubx),
                    options=dict(xtol=1e-12, gtol=1e-10, barrier_tol=1e-10, maxiter=3000))


@pytest.mark.parametrize("name,N,b", [("ZAM_Over-1_1_LF", 30, 0), ("USA_Lanker-2_18_T-1_LF", 50, 1)])
def test_oracle_vs_scipy_trust_constr_at_benchmark_horizons(name, N, b):
    """two independent solvers on the restated NLP at the benchmark sizes (N = 30 / 50, multiple shooting, all 14N+15
    constraint rows): oracle/ipm.py and scipy trust-constr agree to 1e-6 from the same cold start."""
    import mpc_b200
    sc, x0, xref, X0, U0 = mpc_b200.make_batch(name, 4, N, 20261017)
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref[b], sc.static_obstacle)
    w0 = nlp.pack(U0[b], X0[b])
    r = ipm.solve(d, w0)
    s = _trust_constr(d, w0)
    assert r["status"] == 1 and s.constr_violation < 1e-9
    assert np.abs(s.x - r["w"]).max() < 1e-6 and abs(s.fun - r["obj"]) < 1e-7 * max(1.0, abs(r["obj"]))


def test_collision_avoidance_solver_core_point_is_a_local_optimum_for_scipy_too():
    """Collision avoidance is multi-modal from a cold start, so the independent check is local: scipy trust-constr started AT
    the point the solver core (float64 arithmetic, tests/host_sim) converges to must stay there."""
    import hostsim
    import mpc_b200
    from test_host_logic import _cfg
    N = 30
    sc, x0, xref, X0, U0 = mpc_b200.make_batch("ZAM_Over-1_1_CA", 4, N, 20261018)
    X, U, st, it, _ = hostsim.solve(_cfg(sc, N, 1, max_iter=300), xref[:1], X0[:1], U0[:1])
    assert st[0] == 1
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref[0], sc.static_obstacle)
    w = nlp.pack(U[0], X[0])
    assert (nlp.g_fun(d, w)[1 + 5 * (N + 1):].min() - d.r_sum) < 1e-6          # the obstacle row is active at this point
    s = _trust_constr(d, w)
    assert s.constr_violation < 1e-9 and np.abs(s.x - w).max() < 1e-5 and s.fun >= nlp.cost(d, w) - 1e-6
