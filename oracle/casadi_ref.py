"""ORACLE (test infrastructure, NOT product code) -- the reference NLP built LITERALLY with casadi and solved by IPOPT.

This is the branch SURVEY.md 8(c) asks for: when `casadi` is importable on the box that runs the tests / the bench, the NLP of
`CasadiOptimizer` is constructed symbol for symbol the way the reference constructs it and handed to `ca.nlpsol('ipopt')` with
the reference's options, so that the float64 restatement (oracle/nlp.py + oracle/ipm.py) and the CUDA solver can be compared
with IPOPT's own answer (tests/test_casadi_parity.py; `bench.py --impl reference` uses it as the reference arm, `kind: "ipopt"`).

casadi (>= 3.5.1, /root/reference/requirments:5, bundles IPOPT + MUMPS) is an un-vendored third-party dependency that is NOT in
this image: `available()` is False here and on the GPU boxes of this pool (no network, not in /opt/wheelhouse -- recorded in
DESIGN.md), every consumer skips / falls back to the float64 port and says so.  Until a box with casadi runs these tests the
status stays *parity unpinned at the IPOPT boundary* (oracle/nlp.py header).

What each function follows (paths relative to /root/reference/MPC_Planner/):
    build_solver()        optimizer.py:513-560  (symbols :522-542, cost :545 -> :493-511, constraints :548 -> :373-411,
                                                 variable / parameter order :550-552, IPOPT options :556, nlpsol :558)
    bounds()              optimizer.py:413-491  (lbg, ubg, lbx, ubx as the reference's Python lists)
    solve_instance()      optimizer.py:600-609  (p = [vec U_ref; vec X_ref], x0 = [vec U; vec X], sol(...), res['x'])
    closed_loop()         optimizer.py:562-643  (incl. the scrambled initial guess of quirk Q5 when `scramble=True`)
"""
import numpy as np

from . import nlp

try:                                     # the ONLY import of casadi in this repository
    import casadi as ca
except Exception:                        # absent (this image): every entry point below raises / reports unavailable
    ca = None

_P_L, _P_W = 4.508, 1.610                # parameters_vehicle2: p.l, p.w (optimizer.py:385 passes them to the circle helper)
_SOLVER_CACHE = {}


def available():
    return ca is not None


def _require():
    if ca is None:
        raise RuntimeError("casadi is not installed: the verbatim casadi/IPOPT branch of the oracle is unavailable")


def _ks_casadi(x, u):
    """VehicleDynamics.KS_casadi, configuration.py:353-368 (l = p.a + p.b of parameters_vehicle2)."""
    l = nlp.L_WB
    return ca.vertcat(x[3] * ca.cos(x[4]), x[3] * ca.sin(x[4]), u[0], u[1], x[3] / l * ca.tan(x[2]))


def _ego_circles(x_position, y_position, orientation):
    """compute_centers_of_approximation_circles with casadi symbols, configuration.py:69-93."""
    _, disc_distance = nlp.compute_approximating_circle_radius(_P_L, _P_W)
    distance_centers = disc_distance / 2
    center = [x_position, y_position]
    center_fw = [x_position + (distance_centers / 2) * ca.cos(orientation), y_position + (distance_centers / 2) * ca.sin(orientation)]
    center_rw = [x_position - (distance_centers / 2) * ca.cos(orientation), y_position - (distance_centers / 2) * ca.sin(orientation)]
    return center, center_fw, center_rw


def build_solver(N, dt, weights, static_obstacle):
    """CasadiOptimizer.solver(), optimizer.py:513-560, with equal_constraints (:373-411) and cost_function (:493-511) inlined
    in the reference's order.  Returns (nlpsol object, f)."""
    _require()
    states = ca.vertcat(*[ca.SX.sym(n) for n in ("sx", "sy", "delta", "vel", "Psi")])
    controls = ca.vertcat(*[ca.SX.sym(n) for n in ("u0", "u1")])
    f = ca.Function("f", [states, controls], [_ks_casadi(states, controls)], ["input_state", "control_input"], ["rhs"])
    U = ca.SX.sym("U", 2, N)
    X = ca.SX.sym("X", 5, N + 1)
    U_ref = ca.SX.sym("U_ref", 2, N)
    X_ref = ca.SX.sym("X_ref", 5, N + 1)
    # ---- cost_function (optimizer.py:493-511): the terminal term sits on a statement of its own (unary +) and is never added (Q1)
    Q = np.diag([weights["weight_x"], weights["weight_y"], weights["weight_steering_angle"], weights["weight_velocity"],
                 weights["weight_heading_angle"]])
    R = np.diag([weights["weight_velocity_steering_angle"], weights["weight_long_acceleration"]])
    obj = 0
    for i in range(N):
        e = X[:, i] - X_ref[:, i + 1]
        obj = obj + ca.mtimes([e.T, Q, e]) + ca.mtimes([U[:, i].T, R, U[:, i]])
    # ---- equal_constraints (optimizer.py:373-411); `controls[1]`, `states[3]`, `states[2]` are LINEAR indices into the SX
    # matrices (column-major): U[1,0], X[3,0], X[2,0] (Q3)
    oc = nlp.compute_centers_of_approximation_circles(static_obstacle["position_x"], static_obstacle["position_y"],
                                                      static_obstacle["length"], static_obstacle["width"],
                                                      static_obstacle["orientation"])
    g = [ca.sqrt(((U[1]) ** 2 + (X[3] * ((ca.tan(X[2]) * X[3]) / 2.578))) ** 2), X[:, 0] - X_ref[:, 0]]
    for i in range(N):
        x_next_ = f(X[:, i], U[:, i]) * dt + X[:, i]
        g.append(X[:, i + 1] - x_next_)
    for i in range(N + 1):
        ego = _ego_circles(X[0, i], X[1, i], X[4, i])
        for j in range(3):                       # centre, front, rear -- each distance appended three times (Q6)
            d = ca.sqrt((ego[j][0] - oc[j][0]) ** 2 + (ego[j][1] - oc[j][1]) ** 2)
            g += [d, d, d]
    opt_variables = ca.vertcat(ca.reshape(U, -1, 1), ca.reshape(X, -1, 1))
    opt_params = ca.vertcat(ca.reshape(U_ref, -1, 1), ca.reshape(X_ref, -1, 1))
    nlp_prob = {"f": obj, "x": opt_variables, "p": opt_params, "g": ca.vcat(g)}
    opts_setting = {"ipopt.max_iter": 100, "ipopt.print_level": 0, "print_time": 0, "ipopt.acceptable_tol": 1e-8,
                    "ipopt.acceptable_obj_change_tol": 1e-6}
    return ca.nlpsol("solver", "ipopt", nlp_prob, opts_setting), f


def bounds(N, static_obstacle, veh=None):
    """inequal_constraints(), optimizer.py:413-491."""
    veh = veh or nlp.VehicleParams()
    r_obs, _ = nlp.compute_approximating_circle_radius(static_obstacle["length"], static_obstacle["width"])
    r_ego, _ = nlp.compute_approximating_circle_radius(veh.length, veh.width)
    lbg, ubg = [0.0], [veh.a_max]
    for _ in range(N + 1):
        lbg += [0.0] * 5
        ubg += [0.0] * 5
    for _ in range(N + 1):
        lbg += [r_ego + r_obs] * 9
        ubg += [np.inf] * 9
    lbx, ubx = [], []
    for _ in range(N):
        lbx += [veh.deltav_min, -np.inf]
        ubx += [veh.deltav_max, veh.a_max]
    for _ in range(N + 1):
        lbx += [-np.inf, -np.inf, veh.delta_min, veh.v_min, -np.inf]
        ubx += [np.inf, np.inf, veh.delta_max, veh.v_max, np.inf]
    return lbg, ubg, lbx, ubx


def solve_instance(sc, N, xref, X_init, U_init, rebuild=False):
    """One call `sol(x0=init_control, p=c_p, lbg, lbx, ubg, ubx)` (optimizer.py:600-609) for the parameter block `xref`
    [N+1,5] and the initial guess (U_init [N,2], X_init [N+1,5]; stage-major like w, i.e. NOT scrambled).
    rebuild=True re-creates the nlpsol object like the reference does every MPC step (quirk Q10).
    Returns (w* in the reference's order [vec U; vec X], IPOPT return_success)."""
    _require()
    key = (sc.name, N)
    if rebuild or key not in _SOLVER_CACHE:
        _SOLVER_CACHE[key] = build_solver(N, sc.dt, sc.weights_setting, sc.static_obstacle)
    sol, _ = _SOLVER_CACHE[key]
    lbg, ubg, lbx, ubx = bounds(N, sc.static_obstacle)
    c_p = np.concatenate((np.zeros((N, 2)).reshape(-1, 1), np.asarray(xref, float).reshape(-1, 1)))
    init = np.concatenate((np.asarray(U_init, float).reshape(-1, 1), np.asarray(X_init, float).reshape(-1, 1)))
    res = sol(x0=init, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
    return res["x"].full().reshape(-1), bool(sol.stats().get("success", False))


def closed_loop(sc, N, scramble=True):
    """CasadiOptimizer.optimize(), optimizer.py:562-643, noise-free.  scramble=True reproduces the reference's component-major
    initial guess (quirk Q5: `u0.T.reshape(-1,1)` / `next_states.T.reshape(-1,1)` on arrays that are already stage-major).
    Returns (traj_s [T,5], u [T,2]) (Q12)."""
    _require()
    T = sc.iter_length
    init_state = np.array(sc.x0, float).reshape(-1, 1)
    current_state = init_state.copy()
    u0 = np.array([0.0, 0.0] * N).reshape(-1, 2).T
    next_trajectories = np.tile(current_state.reshape(1, -1), N + 1).reshape(N + 1, -1)
    next_states = next_trajectories.copy()
    next_controls = np.zeros((N, 2))
    lbg, ubg, lbx, ubx = bounds(N, sc.static_obstacle)
    u_c, traj = [], []
    for i in range(T):
        c_p = np.concatenate((next_controls.reshape(-1, 1), next_trajectories.reshape(-1, 1)))
        if scramble:
            init_control = np.concatenate((u0.T.reshape(-1, 1), next_states.T.reshape(-1, 1)))
        else:
            uu = u0 if u0.shape == (N, 2) else u0.T
            xx = next_states if next_states.shape == (N + 1, 5) else next_states.T
            init_control = np.concatenate((uu.reshape(-1, 1), xx.reshape(-1, 1)))
        sol, f = build_solver(N, sc.dt, sc.weights_setting, sc.static_obstacle)              # rebuilt every step (Q10)
        res = sol(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        est = res["x"].full()
        u0 = est[:2 * N].reshape(N, 2).T
        x_m = est[2 * N:].reshape(N + 1, 5).T
        u_c.append(u0[:, 0].copy())
        # shift_movement, optimizer.py:645-655
        st = current_state + sc.dt * f(current_state, u0[:, 0]).full()
        u0 = np.concatenate((u0[:, 1:], u0[:, -1:]), axis=1).T
        next_states = np.concatenate((x_m[:, 1:], x_m[:, -1:]), axis=1)
        current_state = st.reshape(-1, 1)
        next_trajectories = nlp.reference_window(i, current_state.reshape(-1), N, T, sc.reference_path, sc.orientation,
                                                 sc.desired_velocity)
        traj.append(current_state.reshape(-1).copy())
    traj_s = np.insert(np.array(traj), 0, init_state.T, axis=0)[:-1]
    return traj_s, np.array(u_c)
