"""ORACLE (test infrastructure, NOT product code) -- float64 restatement of the reference NLP.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (`mpc_b200`, libmpcb200.so) never does.

PARITY STATUS: *parity unpinned at the IPOPT boundary* -- casadi/IPOPT is an un-vendored third-party
dependency (casadi>=3.5.1, /root/reference/requirments:5) that is absent from this image and the reference
holds no numeric assertion for the NLP optimum.  What IS pinned (tests/test_oracle_golden.py):
the plant model and Euler/RK4 steps against 254 recorded transitions (error 0.0), the circle geometry and the
dynamics Jacobians against the CasADi-generated C in test/FORCESNLPsolver/FORCESNLPsolver_model.c (compiled into
oracle/_ref by oracle/Makefile), the exact step-0 optimum a0* = -sqrt(11.5), independent scipy SLSQP (N = 6) and trust-constr (N = 30 / 50, agreement 1e-6) solves,
and the scenario inputs themselves (every recorded RMSD / deviation file of the reference reproduced to rounding,
tests/test_results_format.py).  oracle/casadi_ref.py holds the verbatim casadi/IPOPT branch for boxes where casadi imports.
Statistically pinned in addition: every recorded step of the reference's own CasADi/IPOPT closed loops on ZAM_Over-1_1 and USA_Lanker-2_18_T-1 (N = 10,
applied control = optimum + N(0, sigma^2)) re-solved here leaves residuals with zero median and spread sigma
(tests/test_oracle_golden.py::test_recorded_ipopt_controls_pin_the_oracle_optimum_statistically).

Every function cites the reference lines it follows (paths relative to /root/reference/).

Decision vector / parameter / constraint ORDER is the reference's:
    w = [vec(U) (2N, stage-major: dd_0, a_0, dd_1, a_1, ...) ; vec(X) (5(N+1), stage-major)]   optimizer.py:550
    p = [vec(U_ref) ; vec(X_ref)]                                                             optimizer.py:552
    g = [friction(1) ; X_0 - X_ref_0 (5) ; defects (5N) ; obstacle distances 9(N+1)]          optimizer.py:378-403
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

NX, NU = 5, 2
L_WB = 2.5789128        # p.a + p.b of parameters_vehicle2 (configuration.py:362-363; FORCESNLPsolver_model.c:334)
L_FRICTION = 2.578      # literal in the friction row (optimizer.py:378)
INF = np.inf


# ------------------------------------------------------------------ geometry helpers (configuration.py:40-93)
def compute_approximating_circle_radius(length, width):
    """configuration.py:40-66 -- radius rounded UP to 0.1 m, centre distance round(2*l/3, 1)."""
    assert length >= 0 and width >= 0
    if np.isclose(length, 0.0) and np.isclose(width, 0.0):
        return 0.0, 0.0
    square_length = length / 3
    diagonal_square = np.sqrt((square_length / 2) ** 2 + (width / 2) ** 2)
    if diagonal_square > round(diagonal_square, 1):
        approx_radius = round(diagonal_square, 1) + 0.1
    else:
        approx_radius = round(diagonal_square, 1)
    return approx_radius, round(square_length * 2, 1)


def compute_centers_of_approximation_circles(x, y, length, width, orientation):
    """configuration.py:69-93 -- centre, front (+d/4... i.e. disc_distance/4 along heading), rear."""
    _, disc_distance = compute_approximating_circle_radius(length, width)
    distance_centers = disc_distance / 2
    off = distance_centers / 2
    c, s = np.cos(orientation), np.sin(orientation)
    return [x, y], [x + off * c, y + off * s], [x - off * c, y - off * s]


def ks_dynamics(x, u):
    """VehicleDynamics.KS_casadi, configuration.py:353-368.  x=[sx,sy,delta,v,psi], u=[delta_dot, a]."""
    x = np.asarray(x, float)
    u = np.asarray(u, float)
    return np.stack([x[..., 3] * np.cos(x[..., 4]),
                     x[..., 3] * np.sin(x[..., 4]),
                     u[..., 0] + 0 * x[..., 0],
                     u[..., 1] + 0 * x[..., 0],
                     x[..., 3] / L_WB * np.tan(x[..., 2])], axis=-1)


def euler_step(x, u, dt):
    """shift_movement plant update, optimizer.py:649-650."""
    return np.asarray(x, float) + dt * ks_dynamics(x, u)


def rk4_step(x, u, dt):
    """Forcespro plant / model.eq, optimizer.py:90-98 (single RK4 step of size dt)."""
    x = np.asarray(x, float)
    k1 = ks_dynamics(x, u)
    k2 = ks_dynamics(x + 0.5 * dt * k1, u)
    k3 = ks_dynamics(x + 0.5 * dt * k2, u)
    k4 = ks_dynamics(x + dt * k3, u)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


# ------------------------------------------------------------------ problem data
@dataclass
class VehicleParams:
    """Subset of vehiclemodels.parameters_vehicle2 the optimizer reads (optimizer.py:37-46, 68)."""
    delta_min: float = -1.066
    delta_max: float = 1.066
    deltav_min: float = -0.4
    deltav_max: float = 0.4
    v_min: float = 0.0          # optimizer.py:43 (hard-coded)
    v_max: float = 50.8
    a_max: float = 11.5
    length: float = 4.508
    width: float = 1.610


@dataclass
class NLPData:
    """Everything one call `sol(x0=..., p=..., lbg, ubg, lbx, ubx)` (optimizer.py:607) depends on."""
    N: int
    dt: float
    Q: np.ndarray               # (5,) diag, optimizer.py:500-503
    R: np.ndarray               # (2,) diag, optimizer.py:504
    xref: np.ndarray            # (N+1, 5); row 0 = current state (optimizer.py:667-699)
    obstacle_centers: np.ndarray  # (3,2): centre, front, rear (optimizer.py:60-64)
    r_sum: float                # radius_ego + radius_obstacle (optimizer.py:439)
    ego_offset: float = 0.75    # disc_distance/4 of the ego (configuration.py:80-91)
    veh: VehicleParams = field(default_factory=VehicleParams)
    # Friction row (optimizer.py:378, 424-425).  The reference writes sqrt(q^2) in [0, a_max] with
    # q = a_0^2 + v_0^2 tan(delta_0)/2.578.  That is the same feasible set as -a_max <= q <= a_max, but |q| has a kink at
    # q = 0 (reachable when delta_0 < 0) where its derivative is 0/0 and where an interior-point method jams because
    # the slack of the (never binding) lower bound |q| >= 0 goes to zero.  friction_smooth=True states the row as
    # q in [-a_max, a_max] (default, used for parity); False keeps the verbatim |q| in [0, a_max].
    friction_smooth: bool = True

    @property
    def n(self):
        return NU * self.N + NX * (self.N + 1)

    @property
    def m(self):
        return 1 + NX * (self.N + 1) + 9 * (self.N + 1)


def make_nlp(N, dt, weights, xref, static_obstacle, veh=None):
    """Build NLPData from the reference's configuration fields (Optimizer.__init__, optimizer.py:34-68)."""
    veh = veh or VehicleParams()
    Q = np.array([weights["weight_x"], weights["weight_y"], weights["weight_steering_angle"],
                  weights["weight_velocity"], weights["weight_heading_angle"]], float)
    R = np.array([weights["weight_velocity_steering_angle"], weights["weight_long_acceleration"]], float)
    oc = compute_centers_of_approximation_circles(static_obstacle["position_x"], static_obstacle["position_y"],
                                                  static_obstacle["length"], static_obstacle["width"],
                                                  static_obstacle["orientation"])
    r_obs, _ = compute_approximating_circle_radius(static_obstacle["length"], static_obstacle["width"])
    r_ego, dd = compute_approximating_circle_radius(veh.length, veh.width)
    return NLPData(N=N, dt=dt, Q=Q, R=R, xref=np.asarray(xref, float).reshape(N + 1, NX),
                   obstacle_centers=np.array(oc, float), r_sum=r_ego + r_obs, ego_offset=dd / 4.0, veh=veh)


# ------------------------------------------------------------------ unpack helpers
def split(w, N):
    U = np.asarray(w[:NU * N]).reshape(N, NU)
    X = np.asarray(w[NU * N:]).reshape(N + 1, NX)
    return U, X


def pack(U, X):
    return np.concatenate([np.asarray(U, float).reshape(-1), np.asarray(X, float).reshape(-1)])


def iu(N, k, j):
    return NU * k + j


def ix(N, k, j):
    return NU * N + NX * k + j


# ------------------------------------------------------------------ objective (optimizer.py:493-511, quirks Q1, Q2)
def cost(d, w):
    U, X = split(w, d.N)
    e = X[:d.N] - d.xref[1:d.N + 1]          # X[:, i] pairs with X_ref[:, i+1]  (optimizer.py:509)
    return float(np.sum(e * e * d.Q) + np.sum(U * U * d.R))   # terminal term is dead code (optimizer.py:510)


def cost_grad(d, w):
    U, X = split(w, d.N)
    gU = 2 * U * d.R
    gX = np.zeros_like(X)
    gX[:d.N] = 2 * (X[:d.N] - d.xref[1:d.N + 1]) * d.Q
    return pack(gU, gX)


def cost_hess_diag(d):
    hU = np.tile(2 * d.R, d.N)
    hX = np.concatenate([np.tile(2 * d.Q, d.N), np.zeros(NX)])
    return np.concatenate([hU, hX])


# ------------------------------------------------------------------ constraints g (optimizer.py:373-411)
_SIG = np.array([0.0, 1.0, -1.0])   # centre, front (+), rear (-)  (configuration.py:80-91)


def _obst(d, X):
    """distances (N+1, 3) and the pieces needed for derivatives."""
    c, s = np.cos(X[:, 4]), np.sin(X[:, 4])
    off = d.ego_offset * _SIG[None, :]
    dx = X[:, 0:1] + off * c[:, None] - d.obstacle_centers[None, :, 0]
    dy = X[:, 1:2] + off * s[:, None] - d.obstacle_centers[None, :, 1]
    h = np.sqrt(dx * dx + dy * dy)
    return h, dx, dy, c, s, off


def g_fun(d, w):
    N = d.N
    U, X = split(w, N)
    q = U[0, 1] ** 2 + X[0, 3] * (np.tan(X[0, 2]) * X[0, 3] / L_FRICTION)     # optimizer.py:378 (Q3)
    out = [np.array([q if d.friction_smooth else abs(q)])]                     # sqrt(sq(q)) == |q|
    out.append(X[0] - d.xref[0])                                               # optimizer.py:378, 2nd item
    xn = X[:N] + d.dt * ks_dynamics(X[:N], U)                                   # optimizer.py:380-382
    out.append((X[1:] - xn).reshape(-1))
    h = _obst(d, X)[0]                                                         # optimizer.py:384-403 (Q6)
    out.append(np.repeat(h, 3, axis=1).reshape(-1))                            # each distance listed 3x
    return np.concatenate(out)


def g_bounds(d):
    """inequal_constraints, optimizer.py:413-491 -> lbg, ubg, lbx, ubx as arrays."""
    N, v = d.N, d.veh
    lbg = np.concatenate([[-v.a_max if d.friction_smooth else 0.0], np.zeros(NX * (N + 1)), np.full(9 * (N + 1), d.r_sum)])
    ubg = np.concatenate([[v.a_max], np.zeros(NX * (N + 1)), np.full(9 * (N + 1), INF)])
    lbx = np.concatenate([np.tile([v.deltav_min, -INF], N), np.tile([-INF, -INF, v.delta_min, v.v_min, -INF], N + 1)])
    ubx = np.concatenate([np.tile([v.deltav_max, v.a_max], N), np.tile([INF, INF, v.delta_max, v.v_max, INF], N + 1)])
    return lbg, ubg, lbx, ubx


_STRUCT_CACHE = {}


def _structure(N):
    """Row/column index patterns of the Jacobian and the Lagrangian Hessian (depend on N only)."""
    if N in _STRUCT_CACHE:
        return _STRUCT_CACHE[N]
    k = np.arange(N)
    kk = np.arange(N + 1)
    X = lambda st, j: NU * N + NX * st + j      # noqa: E731
    Uc = lambda st, j: NU * st + j               # noqa: E731
    jr, jc = [], []
    # friction row (3 entries): a0, delta0, v0
    jr += [np.zeros(3, int)]
    jc += [np.array([Uc(0, 1), X(0, 2), X(0, 3)])]
    # pin rows
    jr += [1 + np.arange(NX)]
    jc += [X(0, np.arange(NX))]
    r0 = 1 + NX
    # defects: +I on X_{k+1}, -I on X_k  (10 per stage), then 8 structural entries per stage
    for j in range(NX):
        jr += [r0 + NX * k + j, r0 + NX * k + j]
        jc += [X(k + 1, j), X(k, j)]
    for (rj, cj) in ((0, ("x", 3)), (0, ("x", 4)), (1, ("x", 3)), (1, ("x", 4)), (2, ("u", 0)), (3, ("u", 1)),
                     (4, ("x", 2)), (4, ("x", 3))):
        jr += [r0 + NX * k + rj]
        jc += [X(k, cj[1]) if cj[0] == "x" else Uc(k, cj[1])]
    # obstacle rows: for each stage, circle j, copy rep: entries on sx, sy, psi
    r1 = 1 + NX * (N + 1)
    for j in range(3):
        for rep in range(3):
            for col in (0, 1, 4):
                jr += [r1 + 9 * kk + 3 * j + rep]
                jc += [X(kk, col)]
    jr = np.concatenate([np.atleast_1d(a) for a in jr])
    jc = np.concatenate([np.atleast_1d(a) for a in jc])
    # Hessian pattern
    hr, hc = [], []
    n = NU * N + NX * (N + 1)
    hr += [np.arange(n)]
    hc += [np.arange(n)]                                     # cost diagonal
    ia, idl, iv = Uc(0, 1), X(0, 2), X(0, 3)
    hr += [np.array([ia, iv, idl, idl, iv])]
    hc += [np.array([ia, iv, idl, iv, idl])]                 # friction
    hr += [X(k, 4), X(k, 3), X(k, 4), X(k, 2), X(k, 2), X(k, 3)]
    hc += [X(k, 4), X(k, 4), X(k, 3), X(k, 2), X(k, 3), X(k, 2)]  # dynamics
    cols = (0, 1, 4)
    for a in range(3):
        for b in range(3):
            hr += [X(kk, cols[a])]
            hc += [X(kk, cols[b])]                           # obstacle 3x3 block per stage (summed over circles)
    hr = np.concatenate([np.atleast_1d(a) for a in hr])
    hc = np.concatenate([np.atleast_1d(a) for a in hc])
    _STRUCT_CACHE[N] = (jr, jc, hr, hc)
    return _STRUCT_CACHE[N]


def g_jac(d, w):
    """Sparse Jacobian of g (m x n), analytic, vectorised over stages."""
    N, dt = d.N, d.dt
    U, X = split(w, N)
    jr, jc, _, _ = _structure(N)
    a0, de0, v0 = U[0, 1], X[0, 2], X[0, 3]
    t0 = np.tan(de0)
    q = a0 ** 2 + v0 * v0 * t0 / L_FRICTION
    sg = 1.0 if d.friction_smooth else np.sign(q)
    vals = [np.array([sg * 2 * a0, sg * v0 * v0 * (1 + t0 * t0) / L_FRICTION, sg * 2 * v0 * t0 / L_FRICTION]),
            np.ones(NX)]
    one = np.ones(N)
    for j in range(NX):
        vals += [one, -one]
    de, vv, ps = X[:N, 2], X[:N, 3], X[:N, 4]
    c, s, t = np.cos(ps), np.sin(ps), np.tan(de)
    vals += [-dt * c, dt * vv * s, -dt * s, -dt * vv * c, -dt * one, -dt * one,
             -dt * vv / L_WB * (1 + t * t), -dt * t / L_WB]
    h, dx, dy, c, s, off = _obst(d, X)
    for j in range(3):
        hx, hy = dx[:, j] / h[:, j], dy[:, j] / h[:, j]
        hp = (dx[:, j] * (-off[0, j] * s) + dy[:, j] * (off[0, j] * c)) / h[:, j]
        for rep in range(3):
            vals += [hx, hy, hp]
    vals = np.concatenate(vals)
    return sp.csr_matrix((vals, (jr, jc)), shape=(d.m, d.n))


def lag_hess(d, w, lam, sigma=1.0):
    """Hessian of sigma*f + lam^T g (n x n, symmetric, sparse) -- what CasADi's AD hands IPOPT (optimizer.py:558)."""
    N, dt = d.N, d.dt
    U, X = split(w, N)
    _, _, hr, hc = _structure(N)
    vals = [sigma * cost_hess_diag(d)]
    # friction row: lam0 * sign(q) * hess q
    a0, de0, v0 = U[0, 1], X[0, 2], X[0, 3]
    t0 = np.tan(de0)
    sec2 = 1 + t0 * t0
    q = a0 ** 2 + v0 * v0 * t0 / L_FRICTION
    l0 = lam[0] * (1.0 if d.friction_smooth else np.sign(q))
    cr = l0 * 2 * v0 * sec2 / L_FRICTION
    vals += [np.array([l0 * 2.0, l0 * 2 * t0 / L_FRICTION, l0 * 2 * v0 * v0 * sec2 * t0 / L_FRICTION, cr, cr])]
    # defect rows d_k = X_{k+1} - X_k - dt f(X_k,U_k): contribution -dt * lam_j * hess f_j(X_k)
    r0 = 1 + NX
    l = lam[r0: r0 + NX * N].reshape(N, NX)
    de, vv, ps = X[:N, 2], X[:N, 3], X[:N, 4]
    c, s, t = np.cos(ps), np.sin(ps), np.tan(de)
    sec2 = 1 + t * t
    hpp = -dt * (l[:, 0] * (-vv * c) + l[:, 1] * (-vv * s))
    hvp = -dt * (l[:, 0] * (-s) + l[:, 1] * c)
    hdd = -dt * l[:, 4] * 2 * vv / L_WB * sec2 * t
    hdv = -dt * l[:, 4] * sec2 / L_WB
    vals += [hpp, hvp, hvp, hdd, hdv, hdv]
    # obstacle rows: h = |r|, r = (dx,dy);  hess h = (D^T D + r.d2r - grad grad^T) / h
    h, dx, dy, c, s, off = _obst(d, X)
    r1 = 1 + NX * (N + 1)
    lo = lam[r1:].reshape(N + 1, 3, 3).sum(axis=2)            # the 3 copies share one distance
    blk = np.zeros((N + 1, 3, 3))
    for j in range(3):
        o = off[0, j]
        D = np.zeros((N + 1, 2, 3))
        D[:, 0, 0] = 1.0
        D[:, 1, 1] = 1.0
        D[:, 0, 2] = -o * s
        D[:, 1, 2] = o * c
        r = np.stack([dx[:, j], dy[:, j]], axis=1)
        hh = h[:, j]
        rD = np.einsum("ki,kij->kj", r, D)
        M = np.einsum("kia,kib->kab", D, D)
        M[:, 2, 2] += r[:, 0] * (-o * c) + r[:, 1] * (-o * s)
        Hj = (M - np.einsum("ka,kb->kab", rD, rD) / (hh * hh)[:, None, None]) / hh[:, None, None]
        blk += lo[:, j][:, None, None] * Hj
    for a in range(3):
        for b in range(3):
            vals += [blk[:, a, b]]
    vals = np.concatenate(vals)
    return sp.csr_matrix((vals, (hr, hc)), shape=(d.n, d.n))


# ------------------------------------------------------------------ reference-window rule (optimizer.py:657-702, Q8)
def reference_window(i, x_now, N, iter_length, path, orientation, desired_velocity):
    """desired_command_and_trajectory(i, x0_, N_) -> X_ref (N+1, 5).  Row 0 = current state."""
    rows = [np.asarray(x_now, float).reshape(NX)]
    for k in range(N):
        if i >= iter_length - N:
            j = i + k + 1 - (i - (iter_length - N) + 1)      # optimizer.py:672 == k + iter_length - N
        else:
            j = i + k + 1                                     # optimizer.py:686
        rows.append(np.array([path[j, 0], path[j, 1], 0.0, desired_velocity, orientation[j]]))
    return np.stack(rows)
