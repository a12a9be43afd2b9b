"""ORACLE (test infrastructure, NOT product code) -- float64 restatement of the reference's FORCESPRO formulation of the MPC
problem (/root/reference/MPC_Planner/optimizer.py:86-246), with the module interface oracle/ipm.py solves
(g_bounds, g_fun, g_jac, cost, cost_grad, lag_hess): `ipm.solve(d, w0, model=forces_nlp)`.

    stage variable   z_k = [deltaDot, aLong, xPos, yPos, delta, v, psi], k = 0 .. N-1            optimizer.py:93, 204-205
    parameters       p_k = [path_x, path_y, v_des, psi_ref, obstacle circle centres (6)]          optimizer.py:108-111, 313-317
    equalities       x_0 = xinit (model.xinitidx = 2..6, :224);  x_{k+1} = RK4(x_k, u_k; 0.1 s), k = 0 .. N-2   :90-98, 221
    inequalities     lb <= z_k <= ub (:108-109);  hl <= h(z_k, p_k) <= hu with h = [aLong^2 + (v psiDot)^2 ; nine SQUARED circle
                     distances] (:119-155), hl = [0, (r_ego + r_obs)^2 x 9], hu = [a_max^2, inf x 9] (:110-111)
    objective        sum_{k < N-1} f(z_k, p_k) + f_N(z_{N-1}, p_{N-1})                          :163-195, 217-218

PARITY STATUS: the reference hands this problem to the closed-source FORCESPRO SQP_NLP core (one QP per call, BFGS Hessian,
optimizer.py:226-240) whose iterate after one QP is not a mathematically defined quantity -- **parity unpinned at the solver
boundary**.  What IS pinned: every stage function below and its derivatives against the reference's CasADi-generated C
(test/FORCESNLPsolver/FORCESNLPsolver_model.c, tests/golden/forces_model_kat.npz + oracle/_ref/libforces_model.so) to 1e-10
(tests/test_forces_solver.py); the optimum of the NLP is cross-checked between oracle/ipm.py and scipy's trust-constr.

Two statements of the restated NLP differ from the literal model, neither changes its solution set:
  * the friction row's lower bound hl = 0 on a sum of squares is vacuous (and would put the interior-point slack ON its bound
    whenever aLong = psiDot = 0): the row is stated one-sided, h_0 <= a_max^2;
  * the last stage's controls u_{N-1} appear in no cost term and no dynamics (objectiveN has no input terms, there is no
    x_N): every feasible value is optimal.  The restatement adds R u_{N-1}^2 to select u_{N-1} = 0, the minimum-norm member.
  * stage 0's state bounds are dropped (x_0 is fixed by the initial-value equality, as FORCESPRO does with xinitidx).

Derivatives here are by COMPLEX-STEP differentiation (exact to rounding), deliberately not the hand-derived chain rule of
csrc/forces_model.cuh, so that agreement between the two is evidence.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .nlp import (VehicleParams, L_WB, L_FRICTION, compute_approximating_circle_radius,  # noqa: F401
                  compute_centers_of_approximation_circles)

NZ, NX, NU, NH, NPAR = 7, 5, 2, 10, 10
INF = np.inf


@dataclass
class ForcesData:
    N: int                      # model.N (optimizer.py:204): number of stages = states x_0 .. x_{N-1}
    dt: float                   # integrator step (optimizer.py:97; 0.1)
    Q: np.ndarray               # (5,) stage weights x, y, steering angle, velocity, heading          optimizer.py:172-176
    R: np.ndarray               # (2,) steering rate, acceleration                                     optimizer.py:177-178
    Pt: np.ndarray              # (5,) terminal weights                                                optimizer.py:191-195
    xinit: np.ndarray           # (5,)
    params: np.ndarray          # (N, 10)
    r_sum: float                # radius_ego + radius_obstacle (optimizer.py:110)
    ego_offset: float = 0.75
    l_wb: float = L_WB
    l_fric: float = L_FRICTION  # configuration.wheelbase (optimizer.py:131)
    veh: VehicleParams = field(default_factory=VehicleParams)
    # road-boundary rows the reference left commented out (optimizer.py:18-30, 113-117, 136-161; model.nh = 16): the distance of
    # each ego circle centre to the closest VERTEX of the left / right boundary polyline (ca.mmin over the vertices) >= radius_ego
    left_boundary: np.ndarray = None     # (nl, 2)  configuration.left_road_boundary (configuration.py:432)
    right_boundary: np.ndarray = None    # (nr, 2)  configuration.right_road_boundary (configuration.py:433)
    r_ego: float = 1.2

    @property
    def nh(self):
        return NH + (6 if self.left_boundary is not None else 0)

    @property
    def n(self):
        return NZ * self.N

    @property
    def m(self):
        return NX + NX * (self.N - 1) + self.nh * self.N


# ------------------------------------------------------------------ stage functions (complex-safe, vectorised over rows)
def ks_rhs(x, u, l_wb):
    """VehicleDynamics.KS_casadi (configuration.py:353-368)."""
    return np.stack([x[..., 3] * np.cos(x[..., 4]), x[..., 3] * np.sin(x[..., 4]), u[..., 0] + 0 * x[..., 0],
                     u[..., 1] + 0 * x[..., 0], x[..., 3] / l_wb * np.tan(x[..., 2])], axis=-1)


def dynamics(d, z):
    """forcespro.nlp.integrate(KS_casadi, z[2:7], z[0:2], RK4, stepsize) (optimizer.py:97-98): ONE classical RK4 step."""
    u, x = z[..., :2], z[..., 2:]
    h = d.dt
    k1 = ks_rhs(x, u, d.l_wb)
    k2 = ks_rhs(x + 0.5 * h * k1, u, d.l_wb)
    k3 = ks_rhs(x + 0.5 * h * k2, u, d.l_wb)
    k4 = ks_rhs(x + h * k3, u, d.l_wb)
    return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def inequalities(d, z, p):
    """circles_distance_inequality (optimizer.py:119-155)."""
    psid = z[..., 5] * np.tan(z[..., 4]) / d.l_fric
    out = [z[..., 1] ** 2 + (z[..., 5] * psid) ** 2]
    c, s = np.cos(z[..., 6]), np.sin(z[..., 6])
    for o in (0.0, d.ego_offset, -d.ego_offset):          # ego centre, front, rear (configuration.py:80-91)
        ex, ey = z[..., 2] + o * c, z[..., 3] + o * s
        for j in range(3):
            out.append((ex - p[..., 4 + 2 * j]) ** 2 + (ey - p[..., 5 + 2 * j]) ** 2)
    if d.left_boundary is not None:
        # find_closest_distance_with_road_boundary (optimizer.py:18-30): min over the boundary VERTICES of the distance; rows in
        # the reference's order: left boundary x (centre, front, rear), then right boundary (optimizer.py:156-161)
        for bnd in (d.left_boundary, d.right_boundary):
            for o in (0.0, d.ego_offset, -d.ego_offset):
                ex, ey = z[..., 2] + o * c, z[..., 3] + o * s
                dist = np.sqrt((ex[..., None] - bnd[:, 0]) ** 2 + (ey[..., None] - bnd[:, 1]) ** 2)
                idx = np.argmin(dist.real, axis=-1)
                out.append(np.take_along_axis(dist, idx[..., None], axis=-1)[..., 0])
    return np.stack(out, axis=-1)


def stage_cost(d, z, p, terminal):
    """cost_function / cost_functionN (optimizer.py:163-195); `terminal` is a boolean per row.  The terminal rows carry the
    minimum-norm selector R u^2 (module docstring)."""
    e = np.stack([z[..., 2] - p[..., 0], z[..., 3] - p[..., 1], z[..., 4], z[..., 5] - p[..., 2], z[..., 6] - p[..., 3]], axis=-1)
    w = np.where(np.asarray(terminal)[..., None], d.Pt, d.Q)
    return np.sum(w * e * e, axis=-1) + d.R[0] * z[..., 0] ** 2 + d.R[1] * z[..., 1] ** 2


def _cjac(fun, z, eps=1e-30):
    """Jacobian of fun: (rows, 7) -> (rows, m) by complex step; returns (rows, m, 7)."""
    cols = []
    for j in range(NZ):
        zc = z.astype(complex)
        zc[..., j] += 1j * eps
        cols.append(np.imag(fun(zc)) / eps)
    return np.stack(cols, axis=-1)


def stage_eval(d, z, p):
    """All first-order stage quantities at rows (z, p): dict(c, dc, h, dh, f, df, fN, dfN) -- the layout of the generated C model."""
    z = np.asarray(z, float); p = np.asarray(p, float)
    fl = np.zeros(z.shape[:-1], bool)
    return dict(c=dynamics(d, z), dc=_cjac(lambda q: dynamics(d, q), z), h=inequalities(d, z, p),
                dh=_cjac(lambda q: inequalities(d, q, p), z),
                f=stage_cost(d, z, p, fl) , df=_cjac(lambda q: stage_cost(d, q, p, fl)[..., None], z)[..., 0, :],
                fN=stage_cost(d, z, p, ~fl) - d.R[0] * z[..., 0] ** 2 - d.R[1] * z[..., 1] ** 2,
                dfN=_cjac(lambda q: (stage_cost(d, q, p, ~fl) - d.R[0] * q[..., 0] ** 2 - d.R[1] * q[..., 1] ** 2)[..., None], z)[..., 0, :])


# ------------------------------------------------------------------ the NLP in oracle/ipm.py's interface
def split(d, w):
    return np.asarray(w).reshape(d.N, NZ)


def _terminal(d):
    t = np.zeros(d.N, bool)
    t[-1] = True
    return t


def cost(d, w):
    return float(np.sum(stage_cost(d, split(d, w), d.params, _terminal(d))))


def cost_grad(d, w):
    Z = split(d, w)
    t = _terminal(d)
    return _cjac(lambda q: stage_cost(d, q, d.params, t)[..., None], Z)[:, 0, :].reshape(-1)


def g_fun(d, w):
    Z = split(d, w)
    return np.concatenate([Z[0, 2:] - d.xinit, (Z[1:, 2:] - dynamics(d, Z[:-1])).reshape(-1),
                           inequalities(d, Z, d.params).reshape(-1)])


def g_bounds(d):
    N, v = d.N, d.veh
    nb = d.nh - NH
    hl = np.concatenate([[-INF], np.full(9, d.r_sum ** 2), np.full(nb, d.r_ego)])          # optimizer.py:115-116 (commented): radius_ego
    hu = np.concatenate([[v.a_max ** 2], np.full(9 + nb, INF)])
    lbg = np.concatenate([np.zeros(NX * N), np.tile(hl, N)])
    ubg = np.concatenate([np.zeros(NX * N), np.tile(hu, N)])
    lb = np.array([v.deltav_min, -v.a_max, -INF, -INF, v.delta_min, v.v_min, -INF])       # optimizer.py:108
    ub = np.array([v.deltav_max, v.a_max, INF, INF, v.delta_max, v.v_max, INF])           # optimizer.py:109
    lbx, ubx = np.tile(lb, N), np.tile(ub, N)
    lbx[2:7] = -INF
    ubx[2:7] = INF                                                                        # x_0 is fixed by xinit
    return lbg, ubg, lbx, ubx


def g_jac(d, w):
    N = d.N
    Z = split(d, w)
    rows, cols, vals = [], [], []
    for j in range(NX):                                   # x_0 - xinit
        rows.append(j); cols.append(2 + j); vals.append(1.0)
    dc = _cjac(lambda q: dynamics(d, q), Z[:-1])          # (N-1, 5, 7)
    for k in range(N - 1):
        r0 = NX + NX * k
        for i in range(NX):
            rows.append(r0 + i); cols.append(NZ * (k + 1) + 2 + i); vals.append(1.0)
            for j in range(NZ):
                if dc[k, i, j] != 0.0:
                    rows.append(r0 + i); cols.append(NZ * k + j); vals.append(-dc[k, i, j])
    dh = _cjac(lambda q: inequalities(d, q, d.params), Z)  # (N, nh, 7)
    r1 = NX * N
    for k in range(N):
        for i in range(d.nh):
            for j in range(NZ):
                if dh[k, i, j] != 0.0:
                    rows.append(r1 + d.nh * k + i); cols.append(NZ * k + j); vals.append(dh[k, i, j])
    return sp.csr_matrix((vals, (rows, cols)), shape=(d.m, d.n))


def lag_hess(d, w, lam, sigma=1.0):
    """Hessian of sigma f + lam^T g: block diagonal over the stages (every term depends on one z_k only, x_{k+1} enters the
    defect rows linearly).  Central differences (step 1e-5) of the complex-step gradient of the stage Lagrangian."""
    N = d.N
    Z = split(d, w)
    t = _terminal(d)
    lam_dyn = np.zeros((N, NX))
    lam_dyn[:-1] = lam[NX:NX * N].reshape(N - 1, NX)
    lam_h = lam[NX * N:].reshape(N, d.nh)

    def stage_lag(q):
        val = sigma * stage_cost(d, q, d.params, t) + np.sum(lam_h * inequalities(d, q, d.params), axis=-1)
        return (val - np.sum(lam_dyn * dynamics(d, q), axis=-1))[..., None]

    grad = lambda q: _cjac(stage_lag, q)[:, 0, :]          # noqa: E731  (N, 7)
    H = np.zeros((N, NZ, NZ))
    step = 1e-5
    for j in range(NZ):
        e = np.zeros(NZ); e[j] = step * max(1.0, 1.0)
        H[:, :, j] = (grad(Z + e) - grad(Z - e)) / (2 * step)
    H = 0.5 * (H + np.transpose(H, (0, 2, 1)))
    return sp.block_diag([H[k] for k in range(N)], format="csr")


# ------------------------------------------------------------------ problem construction (ForcesproOptimizer.optimize, optimizer.py:248-323)
def velocity_profile(iter_length, N, desired_velocity):
    """desired velocity for every closed-loop step: constant, then a linear ramp to 0 over the last N steps (optimizer.py:291-294)."""
    return np.hstack((np.ones(iter_length - N) * desired_velocity, np.linspace(desired_velocity, 0, N)))


def stage_parameters(k, N, path, orientation, vel_all, obstacle_centers):
    """all_parameters of MPC step k as (N, 10) rows (optimizer.py:288-317): path points / orientations k+1 .. k+N, replenished with the
    last point, the velocity profile likewise, and the three obstacle circle centres tiled."""
    T = len(path)
    idx = np.minimum(np.arange(k + 1, k + 1 + N), T - 1)
    vi = np.minimum(np.arange(k + 1, k + 1 + N), len(vel_all) - 1)
    oc = np.asarray(obstacle_centers, float).reshape(-1)
    return np.column_stack([np.asarray(path, float)[idx, 0], np.asarray(path, float)[idx, 1], np.asarray(vel_all, float)[vi],
                            np.asarray(orientation, float)[idx], np.tile(oc, (N, 1))])


def make_nlp(N, dt, weights, xinit, params, static_obstacle, veh=None, wheelbase=L_FRICTION, road_boundaries=None):
    veh = veh or VehicleParams()
    Q = np.array([weights["weight_x"], weights["weight_y"], weights["weight_steering_angle"], weights["weight_velocity"],
                  weights["weight_heading_angle"]], float)
    R = np.array([weights["weight_velocity_steering_angle"], weights["weight_long_acceleration"]], float)
    Pt = np.array([weights["weight_x_terminate"], weights["weight_y_terminate"], weights["weight_steering_angle_terminate"],
                   weights["weight_velocity_terminate"], weights["weight_heading_angle_terminate"]], float)
    r_obs, _ = compute_approximating_circle_radius(static_obstacle["length"], static_obstacle["width"])
    r_ego, dd = compute_approximating_circle_radius(veh.length, veh.width)
    return ForcesData(N=N, dt=dt, Q=Q, R=R, Pt=Pt, xinit=np.asarray(xinit, float).reshape(NX),
                      params=np.asarray(params, float).reshape(N, NPAR), r_sum=r_ego + r_obs, ego_offset=dd / 4.0,
                      l_fric=wheelbase, veh=veh, r_ego=r_ego,
                      left_boundary=None if road_boundaries is None else np.asarray(road_boundaries[0], float),
                      right_boundary=None if road_boundaries is None else np.asarray(road_boundaries[1], float))


def initial_guess(d, a0=0.0):
    """problem['x0']: [0, init_acceleration, xinit] tiled over the stages (optimizer.py:267-270)."""
    return np.tile(np.concatenate([[0.0, a0], d.xinit]), d.N)
