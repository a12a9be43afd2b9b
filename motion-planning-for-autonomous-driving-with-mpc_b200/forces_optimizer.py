"""Host-side mirror of the reference's `ForcesproOptimizer` (/root/reference/MPC_Planner/optimizer.py:86-369) on top of
`mpcb200_forces_solve`: same constructor, same `optimize()` contract `(x[T,5], u[T,2], solve_time[T])`, plus the batched calls.

Selection without touching the reference: `MPC_Planner.mpc_planner.ForcesproOptimizer = B200ForcesproOptimizer` with
`framework_name: forcespro` (mpc_planner.py:304-309), see INTEGRATION.md.

What differs from the reference's solver, by construction: FORCESPRO's SQP_NLP core runs ONE QP per call with a BFGS Hessian
(optimizer.py:226-240) -- its output is an intermediate iterate of a closed-source code; this class returns the converged
optimum of the same NLP (checked against oracle/forces_nlp.py + oracle/ipm.py in tests/test_forces_solver.py).
"""
import ctypes
import time

import numpy as np

from .optimizer import B200Optimizer

_TERMINAL_KEYS = ("weight_x_terminate", "weight_y_terminate", "weight_steering_angle_terminate", "weight_velocity_terminate",
                  "weight_heading_angle_terminate")


def velocity_profile(iter_length, N, desired_velocity):
    """optimizer.py:291-294: constant desired velocity, then a linear ramp to 0 over the last N closed-loop steps."""
    return np.hstack((np.ones(iter_length - N) * desired_velocity, np.linspace(desired_velocity, 0, N)))


def stage_parameters(k, N, path, orientation, vel_all, obstacle_centers):
    """`all_parameters` of MPC step k as [N,10] rows (optimizer.py:288-317): path points / headings k+1 .. k+N replenished with the
    last one, the velocity profile likewise, the three obstacle circle centres tiled over the stages."""
    path = np.asarray(path, float)
    idx = np.minimum(np.arange(k + 1, k + 1 + N), len(path) - 1)
    vi = np.minimum(np.arange(k + 1, k + 1 + N), len(vel_all) - 1)
    oc = np.asarray(obstacle_centers, float).reshape(-1)
    return np.column_stack([path[idx, 0], path[idx, 1], np.asarray(vel_all, float)[vi], np.asarray(orientation, float)[idx],
                            np.tile(oc, (N, 1))])


class B200ForcesproOptimizer(B200Optimizer):
    """The FORCESPRO formulation (RK4 dynamics, friction circle at every stage, 3 x 3 circle pairs, terminal weights, symmetric
    acceleration bounds, desired-velocity ramp) solved per ego instance on one B200.  `predict_horizon` is model.N."""

    def __init__(self, configuration, init_values, predict_horizon, **kw):
        # optimizer.py:131: the friction row uses configuration.wheelbase
        kw.setdefault("l_fric", float(getattr(configuration, "wheelbase", 2.578)))
        # the dynamics-curvature term of the Lagrangian Hessian is on by default: pure Gauss-Newton diverges on weight sets like
        # USA_Lanker's (forces_core.cuh, backward_t); hessian="gn" selects the cheaper sweep where it is known to contract
        kw.setdefault("hessian", "exact")
        super(B200ForcesproOptimizer, self).__init__(configuration, init_values, predict_horizon, **kw)
        w = self.weights_setting
        self.weights_terminal = np.array([float(w[k]) for k in _TERMINAL_KEYS])
        self._wt = (ctypes.c_double * 5)(*self.weights_terminal)
        self.vel_all = velocity_profile(int(self.iter_length), self.N, float(self.desired_velocity)) if self.iter_length >= self.N else \
            np.linspace(float(self.desired_velocity), 0, int(self.iter_length))

    def set_road_boundaries(self, left=None, right=None, r_min=None):
        """Switches the six road-boundary rows per stage on (the reference states them and leaves them commented out,
        optimizer.py:18-30, 113-117, 136-161): each ego circle centre keeps at least `r_min` (default radius_ego, :115) from the
        closest vertex of either boundary polyline.  left / right: [n,2] arrays (default: configuration.left_road_boundary /
        right_road_boundary, configuration.py:432-433); `clear_road_boundaries()` switches the rows off."""
        left = getattr(self.configuration, "left_road_boundary", None) if left is None else left
        right = getattr(self.configuration, "right_road_boundary", None) if right is None else right
        h = self.handle
        if left is None or right is None:
            h.check(h.lib.mpcb200_forces_set_road_boundaries(h.h, None, 0, None, 0, 0.0))
            return
        left = np.ascontiguousarray(left, np.float64).reshape(-1, 2)
        right = np.ascontiguousarray(right, np.float64).reshape(-1, 2)
        r_min = float(self.radius_ego if r_min is None else r_min)
        h.check(h.lib.mpcb200_forces_set_road_boundaries(h.h, left.ctypes.data, len(left), right.ctypes.data, len(right), r_min))

    def clear_road_boundaries(self):
        """Road-boundary rows off again (the default, as in the reference where they are commented out)."""
        h = self.handle
        h.check(h.lib.mpcb200_forces_set_road_boundaries(h.h, None, 0, None, 0, 0.0))

    def _obstacle_within_reach(self):
        return True            # the friction circle is a nonlinear row at every stage: keep the float64 pass for float32 stragglers

    # ------------------------------------------------------------------ batched API
    def params_for_step(self, k):
        return stage_parameters(k, self.N, np.asarray(self.resampled_path_points, float)[:, :2], self.orientation, self.vel_all,
                                self.obstacle_circles_centers_tuple)

    def forces_solve_batch(self, xinit, params, Z_init=None):
        """One NLP per row: xinit [B,5], params [B,N,10] (or [N,10], shared), optional initial guess Z_init [B,N,7].
        Returns (Z[B,N,7], status[B], iters[B]) as CUDA tensors; Z rows are [deltaDot, aLong, xPos, yPos, delta, v, psi]."""
        t = self.torch
        xinit = self._dev(xinit).reshape(-1, 5)
        B = xinit.shape[0]
        params = self._dev(params)
        if params.dim() == 2:
            params = params.unsqueeze(0).expand(B, self.N, 10).contiguous()
        assert params.shape == (B, self.N, 10), params.shape
        zin = None if Z_init is None else self._dev(Z_init)
        assert zin is None or zin.shape == (B, self.N, 7)
        Z = t.empty(B, self.N, 7, dtype=t.float64, device=self.device)
        status = t.empty(B, dtype=t.int32, device=self.device)
        iters = t.empty(B, dtype=t.int32, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_forces_solve(h.h, self._wt, xinit.data_ptr(), params.data_ptr(), None if zin is None else zin.data_ptr(),
                                           Z.data_ptr(), status.data_ptr(), iters.data_ptr(), B, self._stream()))
        return Z, status, iters

    def rk4_plant(self, x, u):
        """model.eq: one RK4 step of the kinematic single-track model (optimizer.py:90-98, 359), batched on the device."""
        t = self.torch
        l_wb, h = float(self.cfg.l_wb), float(self.cfg.dt)

        def f(xx):
            return t.stack([xx[:, 3] * t.cos(xx[:, 4]), xx[:, 3] * t.sin(xx[:, 4]), u[:, 0], u[:, 1], xx[:, 3] / l_wb * t.tan(xx[:, 2])], dim=1)
        k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
        return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    def optimize_batch(self, x0, warm_start=True, return_device=False, on_device=None):
        """ForcesproOptimizer.optimize()'s loop (optimizer.py:286-362) for B egos: per closed-loop step the parameters of the next N
        path points, one solve, the first input applied to the RK4 plant.  x0 [B,5] -> (states[B,T,5], inputs[B,T,2], status[B,T],
        iters[B,T]).  warm_start: the previous solution shifted one stage is the next initial guess (the reference re-uses its
        step-0 guess every step, optimizer.py:267-277; the converged optimum does not depend on it).
        on_device (default: True for noise-free runs): the whole loop in ONE launch (`mpcb200_forces_closed_loop`); otherwise one
        `mpcb200_forces_solve` per step driven from here (needed for `configuration.noised`)."""
        t = self.torch
        x = self._dev(x0).reshape(-1, 5).clone()
        B, T = x.shape[0], int(self.iter_length)
        noised = bool(getattr(self.configuration, "noised", False))
        if on_device is None:
            on_device = not noised and warm_start
        if on_device:
            if noised:
                raise ValueError("the device loop is noise-free; use on_device=False for configuration.noised")
            path, orient = self._path_tensors()
            vel = self._dev(np.asarray(self.vel_all, float)[np.minimum(np.arange(T), len(self.vel_all) - 1)])
            traj = t.empty(B, T, 5, dtype=t.float64, device=self.device)
            ctrl = t.empty(B, T, 2, dtype=t.float64, device=self.device)
            status = t.empty(B, T, dtype=t.int32, device=self.device)
            iters = t.empty(B, T, dtype=t.int32, device=self.device)
            h = self.handle
            h.check(h.lib.mpcb200_forces_closed_loop(h.h, self._wt, T, path.data_ptr(), orient.data_ptr(), vel.data_ptr(), x.data_ptr(),
                                                     traj.data_ptr(), ctrl.data_ptr(), status.data_ptr(), iters.data_ptr(), B, self._stream()))
            if return_device:
                return traj, ctrl, status, iters
            return traj.cpu().numpy(), ctrl.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()
        traj = t.empty(B, T, 5, dtype=t.float64, device=self.device)
        ctrl = t.empty(B, T, 2, dtype=t.float64, device=self.device)
        status = t.empty(B, T, dtype=t.int32, device=self.device)
        iters = t.empty(B, T, dtype=t.int32, device=self.device)
        Z = None
        for k in range(T):
            traj[:, k] = x
            zin = None
            if warm_start and Z is not None:
                zin = t.cat([Z[:, 1:], Z[:, -1:]], dim=1).contiguous()
            Z, st, it = self.forces_solve_batch(x, self.params_for_step(k), zin)
            u = Z[:, 0, :2].clone()
            if noised:                                                                   # optimizer.py:347-356
                std = 0.1 if getattr(self.configuration, "use_case", "lane_following") == "lane_following" else 0.05
                u = u + t.as_tensor(np.random.normal(0.0, std, (B, 2)), device=self.device)
            ctrl[:, k] = u
            status[:, k] = st
            iters[:, k] = it
            x = self.rk4_plant(x, u)
        if return_device:
            return traj, ctrl, status, iters
        return traj.cpu().numpy(), ctrl.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()

    # ------------------------------------------------------------------ the reference contract
    def optimize(self, on_device=None):
        """ForcesproOptimizer.optimize() (optimizer.py:248-369): returns (x[T,5], u[T,2], solve_time[T])."""
        x0 = np.array([self.init_position[0], self.init_position[1], 0.0, self.init_velocity, self.init_orientation], float)   # :280
        self.torch.cuda.synchronize(self.device)
        t0 = time.time()
        traj, ctrl, status, _ = self.optimize_batch(x0[None, :])
        dt = (time.time() - t0) / max(1, traj.shape[1])
        if not (status[0] >= 1).all():              # the reference asserts exitflag == 1 (optimizer.py:330)
            bad = np.where(status[0] < 1)[0]
            raise AssertionError(f"bad exitflag {status[0][bad[0]]} at step {bad[0]}")
        return traj[0], ctrl[0], np.full(traj.shape[1], dt)
