"""B200Optimizer -- drop-in for `CasadiOptimizer` behind the reference's `Optimizer` base class.

Reference interface mirrored (paths relative to /root/reference/):
    Optimizer.__init__(configuration, init_values, predict_horizon)      MPC_Planner/optimizer.py:34-68
    CasadiOptimizer.optimize() -> (states[T,5], controls[T,2], t[T])      MPC_Planner/optimizer.py:562-643
    call sites                                                             MPC_Planner/mpc_planner.py:302, 309

When `MPC_Planner.optimizer` is importable (casadi, forcespro, commonroad ... installed) `B200Optimizer` subclasses the
real `Optimizer`; otherwise it subclasses `OptimizerBase`, an attribute-compatible mirror.  Either way the numerical
work is done by libmpcb200.so (hand-written sm_100a CUDA) through ctypes; torch is used only for device buffers and
streams.  There is no CPU path: without CUDA + the built library this module raises.

Additive batched API (new): `solve_batch`, `optimize_batch`.
"""
import time

import numpy as np

from . import _capi
from .scenarios import reference_window

L_WB = 2.5789128      # parameters_vehicle2: p.a + p.b (configuration.py:362-363)
L_FRICTION = 2.578    # literal in the friction row (optimizer.py:378)


# ---------------------------------------------------------------------------------------------------------------
# geometry helpers restated for floats (configuration.py:40-93); the reference's versions call ca.cos/ca.sin
def compute_approximating_circle_radius(length, width):
    """configuration.py:40-66"""
    assert length >= 0 and width >= 0, 'Invalid vehicle dimensions = {}'.format([length, width])
    if np.isclose(length, 0.0) and np.isclose(width, 0.0):
        return 0.0, 0.0
    square_length = length / 3
    diagonal_square = np.sqrt((square_length / 2) ** 2 + (width / 2) ** 2)
    if diagonal_square > round(diagonal_square, 1):
        approx_radius = round(diagonal_square, 1) + 0.1
    else:
        approx_radius = round(diagonal_square, 1)
    return approx_radius, round(square_length * 2, 1)


def compute_centers_of_approximation_circles(x_position, y_position, v_length, v_width, orientation):
    """configuration.py:69-93"""
    _, disc_distance = compute_approximating_circle_radius(v_length, v_width)
    distance_centers = disc_distance / 2
    c, s = float(np.cos(orientation)), float(np.sin(orientation))
    center = [x_position, y_position]
    center_fw = [x_position + (distance_centers / 2) * c, y_position + (distance_centers / 2) * s]
    center_rw = [x_position - (distance_centers / 2) * c, y_position - (distance_centers / 2) * s]
    return center, center_fw, center_rw


class _Steering:
    min, max, v_min, v_max = -1.066, 1.066, -0.4, 0.4


class _Longitudinal:
    v_max, a_max = 50.8, 11.5


class VehicleParameters2:
    """The fields of vehiclemodels.parameters_vehicle2 (BMW 320i) the optimizer reads (optimizer.py:37-46, 68)."""
    steering = _Steering
    longitudinal = _Longitudinal
    l, w = 4.508, 1.610
    a, b = 1.1561957, 1.4227171   # a + b = 2.5789128


def obstacle_circles_and_radius(static_obstacle, p=VehicleParameters2):
    """Optimizer.__init__ lines 60-68: obstacle circle centres, r_ego + r_obs, ego circle offset."""
    circles = compute_centers_of_approximation_circles(static_obstacle["position_x"], static_obstacle["position_y"],
                                                       static_obstacle["length"], static_obstacle["width"],
                                                       static_obstacle["orientation"])
    r_obs, _ = compute_approximating_circle_radius(static_obstacle["length"], static_obstacle["width"])
    r_ego, dd = compute_approximating_circle_radius(p.l, p.w)
    return circles, r_ego + r_obs, dd / 4.0


class OptimizerBase(object):
    """Attribute-compatible mirror of `Optimizer` (optimizer.py:33-83) for environments without casadi et al."""

    def __init__(self, configuration, init_values, predict_horizon):
        self.configuration = configuration
        self.delta_min = configuration.p.steering.min
        self.delta_max = configuration.p.steering.max
        self.deltav_min = configuration.p.steering.v_min
        self.deltav_max = configuration.p.steering.v_max
        self.v_min = 0
        self.v_max = configuration.p.longitudinal.v_max
        self.a_max = configuration.p.longitudinal.a_max
        self.init_position, self.init_velocity, self.init_acceleration, self.init_orientation = \
            init_values[0], init_values[1], init_values[2], init_values[3]
        self.iter_length = configuration.iter_length
        self.delta_t = configuration.delta_t
        self.desired_velocity = configuration.desired_velocity
        self.resampled_path_points = configuration.reference_path
        self.orientation = configuration.orientation
        self.predict_horizon = predict_horizon
        self.weights_setting = configuration.weights_setting
        so = configuration.static_obstacle
        self.obstacle_circles_centers_tuple = compute_centers_of_approximation_circles(
            so["position_x"], so["position_y"], so["length"], so["width"], so["orientation"])
        self.radius_obstacle, _ = compute_approximating_circle_radius(so["length"], so["width"])
        self.radius_ego, _ = compute_approximating_circle_radius(configuration.p.l, configuration.p.w)

    def equal_constraints(self, *args, **kwargs):
        pass

    def inequal_constraints(self, *args, **kwargs):
        pass

    def cost_function(self, *args, **kwargs):
        pass

    def solver(self):
        pass

    def optimize(self):
        pass


try:  # the real base class when the reference and its dependencies are importable
    from MPC_Planner.optimizer import Optimizer as _RefOptimizer  # type: ignore
    _Base = _RefOptimizer
except Exception:  # casadi / forcespro / commonroad absent (this image)
    _Base = OptimizerBase


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise _capi.Mpcb200Error("B200Optimizer needs a CUDA device; there is no CPU fallback")
    return torch


class B200Optimizer(_Base):
    """Batched nonlinear-MPC optimizer on one B200.  `optimize()` keeps the reference's contract (B = 1)."""

    def __init__(self, configuration, init_values, predict_horizon, precision="f32", hessian="gn",
                 max_batch=4096, device=None, max_iter=100, **solver_opts):
        super(B200Optimizer, self).__init__(configuration, init_values, predict_horizon)
        torch = _require_cuda()
        self.torch = torch
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        N = int(predict_horizon)
        cfg = _capi.default_config(N, _capi.F64 if precision in ("f64", "float64", 1) else _capi.F32)
        cfg.device = self.device_index
        cfg.max_batch = int(max_batch)
        cfg.hessian = _capi.HESS_EXACT if hessian in ("exact", 1) else _capi.HESS_GAUSS_NEWTON
        cfg.max_iter = int(max_iter)
        cfg.dt = float(self.delta_t)
        cfg.l_wb = float(getattr(configuration.p, "a", VehicleParameters2.a) + getattr(configuration.p, "b", VehicleParameters2.b))
        cfg.l_fric = L_FRICTION
        w = self.weights_setting
        for i, k in enumerate(("weight_x", "weight_y", "weight_steering_angle", "weight_velocity", "weight_heading_angle")):
            cfg.Q[i] = float(w[k])
        cfg.R[0] = float(w["weight_velocity_steering_angle"])
        cfg.R[1] = float(w["weight_long_acceleration"])
        cfg.deltav_min, cfg.deltav_max = float(self.deltav_min), float(self.deltav_max)
        cfg.a_max = float(self.a_max)
        cfg.delta_min, cfg.delta_max = float(self.delta_min), float(self.delta_max)
        cfg.v_min, cfg.v_max = float(self.v_min), float(self.v_max)
        cfg.r_sum = float(self.radius_ego + self.radius_obstacle)
        _, dd = compute_approximating_circle_radius(configuration.p.l, configuration.p.w)
        cfg.ego_offset = dd / 4.0
        for j in range(3):
            cfg.obstacle[2 * j] = float(self.obstacle_circles_centers_tuple[j][0])
            cfg.obstacle[2 * j + 1] = float(self.obstacle_circles_centers_tuple[j][1])
        if cfg.precision == _capi.F32 and "refine_f64" not in solver_opts:
            cfg.refine_f64 = 1 if self._obstacle_within_reach() else 0
        _capi.set_options(cfg, solver_opts)            # raises on a name that is not a field of mpcb200_config
        self.cfg = cfg
        self.N = N
        with torch.cuda.device(self.device):
            self.handle = _capi.Handle(cfg)
        self._path_d = None

    # ------------------------------------------------------------------ helpers
    OBSTACLE_REACH_M = 25.0

    def _obstacle_within_reach(self):
        """The float64 refinement pass exists for ACTIVE obstacle rows (float32 stalls on their stiff barrier weights, status 3).
        A row can only become active where the ego can touch the obstacle: if every point of the reference path keeps more than
        OBSTACLE_REACH_M + r_ego + r_obs from every obstacle circle (lane following: the reference's dummy obstacle at (-100, 0),
        quirk Q11), the pass could never find work and its (empty) launch is skipped.  The library default is refine_f64 = 1;
        pass refine_f64=... to override this decision either way."""
        path = np.asarray(self.resampled_path_points, float)[:, :2]
        r = float(self.radius_ego + self.radius_obstacle) + self.OBSTACLE_REACH_M
        pts = np.vstack([path, np.array([[self.init_position[0], self.init_position[1]]], float)])
        for c in self.obstacle_circles_centers_tuple:
            if np.hypot(pts[:, 0] - float(c[0]), pts[:, 1] - float(c[1])).min() <= r:
                return True
        return False

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, a):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a.to(device=self.device, dtype=t.float64).contiguous()
        return t.as_tensor(np.ascontiguousarray(a, np.float64), device=self.device)

    def _path_tensors(self):
        if self._path_d is None:
            self._path_d = (self._dev(np.asarray(self.resampled_path_points, float)[:, :2]),
                            self._dev(np.asarray(self.orientation, float)))
        return self._path_d

    # ------------------------------------------------------------------ batched API (device tensors in / out)
    def solve_batch(self, xref, X_init=None, U_init=None, out=None):
        """One NLP solve per row (replaces optimizer.py:605-607).  xref [B,N+1,5] (row 0 = current state).
        Returns (U*[B,N,2], X*[B,N+1,5], status[B] int32, iters[B] int32) as CUDA tensors (float64).
        out=(U, X, status, iters): contiguous CUDA tensors (views of larger ones are fine) that receive the results."""
        t = self.torch
        xref = self._dev(xref)
        B = xref.shape[0]
        assert xref.shape[1:] == (self.N + 1, 5), xref.shape
        cold = X_init is None and U_init is None
        if out is not None:
            U, X, status, iters = out
            assert U.is_contiguous() and X.is_contiguous() and U.shape == (B, self.N, 2) and X.shape == (B, self.N + 1, 5)
            assert U.dtype == t.float64 and X.dtype == t.float64 and status.dtype == t.int32 and iters.dtype == t.int32
            if not cold:
                X.copy_(xref[:, :1, :].expand(B, self.N + 1, 5) if X_init is None else self._dev(X_init))
                U.copy_(t.zeros_like(U) if U_init is None else self._dev(U_init))
        else:
            if cold:
                X = t.empty(B, self.N + 1, 5, dtype=t.float64, device=self.device)
                U = t.empty(B, self.N, 2, dtype=t.float64, device=self.device)
            else:
                X = (xref[:, :1, :].expand(B, self.N + 1, 5).contiguous() if X_init is None else self._dev(X_init).clone())
                U = (t.zeros(B, self.N, 2, dtype=t.float64, device=self.device) if U_init is None else self._dev(U_init).clone())
            status = t.empty(B, dtype=t.int32, device=self.device)
            iters = t.empty(B, dtype=t.int32, device=self.device)
        h = self.handle
        fn = h.lib.mpcb200_solve_cold if cold else h.lib.mpcb200_solve
        h.check(fn(h.h, xref.data_ptr(), X.data_ptr(), U.data_ptr(), status.data_ptr(), iters.data_ptr(), B, self._stream()))
        return U, X, status, iters

    def set_scenarios(self, scenarios):
        """Scenario table for `solve_batch_scenarios` (BASELINE configs[4]: several scenarios in ONE launch).  `scenarios`: objects
        with the fields `Optimizer.__init__` reads from a configuration (`dt` / `delta_t`, `weights_setting`, `static_obstacle`),
        e.g. `mpc_b200.load_scenario(...)`.  Bounds, horizon and solver options stay this optimizer's."""
        tab = (_capi.Scenario * len(scenarios))()
        for i, sc in enumerate(scenarios):
            w = sc.weights_setting
            tab[i].dt = float(getattr(sc, "dt", getattr(sc, "delta_t", self.delta_t)))
            for j, k in enumerate(("weight_x", "weight_y", "weight_steering_angle", "weight_velocity", "weight_heading_angle")):
                tab[i].Q[j] = float(w[k])
            tab[i].R[0], tab[i].R[1] = float(w["weight_velocity_steering_angle"]), float(w["weight_long_acceleration"])
            circles, r_sum, _ = obstacle_circles_and_radius(sc.static_obstacle, self.configuration.p)
            tab[i].r_sum = float(r_sum)
            for j in range(3):
                tab[i].obstacle[2 * j], tab[i].obstacle[2 * j + 1] = float(circles[j][0]), float(circles[j][1])
        h = self.handle
        h.check(h.lib.mpcb200_set_scenarios(h.h, tab, len(scenarios)))
        self.n_scenarios = len(scenarios)

    def solve_batch_scenarios(self, xref, scenario_id):
        """One cold-start NLP solve per row with the constants of row b taken from scenario `scenario_id[b]` of the table set by
        `set_scenarios` (one launch for the whole mixed batch).  Returns (U, X, status, iters) like `solve_batch`."""
        t = self.torch
        xref = self._dev(xref)
        B = xref.shape[0]
        sid = t.as_tensor(np.ascontiguousarray(scenario_id, np.int32) if not isinstance(scenario_id, t.Tensor) else scenario_id,
                          device=self.device).to(t.int32).contiguous()
        assert sid.shape == (B,) and xref.shape[1:] == (self.N + 1, 5)
        X = t.empty(B, self.N + 1, 5, dtype=t.float64, device=self.device)
        U = t.empty(B, self.N, 2, dtype=t.float64, device=self.device)
        status = t.empty(B, dtype=t.int32, device=self.device)
        iters = t.empty(B, dtype=t.int32, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_solve_scenarios(h.h, xref.data_ptr(), sid.data_ptr(), X.data_ptr(), U.data_ptr(), status.data_ptr(),
                                              iters.data_ptr(), B, self._stream()))
        return U, X, status, iters

    def solve_batch_dual(self, xref, X_init, U_init, lam=None):
        """`solve_batch` with the inequality multipliers / obstacle slacks in and out (`mpcb200_solve_dual`).  lam: CUDA tensor
        [B, lam_words] from a previous call (warm duals) or None (cold duals; a fresh block is returned).
        Returns (U, X, status, iters, lam)."""
        t = self.torch
        xref = self._dev(xref)
        B = xref.shape[0]
        X = self._dev(X_init).clone()
        U = self._dev(U_init).clone()
        h = self.handle
        if lam is None:
            lam = t.zeros(B, int(h.lib.mpcb200_lam_words(h.h)), dtype=t.float64, device=self.device)
        assert lam.is_contiguous() and lam.dtype == t.float64 and lam.shape == (B, int(h.lib.mpcb200_lam_words(h.h)))
        status = t.empty(B, dtype=t.int32, device=self.device)
        iters = t.empty(B, dtype=t.int32, device=self.device)
        h.check(h.lib.mpcb200_solve_dual(h.h, xref.data_ptr(), X.data_ptr(), U.data_ptr(), lam.data_ptr(), status.data_ptr(),
                                         iters.data_ptr(), B, self._stream()))
        return U, X, status, iters, lam

    def solve_batch_stepwise(self, xref, X_init=None, U_init=None, n_iter=None):
        """Same solve with one kernel launch per SQP iteration (KKT slab staged HBM<->smem by TMA each launch)."""
        t = self.torch
        xref = self._dev(xref)
        B = xref.shape[0]
        X = (xref[:, :1, :].expand(B, self.N + 1, 5).contiguous() if X_init is None else self._dev(X_init).clone())
        U = (t.zeros(B, self.N, 2, dtype=t.float64, device=self.device) if U_init is None else self._dev(U_init).clone())
        status = t.empty(B, dtype=t.int32, device=self.device)
        iters = t.empty(B, dtype=t.int32, device=self.device)
        h = self.handle
        s = self._stream()
        h.check(h.lib.mpcb200_sqp_begin(h.h, xref.data_ptr(), X.data_ptr(), U.data_ptr(), B, s))
        for _ in range(self.cfg.max_iter if n_iter is None else n_iter):
            h.check(h.lib.mpcb200_sqp_iter(h.h, 1, s))
        h.check(h.lib.mpcb200_sqp_end(h.h, X.data_ptr(), U.data_ptr(), status.data_ptr(), iters.data_ptr(), s))
        return U, X, status, iters

    def alloc_host_buffers(self, B):
        """Pinned host arrays for `solve_batch_host`'s zero-copy route: numpy views (xref [B,N+1,5], X [B,N+1,5], U [B,N,2],
        float64) of page-locked torch tensors the device can address; fill `xref` in place, pass `out=(X, U)`.
        The tensors are kept alive by the returned arrays (`.base`)."""
        t = self.torch
        mk = lambda *shape: t.empty(*shape, dtype=t.float64).pin_memory().numpy()          # noqa: E731
        return mk(B, self.N + 1, 5), mk(B, self.N + 1, 5), mk(B, self.N, 2)

    def solve_batch_host(self, xref, X_init=None, U_init=None, inplace=False, out=None):
        """End-to-end call with HOST numpy buffers (H2D + solve + D2H inside the library, synchronous).
        X_init = U_init = None: cold start (the reference's step-0 guess), only xref is uploaded.
        out=(X_out, U_out): C-contiguous float64 arrays (ideally pinned) that receive the solution;
        inplace=True: X_init/U_init themselves are overwritten; otherwise fresh arrays are returned."""
        xref = np.ascontiguousarray(xref, np.float64)
        B = xref.shape[0]
        cold = X_init is None and U_init is None
        if cold:
            X_in = U_in = None
            X, U = out if out is not None else (np.empty((B, self.N + 1, 5)), np.empty((B, self.N, 2)))
        else:
            X_in = np.ascontiguousarray(X_init, np.float64)
            U_in = np.ascontiguousarray(U_init, np.float64)
            if out is not None:
                X, U = out
            elif inplace:
                X, U = X_in, U_in
                assert X is X_init and U is U_init, "inplace needs C-contiguous float64 arrays"
            else:
                X, U = np.empty_like(X_in), np.empty_like(U_in)
        assert X.flags.c_contiguous and U.flags.c_contiguous and X.dtype == np.float64 and U.dtype == np.float64
        assert X.shape == (B, self.N + 1, 5) and U.shape == (B, self.N, 2)
        status = np.empty(B, np.int32)
        iters = np.empty(B, np.int32)
        h = self.handle
        ptr = lambda a: a.__array_interface__["data"][0]          # noqa: E731  (a.ctypes.data builds a ctypes object per call)
        h.check(h.lib.mpcb200_solve_host(h.h, ptr(xref), ptr(X_in) if not cold else None, ptr(U_in) if not cold else None,
                                         ptr(X), ptr(U), ptr(status), ptr(iters), B))
        return U, X, status, iters

    def forces_stage_eval(self, z, p, weights_terminal=None):
        """Stage functions + first derivatives of the reference's FORCESPRO formulation (optimizer.py:90-195) for n points:
        z [n,7], p [n,10] -> dict(c[n,5], dc[n,5,7], h[n,10], dh[n,10,7], f[n], df[n,7], fN[n], dfN[n,7]) as CUDA tensors.
        `weights_terminal`: the five weight_*_terminate values (default: from weights_setting)."""
        import ctypes
        t = self.torch
        z, p = self._dev(z), self._dev(p)
        n = z.shape[0]
        assert z.shape == (n, 7) and p.shape == (n, 10)
        if weights_terminal is None:
            w = self.weights_setting
            weights_terminal = [w[k] for k in ("weight_x_terminate", "weight_y_terminate", "weight_steering_angle_terminate",
                                               "weight_velocity_terminate", "weight_heading_angle_terminate")]
        wt = (ctypes.c_double * 5)(*[float(v) for v in weights_terminal])
        out = t.empty(n, 136, dtype=t.float64, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_forces_stage_eval(h.h, wt, z.data_ptr(), p.data_ptr(), out.data_ptr(), n, self._stream()))
        return dict(c=out[:, 0:5], dc=out[:, 5:40].reshape(n, 5, 7), h=out[:, 40:50], dh=out[:, 50:120].reshape(n, 10, 7),
                    f=out[:, 120], df=out[:, 121:128], fN=out[:, 128], dfN=out[:, 129:136])

    def plant_step_shift(self, x, U, X):
        """shift_movement (optimizer.py:645-655) on the device, in place.  Returns applied controls [B,2]."""
        t = self.torch
        B = x.shape[0]
        u_applied = t.empty(B, 2, dtype=t.float64, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_plant_step_shift(h.h, x.data_ptr(), U.data_ptr(), X.data_ptr(), u_applied.data_ptr(), B,
                                               self._stream()))
        return u_applied

    def build_ref_window(self, i, x):
        """desired_command_and_trajectory (optimizer.py:657-702) on the device -> X_ref [B,N+1,5]."""
        t = self.torch
        path, orient = self._path_tensors()
        B = x.shape[0]
        xref = t.empty(B, self.N + 1, 5, dtype=t.float64, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_build_ref_window(h.h, int(i), int(self.iter_length), path.data_ptr(), orient.data_ptr(),
                                               float(self.desired_velocity), x.data_ptr(), xref.data_ptr(), B,
                                               self._stream()))
        return xref

    def optimize_batch(self, x0, return_device=False):
        """The reference's whole receding-horizon loop (optimizer.py:596-631) for B egos, entirely on the device.
        x0 [B,5] -> (states[B,T,5], controls[B,T,2], status[B,T], iters[B,T])."""
        t = self.torch
        x0 = self._dev(x0)
        B, T = x0.shape[0], int(self.iter_length)
        path, orient = self._path_tensors()
        traj = t.empty(B, T, 5, dtype=t.float64, device=self.device)
        ctrl = t.empty(B, T, 2, dtype=t.float64, device=self.device)
        status = t.empty(B, T, dtype=t.int32, device=self.device)
        iters = t.empty(B, T, dtype=t.int32, device=self.device)
        h = self.handle
        h.check(h.lib.mpcb200_closed_loop(h.h, T, path.data_ptr(), orient.data_ptr(), float(self.desired_velocity),
                                          x0.data_ptr(), traj.data_ptr(), ctrl.data_ptr(), status.data_ptr(),
                                          iters.data_ptr(), B, self._stream()))
        if return_device:
            return traj, ctrl, status, iters
        return traj.cpu().numpy(), ctrl.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()

    # ------------------------------------------------------------------ the reference contract
    def optimize(self, on_device=None):
        """CasadiOptimizer.optimize() (optimizer.py:562-643): returns (traj_s[T,5], u[T,2], t_v[T]).

        Noise-free runs (`configuration.noised` false) execute the whole receding-horizon loop on the device in ONE launch
        (`mpcb200_closed_loop` = `optimize_batch` with B = 1); `t_v` then holds the measured loop time divided evenly over the
        T steps (the device loop has no per-step host clock).  With `noised`, with a float32 handle that needs the float64
        refinement pass (obstacle within reach of the path), or with `on_device=False`, the loop runs step by
        step from the host in the reference's own shape -- solve -> first control (+ noise) -> plant step + shift -> next
        window -- with the solve time of each step measured like the reference does (wall clock around the solve,
        optimizer.py:603-608).  `noised` uses the reference's noise law and needs N == 10 (optimizer.py:611-615, quirk Q9)."""
        t = self.torch
        N, T = self.N, int(self.iter_length)
        init_state = np.array([self.init_position[0], self.init_position[1], 0.0, self.init_velocity,
                               self.init_orientation], float)
        noised = bool(getattr(self.configuration, "noised", False))
        if on_device is None:
            # the device loop has no float64 refinement pass: a float32 handle whose obstacle is within reach (refine_f64 set)
            # runs the loop step by step from the host instead, where every solve is followed by its refinement launch
            on_device = not noised and not (self.cfg.precision == _capi.F32 and self.cfg.refine_f64)
        if on_device:
            if noised:
                raise ValueError("the device loop is noise-free; use on_device=False with configuration.noised")
            t.cuda.synchronize(self.device)
            t_ = time.time()
            traj, ctrl, status, iters = self.optimize_batch(init_state[None, :])
            dt_loop = time.time() - t_
            self.last_status, self.last_iters = status[0], iters[0]
            return traj[0], ctrl[0], np.full(T, dt_loop / T)
        x = self._dev(init_state[None, :])
        xref = x[:, None, :].expand(1, N + 1, 5).contiguous()      # Q4: first parameter block = x0 tiled
        X = xref.clone()
        U = t.zeros(1, N, 2, dtype=t.float64, device=self.device)
        traj, u_c, t_v, sts, its = [], [], [], [], []
        for i in range(T):
            t.cuda.synchronize(self.device)
            t_ = time.time()
            U, X, status, iters = self.solve_batch(xref, X, U)
            t.cuda.synchronize(self.device)
            t_v.append(time.time() - t_)
            sts.append(status); its.append(iters)
            if noised:
                if N != 10:
                    raise ValueError("noised=True draws 20 = 2*10 samples in the reference (optimizer.py:613); N must be 10")
                sigma = 0.1 if getattr(self.configuration, "use_case", "lane_following") == "lane_following" else 0.05
                noise = np.random.normal(0, sigma, 20).reshape(2, N).T
                U = U + self._dev(noise[None])
            u_c.append(self.plant_step_shift(x, U, X))
            traj.append(x.clone())
            xref = self.build_ref_window(i, x)
        # one device -> host transfer for the whole run
        traj_s = t.cat(traj, dim=0).cpu().numpy()
        u_out = t.cat(u_c, dim=0).cpu().numpy()
        self.last_status, self.last_iters = t.cat(sts).cpu().numpy(), t.cat(its).cpu().numpy()
        traj_s = np.insert(traj_s, 0, init_state, axis=0)
        traj_s = np.delete(traj_s, -1, axis=0)
        return traj_s, u_out, np.array(t_v)

    def save_results(self, save_path, states, controls, solve_time):
        """Writes `planned states.txt`, `control inputs.txt`, `solve time.txt`, `deviation.txt`, `RMSD.txt` in the format
        MPCPlanner.plot_* writes them (mpc_planner.py:190-290), so a run diffs against the reference's recorded fixtures."""
        from . import results
        origin = getattr(self.configuration, "origin_reference_path", None)
        return results.write_result_files(save_path, states, controls, solve_time,
                                          np.asarray(self.resampled_path_points, float), origin)


def make_configuration(scenario, predict_horizon=None, framework_name="casadi", noised=False):
    """A `PlanningConfiguration`-shaped object (configuration.py:106-336) built from mpc_b200.scenarios data, for use
    where commonroad is not installed.  Only the fields Optimizer reads are populated."""
    from types import SimpleNamespace
    return SimpleNamespace(p=VehicleParameters2, iter_length=scenario.iter_length, delta_t=scenario.dt,
                           desired_velocity=scenario.desired_velocity, reference_path=scenario.reference_path,
                           orientation=scenario.orientation, weights_setting=scenario.weights_setting,
                           static_obstacle=scenario.static_obstacle, noised=noised, use_case=scenario.use_case,
                           wheelbase=scenario.wheelbase, framework_name=framework_name,
                           predict_horizon=predict_horizon, origin_reference_path=getattr(scenario, "origin_reference_path", None),
                           left_road_boundary=getattr(scenario, "left_road_boundary", None),
                           right_road_boundary=getattr(scenario, "right_road_boundary", None))


def init_values_from_state(x0):
    """(position, velocity, acceleration, orientation) as MPCPlanner.get_init_values returns (mpc_planner.py:30-59)."""
    return (np.array([x0[0], x0[1]]), float(x0[3]), 0.0, float(x0[4]))
