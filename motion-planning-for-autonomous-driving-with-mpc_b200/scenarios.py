"""Scenario data (the optimizer's input schema) and the synthetic batch generator of SURVEY.md §8(d).

`data/scenarios.json` is produced by tools/extract_scenarios.py from the reference's CommonRoad XML + YAML
files; it holds exactly the fields `Optimizer.__init__` reads from `configuration`
(/root/reference/MPC_Planner/optimizer.py:51-68): iter_length, delta_t, desired_velocity, reference_path,
orientation, weights_setting, static_obstacle (+ x0 from the planning problem).
"""
import json
import os
from types import SimpleNamespace

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scenarios.json")
_cache = None

# perturbation law of the batched configs: sigma per state component
PERTURB_SIGMA = np.array([0.5, 0.3, 0.01, 1.0, 0.05])


def _all():
    global _cache
    if _cache is None:
        with open(_DATA) as f:
            _cache = json.load(f)
    return _cache


def scenario_names():
    return list(_all().keys())


def load_scenario(name):
    """Returns a namespace with numpy fields: x0[5], reference_path[T,2], orientation[T], desired_velocity,
    iter_length, dt, weights_setting (dict), static_obstacle (dict), use_case."""
    d = _all()[name]
    return SimpleNamespace(name=name, x0=np.array(d["x0"], float), reference_path=np.array(d["reference_path"], float),
                           orientation=np.array(d["orientation"], float), desired_velocity=float(d["desired_velocity"]),
                           iter_length=int(d["iter_length"]), dt=float(d["dt"]), weights_setting=dict(d["weights_setting"]),
                           static_obstacle=dict(d["static_obstacle"]), use_case=d["use_case"],
                           wheelbase=float(d.get("wheelbase", 2.578)), synthesised=bool(d.get("synthesised", False)),
                           origin_reference_path=(np.array(d["origin_reference_path"], float) if "origin_reference_path" in d else None),
                           left_road_boundary=(np.array(d["left_road_boundary"], float) if "left_road_boundary" in d else None),
                           right_road_boundary=(np.array(d["right_road_boundary"], float) if "right_road_boundary" in d else None))


def perturbed_initial_states(sc, B, seed, r_clear=None, obstacle_circles=None, ego_offset=0.75):
    """x0 + eps, eps ~ N(0, diag(PERTURB_SIGMA)^2); |delta0| <= 0.05, v0 in [1, 25] (friction row feasible and
    convex).  With `r_clear`, instances whose three ego circles are closer than r_clear to the matching
    obstacle circle are redrawn (config 3)."""
    rng = np.random.default_rng(seed)
    out = np.empty((B, 5))
    n = 0
    while n < B:
        m = B - n
        x = sc.x0[None, :] + rng.normal(size=(m, 5)) * PERTURB_SIGMA[None, :]
        x[:, 2] = np.clip(x[:, 2], -0.05, 0.05)
        x[:, 3] = np.clip(x[:, 3], 1.0, 25.0)
        if r_clear is not None:
            c, s = np.cos(x[:, 4]), np.sin(x[:, 4])
            ok = np.ones(m, bool)
            for j, sg in enumerate((0.0, 1.0, -1.0)):
                dx = x[:, 0] + sg * ego_offset * c - obstacle_circles[j][0]
                dy = x[:, 1] + sg * ego_offset * s - obstacle_circles[j][1]
                ok &= np.hypot(dx, dy) > r_clear
            x = x[ok]
        out[n:n + len(x)] = x
        n += len(x)
    return out


def reference_window(i, x_now, N, iter_length, path, orientation, desired_velocity):
    """Host mirror of desired_command_and_trajectory (optimizer.py:657-702) for a batch of current states
    x_now [B,5] -> X_ref [B,N+1,5] (row 0 = current state; window freezes for i >= iter_length-N, quirk Q8)."""
    x_now = np.atleast_2d(np.asarray(x_now, float))
    B = x_now.shape[0]
    k = np.arange(N)
    idx = k + (iter_length - N) if i >= iter_length - N else i + k + 1
    rows = np.stack([path[idx, 0], path[idx, 1], np.zeros(N), np.full(N, desired_velocity), orientation[idx]], axis=1)
    out = np.empty((B, N + 1, 5))
    out[:, 0] = x_now
    out[:, 1:] = rows[None]
    return out


def make_batch(name, B, N, seed, mpc_step=0):
    """Synthetic batched instance of one scenario: (sc, x0[B,5], xref[B,N+1,5], X_init[B,N+1,5], U_init[B,N,2]).
    Cold start exactly as the reference's first MPC step: states = x0 tiled, controls = 0 (optimizer.py:580-583)."""
    sc = load_scenario(name)
    from .optimizer import obstacle_circles_and_radius
    circles, r_sum, off = obstacle_circles_and_radius(sc.static_obstacle)
    x0 = perturbed_initial_states(sc, B, seed, r_clear=r_sum + 0.05, obstacle_circles=circles, ego_offset=off)
    T = sc.iter_length
    if N > T:
        raise ValueError(f"N={N} needs a path of at least N points, scenario {name} has {T}")
    xref = reference_window(mpc_step, x0, N, T, sc.reference_path, sc.orientation, sc.desired_velocity)
    X = np.repeat(x0[:, None, :], N + 1, axis=1)
    U = np.zeros((B, N, 2))
    return sc, x0, xref, X, U
