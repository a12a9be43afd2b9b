"""Result files in the reference's own format, so a run diffs directly against its recorded fixtures
(/root/reference/test/2D_plots_<framework>_<scenario>_<use_case>/{planned states,control inputs,solve time,deviation,RMSD}.txt).

Restates the numeric part of MPCPlanner.plot_* (plots themselves are out of scope):
    planned states.txt   np.savetxt(x)            mpc_planner.py:255
    control inputs.txt   np.savetxt(u)            mpc_planner.py:211
    solve time.txt       np.savetxt(solve_time)   mpc_planner.py:237
    deviation.txt        distance to the closest point of the original reference path   mpc_planner.py:190-197
    RMSD.txt             sqrt(sum (ref - x)^2 / (T-1)) in x and y                       mpc_planner.py:279-290
"""
import os

import numpy as np

FILES = ("planned states.txt", "control inputs.txt", "solve time.txt", "deviation.txt", "RMSD.txt")


def find_closest_point(path_points, current_point):
    """configuration.py:26-37"""
    d = np.asarray(path_points, float)[:, :2] - np.asarray(current_point, float).reshape(1, 2)
    return int(np.argmin(d[:, 0] ** 2 + d[:, 1] ** 2))


def deviation(x, origin_reference_path):
    """mpc_planner.py:190-197"""
    x = np.asarray(x, float)
    near = np.array([origin_reference_path[find_closest_point(origin_reference_path, x[i, 0:2])][:2] for i in range(x.shape[0])])
    return np.sqrt((near[:, 0] - x[:, 0]) ** 2 + (near[:, 1] - x[:, 1]) ** 2)


def rmsd(x, reference_path, iter_length=None):
    """mpc_planner.py:279-288 -- note the (T - 1) denominator."""
    x = np.asarray(x, float)
    T = int(iter_length if iter_length is not None else x.shape[0])
    ref = np.asarray(reference_path, float)
    sx = float(((ref[:T, 0] - x[:T, 0]) ** 2).sum())
    sy = float(((ref[:T, 1] - x[:T, 1]) ** 2).sum())
    return np.array([np.sqrt(sx / (T - 1)), np.sqrt(sy / (T - 1))])


def write_result_files(save_path, x, u, solve_time, reference_path, origin_reference_path=None):
    """Writes the five txt files of a closed-loop run; returns {file name: array}."""
    os.makedirs(save_path, exist_ok=True)
    origin = reference_path if origin_reference_path is None else origin_reference_path
    out = {"planned states.txt": np.asarray(x, float), "control inputs.txt": np.asarray(u, float),
           "solve time.txt": np.asarray(solve_time, float), "deviation.txt": deviation(x, origin),
           "RMSD.txt": rmsd(x, reference_path)}
    for name, a in out.items():
        np.savetxt(os.path.join(save_path, name), a)
    return out


def read_result_files(path):
    return {name: np.loadtxt(os.path.join(path, name)) for name in FILES if os.path.exists(os.path.join(path, name))}
