"""Result files in the reference's format (mpc_b200/results.py) against the reference's recorded runs
(tests/golden/recorded_runs.npz = /root/reference/test/2D_plots_*/*.txt verbatim, made by tools/make_golden.py)."""
import os

import numpy as np

import mpc_b200
from mpc_b200 import results

G = os.path.join(os.path.dirname(__file__), "golden")


def test_restated_paths_reproduce_every_recorded_rmsd_and_deviation_file():
    """The commonroad-free scenario front-end (tools/extract_scenarios.py -> data/scenarios.json) against ALL six recorded runs
    of the reference: `deviation.txt` is the distance of each recorded state to the closest vertex of the route planner's
    ORIGIN reference path (mpc_planner.py:190-197), `RMSD.txt` the RMS distance to the clipped / Chaikin-smoothed / resampled
    path the optimizer tracks (mpc_planner.py:279-290).  Both are reproduced to rounding for ZAM_Over-1_1 (lane following and
    collision avoidance) and USA_Lanker-2_18_T-1 (lane changes 3452 -> 3454 -> 3456), CasADi and Forcespro runs: the restated
    paths, the desired velocity rule and the resampling ARE the reference's, not an approximation of them."""
    g = np.load(os.path.join(G, "recorded_runs.npz"))
    n_dev = 0
    for name, key in (("ZAM_Over-1_1_LF", "zam_lf"), ("ZAM_Over-1_1_CA", "zam_ca"), ("USA_Lanker-2_18_T-1_LF", "lanker_lf")):
        sc = mpc_b200.load_scenario(name)
        assert sc.origin_reference_path is not None and sc.reference_path.shape[0] == sc.iter_length
        for solver in ("casadi", "forcespro"):
            x = g[f"{solver}_{key}_x"]
            assert x.shape[0] == sc.iter_length and np.array_equal(x[0], sc.x0)
            dev = results.deviation(x, sc.origin_reference_path)
            assert np.abs(dev - g[f"{solver}_{key}_dev"]).max() < 1e-9
            n_dev += len(dev)
            if f"{solver}_{key}_rmsd" in g.files:                       # the reference writes RMSD.txt for lane following only
                assert np.abs(results.rmsd(x, sc.reference_path) - g[f"{solver}_{key}_rmsd"]).max() < 1e-7
    assert n_dev == 2 * (30 + 30 + 70)
    # first recorded CasADi step brakes at the friction limit (quirks Q3/Q4): a0 = -sqrt(11.5) + N(0, 0.1^2) noise
    assert abs(g["casadi_zam_lf_u"][0, 1] + np.sqrt(11.5)) < 0.5


def test_result_files_round_trip_in_the_reference_format(tmp_path):
    g = np.load(os.path.join(G, "recorded_runs.npz"))
    sc = mpc_b200.load_scenario("USA_Lanker-2_18_T-1_LF")
    x, u, t = g["casadi_lanker_lf_x"], g["casadi_lanker_lf_u"], g["casadi_lanker_lf_t"]
    out = results.write_result_files(str(tmp_path / "run"), x, u, t, sc.reference_path)
    back = results.read_result_files(str(tmp_path / "run"))
    assert sorted(back) == sorted(results.FILES)
    assert back["planned states.txt"].shape == (70, 5) and back["control inputs.txt"].shape == (70, 2)
    assert back["solve time.txt"].shape == (70,) and back["deviation.txt"].shape == (70,) and back["RMSD.txt"].shape == (2,)
    assert np.array_equal(back["planned states.txt"], x) and np.array_equal(back["control inputs.txt"], u)     # %.18e is lossless
    assert np.allclose(back["RMSD.txt"], out["RMSD.txt"])
    with open(tmp_path / "run" / "planned states.txt") as f:
        assert len(f.readline().split()) == 5
