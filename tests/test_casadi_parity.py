"""The verbatim casadi/IPOPT branch of the oracle (oracle/casadi_ref.py).  casadi is an un-vendored dependency of
the reference that is absent from this image: the IPOPT comparisons skip when it cannot be imported (and say so); the parts of
the branch that need no casadi -- the reference's bound lists -- are always checked against the restatement."""
import os

import numpy as np
import pytest

import mpc_b200
from oracle import casadi_ref, ipm, nlp

G = os.path.join(os.path.dirname(__file__), "golden")
needs_casadi = pytest.mark.skipif(not casadi_ref.available(), reason="casadi (>=3.5.1, IPOPT+MUMPS) is not installed in this image: "
                                                                    "parity stays unpinned at the IPOPT boundary")


@pytest.mark.parametrize("name,N", [("ZAM_Over-1_1_LF", 30), ("ZAM_Over-1_1_CA", 30), ("USA_Lanker-2_18_T-1_LF", 50), ("ZAM_Over-1_1_LF", 10)])
def test_reference_bound_lists_equal_the_restated_bounds(name, N):
    """inequal_constraints (optimizer.py:413-491) written out as the reference's Python lists == oracle.nlp.g_bounds, except the
    one documented difference: the friction row is stated as q in [-a_max, a_max] instead of |q| in [0, a_max] (same set)."""
    sc = mpc_b200.load_scenario(name)
    d = nlp.make_nlp(N, sc.dt, sc.weights_setting, np.tile(sc.x0, (N + 1, 1)), sc.static_obstacle)
    lbg, ubg, lbx, ubx = casadi_ref.bounds(N, sc.static_obstacle)
    Lg, Ug, Lx, Ux = nlp.g_bounds(d)
    assert len(lbg) == d.m and len(lbx) == d.n
    assert lbg[0] == 0.0 and Lg[0] == -ubg[0]
    assert np.array_equal(lbg[1:], Lg[1:]) and np.array_equal(ubg, Ug) and np.array_equal(lbx, Lx) and np.array_equal(ubx, Ux)
    d.friction_smooth = False
    assert np.array_equal(lbg, nlp.g_bounds(d)[0])


@needs_casadi
@pytest.mark.parametrize("key,name,N", [("lf_zam_n30", "ZAM_Over-1_1_LF", 30), ("lf_lanker_n50", "USA_Lanker-2_18_T-1_LF", 50),
                                         ("lf_zam_n10", "ZAM_Over-1_1_LF", 10)])
def test_oracle_equals_ipopt_on_the_golden_instances(key, name, N):
    """oracle (restated NLP + own IPM) == casadi/IPOPT on the verbatim reference NLP, <= 1e-6."""
    g = np.load(os.path.join(G, "nlp_solutions.npz"))
    sc = mpc_b200.load_scenario(name)
    for b, xref in enumerate(g[key + "_xref"]):
        X0 = np.tile(xref[0], (N + 1, 1))
        w, ok = casadi_ref.solve_instance(sc, N, xref, X0, np.zeros((N, 2)))
        assert ok
        Ui, Xi = nlp.split(w, N)
        assert np.abs(Ui - g[key + "_U"][b]).max() < 1e-6 and np.abs(Xi - g[key + "_X"][b]).max() < 1e-6
        d = nlp.make_nlp(N, sc.dt, sc.weights_setting, xref, sc.static_obstacle)
        assert ipm.kkt_error(d, w)[0] < 1e-6                        # IPOPT's point is a KKT point of the RESTATED NLP


@needs_casadi
def test_oracle_closed_loop_equals_ipopt_closed_loop():
    from oracle import closed_loop
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    tr_i, u_i = casadi_ref.closed_loop(sc, 10, scramble=True)
    tr_o, u_o = closed_loop.optimize(sc, 10)
    assert np.abs(tr_i - tr_o).max() < 1e-5 and np.abs(u_i - u_o).max() < 1e-5


@needs_casadi
@pytest.mark.gpu
def test_cuda_solver_equals_ipopt_on_config2_sample():
    """The CUDA solver against IPOPT itself (not the port): 64 instances of BASELINE configs[1], stated tolerance 1e-3."""
    import torch
    from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
    N, B = 30, 64
    sc, x0, xref, X0, U0 = mpc_b200.make_batch("ZAM_Over-1_1_LF", B, N, 20261017)
    opt = B200Optimizer(make_configuration(sc, N), init_values_from_state(sc.x0), N, max_batch=B)
    U, X, st, it = [t.cpu().numpy() for t in opt.solve_batch(xref)]
    assert (st == 1).all()
    for b in range(B):
        w, ok = casadi_ref.solve_instance(sc, N, xref[b], X0[b], U0[b])
        assert ok
        Ui, Xi = nlp.split(w, N)
        assert np.abs(Ui - U[b]).max() < 1e-3 and np.abs(Xi - X[b]).max() < 1e-3
