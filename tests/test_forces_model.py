"""The FORCESPRO-formulation stage functions (csrc/forces_model.cuh: RK4 dynamics, friction circle + nine squared circle distances,
stage / terminal objective, and all first derivatives) against the reference's OWN CasADi-generated C model
(test/FORCESNLPsolver/FORCESNLPsolver_model.c): the committed known answers (tests/golden/forces_model_kat.npz, produced by
tools/make_golden.py from oracle/_ref/libforces_model.so) and, where the compiled model is present, live on fresh random points.
CPU: the same source through tests/host_sim.  GPU (-m gpu): `mpcb200_forces_stage_eval` through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import hostsim

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libforces_model.so")
# weights baked into the generated model (FORCESNLPsolver_model.c:77-117 stage, :1223-1254 terminal), dt, p.a+p.b, wheelbase, circle offset
CONSTS = np.array([0.1, 2.5789128, 2.578, 0.75, 2, 2, 50, 0.1, 5, 2, 0.2, 4, 4, 100, 0.2, 10], float)
TOL = 1e-12


def _check(r, k):
    sc = lambda a: max(1.0, float(np.abs(a).max()))          # noqa: E731
    assert np.abs(r["c"] - k["dynamics"]).max() < TOL
    assert np.abs(r["dc"] - k["ddynamics"]).max() < TOL
    assert np.abs(r["h"] - k["inequalities"]).max() < TOL * sc(k["inequalities"])
    assert np.abs(r["dh"] - k["dinequalities"]).max() < TOL * sc(k["dinequalities"])
    assert np.abs(r["f"] - k["objective"][:, 0]).max() < TOL * sc(k["objective"])
    assert np.abs(r["fN"] - k["objective"][:, 1]).max() < TOL * sc(k["objective"])
    assert np.abs(r["df"] - k["dobjective"][:, 0]).max() < TOL * sc(k["dobjective"])
    assert np.abs(r["dfN"] - k["dobjective"][:, 1]).max() < TOL * sc(k["dobjective"])
    # structure the generated code declares: dc has 21 non-zeros, dh 27 (FORCESNLPsolver_model.c casadi_s5 / casadi_s6)
    assert int((np.abs(k["ddynamics"]).max(axis=0) > 0).sum()) == 21 and int((np.abs(r["dc"]).max(axis=0) > 0).sum()) == 21
    assert int((np.abs(r["dh"]).max(axis=0) > 0).sum()) == 27


def test_host_build_matches_the_generated_model_known_answers():
    k = np.load(os.path.join(G, "forces_model_kat.npz"))
    _check(hostsim.forces_eval(CONSTS, k["z"], k["p"]), k)


def _ref_eval(z, p):
    """Call the compiled reference model (oracle/_ref, built by oracle/Makefile from the file under /root/reference)."""
    lib = C.CDLL(REF_SO)

    def call(name, zz, pp, nout):
        arg = (C.POINTER(C.c_double) * 2)(zz.ctypes.data_as(C.POINTER(C.c_double)), pp.ctypes.data_as(C.POINTER(C.c_double)))
        o = np.zeros(nout)
        res = (C.POINTER(C.c_double) * 1)(o.ctypes.data_as(C.POINTER(C.c_double)))
        iw = (C.c_int * 64)(); w = (C.c_double * 512)()
        getattr(lib, name)(arg, res, iw, w, 0)
        return o

    def dense(name, vals):
        sp = getattr(lib, name + "_sparsity_out")
        sp.restype = C.POINTER(C.c_int)
        s = sp(0)
        nrow, ncol = s[0], s[1]
        colind = [s[2 + i] for i in range(ncol + 1)]
        rows = [s[2 + ncol + 1 + i] for i in range(colind[-1])]
        M = np.zeros((nrow, ncol))
        for c in range(ncol):
            for q in range(colind[c], colind[c + 1]):
                M[rows[q], c] = vals[q]
        return M

    def nnz(name):
        sp = getattr(lib, name + "_sparsity_out")
        sp.restype = C.POINTER(C.c_int)
        s = sp(0)
        return [s[2 + i] for i in range(s[1] + 1)][-1]

    n = z.shape[0]
    out = dict(dynamics=np.zeros((n, 5)), ddynamics=np.zeros((n, 5, 7)), inequalities=np.zeros((n, 10)), dinequalities=np.zeros((n, 10, 7)),
               objective=np.zeros((n, 2)), dobjective=np.zeros((n, 2, 7)))
    for i in range(n):
        zz, pp = np.ascontiguousarray(z[i]), np.ascontiguousarray(p[i])
        out["dynamics"][i] = call("FORCESNLPsolver_dynamics_0", zz, pp, 5)
        M = dense("FORCESNLPsolver_ddynamics_0", call("FORCESNLPsolver_ddynamics_0", zz, pp, nnz("FORCESNLPsolver_ddynamics_0")))
        out["ddynamics"][i] = M if M.shape == (5, 7) else M.T
        out["inequalities"][i] = call("FORCESNLPsolver_inequalities_0", zz, pp, 10)
        M = dense("FORCESNLPsolver_dinequalities_0", call("FORCESNLPsolver_dinequalities_0", zz, pp, nnz("FORCESNLPsolver_dinequalities_0")))
        out["dinequalities"][i] = M if M.shape == (10, 7) else M.T
        for t in (0, 1):
            out["objective"][i, t] = call(f"FORCESNLPsolver_objective_{t}", zz, pp, 1)[0]
            nm = f"FORCESNLPsolver_dobjective_{t}"
            out["dobjective"][i, t] = dense(nm, call(nm, zz, pp, nnz(nm))).reshape(-1)
    return out


def _random_points(n, seed):
    rng = np.random.default_rng(seed)
    z = np.stack([rng.uniform(-0.4, 0.4, n), rng.uniform(-11, 11, n), rng.uniform(-50, 150, n), rng.uniform(-20, 20, n),
                  rng.uniform(-1.0, 1.0, n), rng.uniform(0.0, 40, n), rng.uniform(-3.2, 3.2, n)], axis=1)
    p = np.concatenate([rng.uniform(-50, 150, (n, 1)), rng.uniform(-20, 20, (n, 1)), rng.uniform(0, 25, (n, 1)), rng.uniform(-3.2, 3.2, (n, 1)),
                        rng.uniform(-50, 150, (n, 6))], axis=1)
    return z, p


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libforces_model.so not built (needs /root/reference at build time)")
def test_host_build_matches_the_compiled_reference_model_on_random_points():
    z, p = _random_points(512, 7)
    _check(hostsim.forces_eval(CONSTS, z, p), _ref_eval(z, p))


@pytest.mark.gpu
def test_cuda_forces_stage_eval_matches_the_generated_model():
    import mpc_b200
    from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    ws = dict(sc.weights_setting)
    for key, v in zip(("weight_x", "weight_y", "weight_steering_angle", "weight_velocity", "weight_heading_angle",
                       "weight_velocity_steering_angle", "weight_long_acceleration"), CONSTS[4:11]):
        ws[key] = float(v)
    sc.weights_setting = ws
    opt = B200Optimizer(make_configuration(sc, 10), init_values_from_state(sc.x0), 10, max_batch=8)
    k = np.load(os.path.join(G, "forces_model_kat.npz"))
    r = {a: b.cpu().numpy() for a, b in opt.forces_stage_eval(k["z"], k["p"], weights_terminal=CONSTS[11:16]).items()}
    _check(r, k)
    if os.path.exists(REF_SO):                       # the compiled reference model travels to the GPU box with the snapshot
        z, p = _random_points(4099, 11)             # ragged last CTA
        r = {a: b.cpu().numpy() for a, b in opt.forces_stage_eval(z, p, weights_terminal=CONSTS[11:16]).items()}
        _check(r, _ref_eval(z, p))
