"""ORACLE (test infrastructure, NOT product code) -- float64 primal-dual interior-point solve of the reference NLP.

Stands in for `ca.nlpsol('solver', 'ipopt', ...)` + `sol(x0, p, lbg, ubg, lbx, ubx)`
(/root/reference/MPC_Planner/optimizer.py:554-558, 607): same decision vector, same g, same bounds
(oracle/nlp.py), exact Lagrangian Hessian, log-barrier on bounds, slacks on the inequality rows of g,
monotone barrier update, fraction-to-the-boundary rule, l1-merit backtracking line search -- the published IPOPT
algorithm (Waechter & Biegler 2006) minus its filter/restoration phase.  casadi/IPOPT itself is absent from this
image (parity unpinned, see oracle/nlp.py header); `kkt_error()` below is the algorithm-independent check the tests
use: any point it accepts at 1e-8 is a KKT point of the reference NLP no matter which solver produced it.

Deliberately a DIFFERENT algorithm from the CUDA kernel (exact Hessian, sparse LU on the full KKT system, reference
variable ordering) so that agreement between the two is evidence, not tautology.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import nlp

INF = np.inf


class _Layout:
    """v = [w ; s]; c(v) = [g_E(w) - b_E ; g_I(w) - s]; bounds on v."""

    def __init__(self, d, model=None):
        lbg, ubg, lbx, ubx = (model or nlp).g_bounds(d)
        self.eq = np.where(lbg == ubg)[0]
        self.iq = np.where(lbg != ubg)[0]
        self.b_eq = lbg[self.eq]
        self.n, self.ms = d.n, len(self.iq)
        self.L = np.concatenate([lbx, lbg[self.iq]])
        self.U = np.concatenate([ubx, ubg[self.iq]])
        self.hasL = np.isfinite(self.L)
        self.hasU = np.isfinite(self.U)
        self.nv = self.n + self.ms
        self.mc = len(self.eq) + self.ms
        # selection matrices
        m = d.m
        self.S_eq = sp.csr_matrix((np.ones(len(self.eq)), (np.arange(len(self.eq)), self.eq)), shape=(len(self.eq), m))
        self.S_iq = sp.csr_matrix((np.ones(self.ms), (np.arange(self.ms), self.iq)), shape=(self.ms, m))


def _push_interior(v, L, U, hasL, hasU, k1=1e-2, k2=1e-2):
    """IPOPT's initial-point projection (bound_push / bound_frac)."""
    v = v.copy()
    both = hasL & hasU
    pl = np.where(hasL, k1 * np.maximum(1.0, np.abs(np.where(hasL, L, 0.0))), 0.0)
    pu = np.where(hasU, k1 * np.maximum(1.0, np.abs(np.where(hasU, U, 0.0))), 0.0)
    span = np.where(both, U - L, INF)
    pl = np.where(both, np.minimum(pl, k2 * span), pl)
    pu = np.where(both, np.minimum(pu, k2 * span), pu)
    lo = np.where(hasL, L + pl, -INF)
    hi = np.where(hasU, U - pu, INF)
    return np.minimum(np.maximum(v, lo), hi)


def solve(d, w0, tol=1e-9, max_iter=300, mu0=0.1, verbose=False, model=None):
    """Returns dict(w, lam_g, iters, status, kkt, obj).  status 1 = converged.
    `model`: the module that states the NLP (g_bounds, g_fun, g_jac, cost, cost_grad, lag_hess); default oracle.nlp (the
    CasADi formulation), oracle.forces_nlp for the FORCESPRO formulation."""
    nlp = model or globals()["nlp"]
    lay = _Layout(d, nlp)
    n, ms, nv = lay.n, lay.ms, lay.nv
    L, U, hasL, hasU = lay.L, lay.U, lay.hasL, lay.hasU
    Lz = np.where(hasL, L, 0.0)
    Uz = np.where(hasU, U, 0.0)

    w = np.asarray(w0, float).copy()
    g = nlp.g_fun(d, w)
    v = _push_interior(np.concatenate([w, g[lay.iq]]), L, U, hasL, hasU)
    zL = np.where(hasL, 1.0, 0.0)
    zU = np.where(hasU, 1.0, 0.0)
    lam = np.zeros(lay.mc)
    mu = mu0
    I_s = sp.identity(ms, format="csr")

    def evaluate(v):
        w = v[:n]
        g = nlp.g_fun(d, w)
        c = np.concatenate([g[lay.eq] - lay.b_eq, g[lay.iq] - v[n:]])
        return g, c

    def jac(v):
        Jg = nlp.g_jac(d, v[:n])
        top = sp.hstack([lay.S_eq @ Jg, sp.csr_matrix((len(lay.eq), ms))])
        bot = sp.hstack([lay.S_iq @ Jg, -I_s])
        return sp.vstack([top, bot]).tocsr()

    def grad_f(v):
        return np.concatenate([nlp.cost_grad(d, v[:n]), np.zeros(ms)])

    def lam_to_g(lam):
        lg = np.zeros(d.m)
        lg[lay.eq] = lam[:len(lay.eq)]
        lg[lay.iq] = lam[len(lay.eq):]
        return lg

    def barrier(v, mu):
        dl = np.where(hasL, v - Lz, 1.0)
        du = np.where(hasU, Uz - v, 1.0)
        if np.any(dl <= 0) or np.any(du <= 0):
            return INF
        return nlp.cost(d, v[:n]) - mu * (np.sum(np.log(dl[hasL])) + np.sum(np.log(du[hasU])))

    def errors(v, lam, zL, zU, mu, J, c, gf):
        r_d = gf + J.T @ lam - zL + zU
        dl = np.where(hasL, v - Lz, 1.0)
        du = np.where(hasU, Uz - v, 1.0)
        cl = np.where(hasL, dl * zL - mu, 0.0)
        cu = np.where(hasU, du * zU - mu, 0.0)
        nz = max(1, hasL.sum() + hasU.sum())
        s_d = max(100.0, (np.abs(lam).sum() + zL.sum() + zU.sum()) / (len(lam) + nz)) / 100.0
        s_c = max(100.0, (zL.sum() + zU.sum()) / nz) / 100.0
        return max(np.abs(r_d).max() / s_d, np.abs(c).max(), np.abs(cl).max() / s_c, np.abs(cu).max() / s_c), r_d

    # least-squares multiplier initialisation (IPOPT default), dropped if large
    g, c = evaluate(v)
    J = jac(v)
    gf = grad_f(v)
    try:
        K = sp.bmat([[sp.identity(nv), J.T], [J, None]], format="csc")
        sol = spla.splu(K).solve(np.concatenate([-(gf - zL + zU), np.zeros(lay.mc)]))
        lam = sol[nv:]
        if not np.all(np.isfinite(lam)) or np.abs(lam).max() > 1e3:
            lam = np.zeros(lay.mc)
    except RuntimeError:
        lam = np.zeros(lay.mc)

    nu_pen = 1.0
    status, it = 0, 0
    delta_w_last = 0.0
    for it in range(max_iter):
        g, c = evaluate(v)
        J = jac(v)
        gf = grad_f(v)
        e0, _ = errors(v, lam, zL, zU, 0.0, J, c, gf)
        if e0 <= tol:
            status = 1
            break
        emu, r_d = errors(v, lam, zL, zU, mu, J, c, gf)
        while emu <= 10.0 * mu and mu > tol / 10.0:
            mu = max(tol / 10.0, min(0.2 * mu, mu ** 1.5))
            emu, r_d = errors(v, lam, zL, zU, mu, J, c, gf)
            nu_pen = 1.0
        tau = max(0.99, 1.0 - mu)
        dl = np.where(hasL, v - Lz, 1.0)
        du = np.where(hasU, Uz - v, 1.0)
        Sigma = np.where(hasL, zL / dl, 0.0) + np.where(hasU, zU / du, 0.0)
        W = nlp.lag_hess(d, v[:n], lam_to_g(lam), 1.0)
        W = sp.block_diag([W, sp.csr_matrix((ms, ms))], format="csr")
        gphi = gf - np.where(hasL, mu / dl, 0.0) + np.where(hasU, mu / du, 0.0)
        rhs = -np.concatenate([gphi + J.T @ lam, c])
        # regularised solve with curvature test (inertia-free, Chiang & Zavala 2016)
        delta_w = 0.0
        dv = None
        for _try in range(40):
            Hreg = W + sp.diags(Sigma + delta_w)
            K = sp.bmat([[Hreg, J.T], [J, -1e-10 * sp.identity(lay.mc)]], format="csc")
            try:
                sol = spla.splu(K).solve(rhs)
            except RuntimeError:
                sol = None
            if sol is not None and np.all(np.isfinite(sol)):
                dv_t, dlam_t = sol[:nv], sol[nv:]
                curv = dv_t @ (Hreg @ dv_t)
                if curv >= 1e-10 * (dv_t @ dv_t) or np.abs(dv_t).max() < 1e-14:
                    dv, dlam = dv_t, dlam_t
                    break
            delta_w = max(1e-4, delta_w_last / 3.0) if delta_w == 0.0 else delta_w * 8.0
        if dv is None:
            status = -7
            break
        if delta_w > 0:
            delta_w_last = delta_w
        dzL = np.where(hasL, mu / dl - zL - zL / dl * dv, 0.0)
        dzU = np.where(hasU, mu / du - zU + zU / du * dv, 0.0)

        def max_step(x, dx, mask):
            neg = mask & (dx < 0)
            if not np.any(neg):
                return 1.0
            return min(1.0, float(np.min(-tau * x[neg] / dx[neg])))

        a_p = min(max_step(dl, dv, hasL), max_step(du, -dv, hasU))
        a_d = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))
        # l1 merit
        c1 = np.abs(c).sum()
        dphi = gphi @ dv
        quad = max(0.0, dv @ ((W + sp.diags(Sigma + delta_w)) @ dv))
        if c1 > 1e-14:
            nu_need = (dphi + 0.5 * quad) / (0.9 * c1)
            if nu_pen < nu_need:
                nu_pen = nu_need + 1.0
        Dm = dphi - nu_pen * c1
        phi0 = barrier(v, mu) + nu_pen * c1
        a = a_p
        accepted = False
        for _ls in range(40):
            vt = v + a * dv
            pt = barrier(vt, mu)
            if np.isfinite(pt):
                _, ct = evaluate(vt)
                if pt + nu_pen * np.abs(ct).sum() <= phi0 + 1e-8 * a * Dm + 1e-13 * abs(phi0):
                    accepted = True
                    break
            a *= 0.5
        if not accepted:
            # tiny step: take it anyway if we are essentially converged, else give up
            if np.abs(dv).max() < 1e-9:
                a = a_p
            else:
                status = -7
                break
        v = v + a * dv
        lam = lam + a * dlam
        zL = zL + a_d * dzL
        zU = zU + a_d * dzU
        # IPOPT's kappa_sigma safeguard on the bound multipliers
        dl = np.where(hasL, v - Lz, 1.0)
        du = np.where(hasU, Uz - v, 1.0)
        zL = np.where(hasL, np.clip(zL, mu / (1e10 * dl), 1e10 * mu / dl), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (1e10 * du), 1e10 * mu / du), 0.0)
        if verbose:
            print(f"it {it:3d} mu {mu:8.1e} err {emu:9.2e} |c| {c1:9.2e} a_p {a:6.3f} a_d {a_d:6.3f} dw {delta_w:7.1e} "
                  f"f {nlp.cost(d, v[:n]):.6f}")
    w = v[:n]
    return dict(w=w, lam_g=lam_to_g(lam), z_lo=zL, z_hi=zU, iters=it, status=status,
                kkt=kkt_error(d, w, model=nlp)[0], obj=nlp.cost(d, w), mu=mu)


def kkt_error(d, w, act_tol=1e-6, model=None):
    """Algorithm-independent KKT check of a candidate primal point `w` for the reference NLP.

    Identifies the active set at tolerance `act_tol`, solves the least-squares multiplier problem with sign
    constraints (NNLS), and returns (max(stationarity, primal infeasibility, wrong-sign), details).
    Accepts points produced by ANY solver (the oracle IPM, scipy SLSQP, the CUDA kernel)."""
    from scipy.optimize import lsq_linear
    nlp = model or globals()["nlp"]
    lbg, ubg, lbx, ubx = nlp.g_bounds(d)
    g = nlp.g_fun(d, w)
    J = nlp.g_jac(d, w).toarray()
    gf = nlp.cost_grad(d, w)
    prim = max(np.max(np.maximum(lbg - g, 0)), np.max(np.maximum(g - ubg, 0)),
               np.max(np.maximum(lbx - w, 0)), np.max(np.maximum(w - ubx, 0)))
    cols, lo, hi = [], [], []
    seen_rows = set()
    for i in range(d.m):
        eq = lbg[i] == ubg[i]
        al = np.isfinite(lbg[i]) and g[i] - lbg[i] <= act_tol
        au = np.isfinite(ubg[i]) and ubg[i] - g[i] <= act_tol
        if eq or al or au:
            key = tuple(np.round(J[i], 12))
            if not eq and key in seen_rows:      # the 3x duplicated obstacle rows (Q6) share a multiplier
                continue
            seen_rows.add(key)
            cols.append(J[i])
            lo.append(-INF if (eq or al) else 0.0)
            hi.append(INF if (eq or au) else 0.0)
    for i in range(d.n):
        al = np.isfinite(lbx[i]) and w[i] - lbx[i] <= act_tol
        au = np.isfinite(ubx[i]) and ubx[i] - w[i] <= act_tol
        if al or au:
            e = np.zeros(d.n)
            e[i] = 1.0
            cols.append(e)
            lo.append(-INF if al else 0.0)
            hi.append(INF if au else 0.0)
    A = np.array(cols).T
    # gf + A y = 0 with y <= 0 for lower-active (multiplier pushes up), y >= 0 for upper-active
    res = lsq_linear(A, -gf, bounds=(np.array(lo), np.array(hi)), method="bvls", tol=1e-14)
    stat = np.abs(A @ res.x + gf).max()
    return max(stat, prim), dict(stationarity=stat, primal=prim, n_active=A.shape[1])
