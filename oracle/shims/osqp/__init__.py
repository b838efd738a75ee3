"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for the PyPI ``osqp`` package (constraint ``osqp>=0.6`` in the
reference's pyproject.toml:19; unpinned, no lock file, not installable offline).
The reference's solver calls exactly this surface:

    osqp.OSQP()                                  scp.py:326, scp.py:441
    .setup(P=, q=, A=, l=, u=, verbose=False
           [, warm_start=True, max_iter=10000])  scp.py:360, scp.py:442
    .warm_start(x=...)                           scp.py:443
    .solve() -> .x, .info.status_val, .info.status   scp.py:362-367, 445-449

What is restated here is the *published* OSQP algorithm (Stellato, Banjac,
Goulart, Bemporad, Boyd: "OSQP: an operator splitting solver for quadratic
programs", Math. Prog. Comp. 2020, Algorithm 1 + sections 3.4, 4, 5): Ruiz
equilibration, the ADMM iteration with relaxation alpha, per-row step sizes
(rho, 1e3*rho on equality rows), adaptive rho, the residual based termination
test evaluated every ``check_termination`` iterations, the primal
infeasibility certificate, and "polish".  The C sources of OSQP/QDLDL are not
on this machine, so agreement with a particular OSQP release cannot be checked
here; the well-defined target of parity is the unique minimiser of each
strictly convex subproblem (objective sum ||a||^2, scp.py:328-330).

Two uses:
  * reference-like mode (defaults = OSQP defaults, eps 1e-3): the timed CPU
    baseline and the "how far is the loose reference from the minimiser" figure;
  * truth mode (``OVERRIDES`` with ``certify=True``): ADMM to moderate accuracy,
    then an active-set refinement that ends with an explicit KKT certificate
    (primal feasibility, stationarity, dual sign) at ~1e-10.  A certified
    point is the exact minimiser no matter which algorithm produced it.

Linear algebra: the x-update is solved in its reduced (normal equation) form
(P + sigma I + A' diag(rho) A) x = sigma x_k - q + A'(rho z_k - y_k), which is
algebraically identical to OSQP's quasi-definite KKT solve; dense Cholesky for
n <= DENSE_LIMIT, sparse LU of the KKT matrix beyond.
"""

from __future__ import annotations

import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

__version__ = "0.0-oracle-shim"

# Settings forced from outside (oracle harness).  Keys as in OSQP settings plus
# "certify" (bool) and "certify_tol" (float).
OVERRIDES: dict = {}
# One record per solve() call; the harness reads and clears it.
STATS: list = []

DENSE_LIMIT = 12000
OSQP_INFTY = 1e30
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_EQ_FACTOR = 1e3
RHO_TOL = 1e-4
MIN_SCALING, MAX_SCALING = 1e-4, 1e4

_DEFAULTS = dict(
    rho=0.1,
    sigma=1e-6,
    alpha=1.6,
    scaling=10,
    max_iter=4000,
    eps_abs=1e-3,
    eps_rel=1e-3,
    eps_prim_inf=1e-4,
    eps_dual_inf=1e-4,
    check_termination=25,
    adaptive_rho=True,
    adaptive_rho_interval=100,  # OSQP derives it from setup time (nondeterministic); fixed here
    adaptive_rho_tolerance=5.0,
    warm_start=True,
    polish=False,
    verbose=False,
    certify=False,
    certify_tol=1e-9,
)

_STATUS = {
    1: "solved",
    2: "solved inaccurate",
    -2: "maximum iterations reached",
    -3: "primal infeasible",
    3: "primal infeasible inaccurate",
    -4: "dual infeasible",
    -10: "unsolved",
}


class _Info:
    def __init__(self):
        self.status_val = -10
        self.status = _STATUS[-10]
        self.iter = 0
        self.obj_val = float("nan")
        self.pri_res = float("nan")
        self.dua_res = float("nan")
        self.rho_updates = 0
        self.rho_estimate = float("nan")
        self.setup_time = 0.0
        self.solve_time = 0.0
        self.run_time = 0.0
        self.status_polish = 0
        self.kkt_certificate = float("nan")  # max KKT residual when certify=True


class _Results:
    def __init__(self, n, m):
        self.x = np.full(n, np.nan)
        self.y = np.full(m, np.nan)
        self.info = _Info()


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _col_inf_norm(M):
    if M.shape[0] == 0 or M.nnz == 0:
        return np.zeros(M.shape[1])
    return np.asarray(abs(M).max(axis=0).todense()).ravel()


def _row_inf_norm(M):
    if M.shape[1] == 0 or M.nnz == 0:
        return np.zeros(M.shape[0])
    return np.asarray(abs(M).max(axis=1).todense()).ravel()


class OSQP:
    def __init__(self):
        self._is_setup = False

    # ------------------------------------------------------------------ setup
    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):  # noqa: E741
        t0 = time.perf_counter()
        s = dict(_DEFAULTS)
        unknown = set(settings) - set(s)
        if unknown:
            raise TypeError(f"unknown OSQP settings: {sorted(unknown)}")
        s.update(settings)
        s.update(OVERRIDES)
        self.s = s

        A = sp.csc_matrix(A, dtype=float)
        m, n = A.shape
        P = sp.csc_matrix(P, dtype=float) if P is not None else sp.csc_matrix((n, n))
        # OSQP takes the upper triangle of P; the reference passes a diagonal P.
        P = sp.triu(P, format="csc")
        P = (P + sp.triu(P, 1).T).tocsc()
        q = np.zeros(n) if q is None else np.asarray(q, dtype=float).copy()
        l = np.full(m, -np.inf) if l is None else np.asarray(l, dtype=float).copy()  # noqa: E741
        u = np.full(m, np.inf) if u is None else np.asarray(u, dtype=float).copy()
        if np.any(l > u):
            raise ValueError("lower bound must be lower than or equal to upper bound")
        l = np.maximum(l, -OSQP_INFTY)  # noqa: E741
        u = np.minimum(u, OSQP_INFTY)
        self.n, self.m = n, m
        self.P, self.q, self.A, self.l, self.u = P, q, A, l, u

        self._scale()
        self._rho_vectors(s["rho"])
        self._gram = None
        self._factor()

        self.x = np.zeros(n)
        self.z = np.zeros(m)
        self.y = np.zeros(m)
        self._is_setup = True
        self.setup_time = time.perf_counter() - t0

    def update_settings(self, **kw):
        self.s.update(kw)

    def warm_start(self, x=None, y=None):
        # scp.py:443 passes x only; OSQP then keeps y (zeros after setup) and
        # sets z = A x in the scaled space.
        if x is not None:
            self.x = np.asarray(x, dtype=float) / self.D
            self.z = self.As @ self.x
        if y is not None:
            self.y = np.asarray(y, dtype=float) * self.c / self.E

    # ---------------------------------------------------------------- scaling
    def _scale(self):
        n, m = self.n, self.m
        P, A, q = self.P.copy(), self.A.copy(), self.q.copy()
        D = np.ones(n)
        E = np.ones(m)
        c = 1.0
        for _ in range(int(self.s["scaling"])):
            dx = 1.0 / np.sqrt(_limit_scaling(np.maximum(_col_inf_norm(P), _col_inf_norm(A))))
            dz = 1.0 / np.sqrt(_limit_scaling(_row_inf_norm(A)))
            Dm, Em = sp.diags(dx), sp.diags(dz)
            P = (Dm @ P @ Dm).tocsc()
            A = (Em @ A @ Dm).tocsc()
            q = dx * q
            D *= dx
            E *= dz
            cn = _col_inf_norm(P)
            cost = max(float(np.mean(cn)) if n else 0.0, float(np.max(np.abs(q))) if n else 0.0)
            cost = float(_limit_scaling(np.array([cost]))[0])
            g = 1.0 / cost
            P = P * g
            q = q * g
            c *= g
        self.Ps, self.As, self.qs = P.tocsc(), A.tocsc(), q
        self.AsT = self.As.T.tocsc()
        self.D, self.E, self.c = D, E, c
        self.ls = self.E * self.l
        self.us = self.E * self.u
        # keep "infinite" bounds infinite after scaling
        self.ls[self.l <= -OSQP_INFTY] = -OSQP_INFTY * MAX_SCALING
        self.us[self.u >= OSQP_INFTY] = OSQP_INFTY * MAX_SCALING

    def _rho_vectors(self, rho):
        rho = float(min(max(rho, RHO_MIN), RHO_MAX))
        self.rho = rho
        lo_inf = self.l <= -OSQP_INFTY
        up_inf = self.u >= OSQP_INFTY
        eq = (self.us - self.ls) < RHO_TOL
        self.is_eq = eq
        self.w = np.where(eq, RHO_EQ_FACTOR, 1.0)
        rv = rho * self.w
        rv[lo_inf & up_inf] = RHO_MIN
        self.rho_vec = rv

    # ----------------------------------------------------------- factorisation
    def _factor(self):
        n, m = self.n, self.m
        sigma = self.s["sigma"]
        if n <= DENSE_LIMIT:
            loose = (self.l <= -OSQP_INFTY) & (self.u >= OSQP_INFTY)
            if self._gram is None:
                w = self.w.copy()
                w[loose] = 0.0
                self._gram = np.asarray((self.AsT @ sp.diags(w) @ self.As).todense())
                self._gram_loose = (
                    np.asarray((self.AsT @ sp.diags(loose.astype(float)) @ self.As).todense())
                    if loose.any()
                    else None
                )
                self._Pd = np.asarray(self.Ps.todense())
            K = self._Pd + self.rho * self._gram
            if self._gram_loose is not None:
                K = K + RHO_MIN * self._gram_loose
            K[np.diag_indices(n)] += sigma
            self._chol = sla.cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
            self._lu = None
        else:
            KKT = sp.bmat(
                [[self.Ps + sigma * sp.eye(n), self.AsT], [self.As, -sp.diags(1.0 / self.rho_vec)]],
                format="csc",
            )
            self._lu = spla.splu(KKT)
            self._chol = None

    def _solve_lin(self, x, z, y):
        sigma = self.s["sigma"]
        if self._chol is not None:
            rhs = sigma * x - self.qs + self.AsT @ (self.rho_vec * z - y)
            xt = sla.cho_solve(self._chol, rhs, check_finite=False)
            return xt, self.As @ xt
        rhs = np.concatenate([sigma * x - self.qs, z - y / self.rho_vec])
        sol = self._lu.solve(rhs)
        xt, nu = sol[: self.n], sol[self.n :]
        return xt, z + (nu - y) / self.rho_vec

    # ------------------------------------------------------------------ solve
    def solve(self):
        if not self._is_setup:
            raise RuntimeError("setup() first")
        t0 = time.perf_counter()
        s = self.s
        n, m = self.n, self.m
        res = _Results(n, m)
        info = res.info
        alpha = s["alpha"]
        if not s["warm_start"]:
            self.x[:] = 0
            self.z[:] = 0
            self.y[:] = 0
        x, z, y = self.x.copy(), self.z.copy(), self.y.copy()
        status = -2
        it = 0
        rho_updates = 0
        pri = dua = np.inf
        Einv, Dinv = 1.0 / self.E, 1.0 / self.D
        for it in range(1, int(s["max_iter"]) + 1):
            x_prev, y_prev = x, y
            xt, zt = self._solve_lin(x, z, y)
            x = alpha * xt + (1 - alpha) * x_prev
            zr = alpha * zt + (1 - alpha) * z
            z_new = np.clip(zr + y / self.rho_vec, self.ls, self.us)
            y = y + self.rho_vec * (zr - z_new)
            z = z_new

            check = s["check_termination"] and it % s["check_termination"] == 0
            adapt = s["adaptive_rho"] and s["adaptive_rho_interval"] and it % s["adaptive_rho_interval"] == 0
            if not (check or adapt or it == s["max_iter"]):
                continue
            Ax = self.As @ x
            Px = self.Ps @ x
            Aty = self.AsT @ y
            pri = np.max(np.abs(Einv * (Ax - z))) if m else 0.0
            dua = np.max(np.abs(Dinv * (Px + self.qs + Aty))) / self.c
            npri = max(np.max(np.abs(Einv * Ax)), np.max(np.abs(Einv * z))) if m else 0.0
            ndua = max(np.max(np.abs(Dinv * Px)), np.max(np.abs(Dinv * Aty)), np.max(np.abs(Dinv * self.qs))) / self.c
            if check or it == s["max_iter"]:
                if pri <= s["eps_abs"] + s["eps_rel"] * npri and dua <= s["eps_abs"] + s["eps_rel"] * ndua:
                    status = 1
                    break
                if m and self._primal_infeasible(y - y_prev):
                    status = -3
                    break
                if self._dual_infeasible(x - x_prev):
                    status = -4
                    break
            if adapt:
                # rho estimate from scaled residuals (paper section 5.2)
                Axs, zs = np.max(np.abs(Ax)), np.max(np.abs(z))
                pr_s = np.max(np.abs(Ax - z)) / (max(Axs, zs) + 1e-10)
                du_s = np.max(np.abs(Px + self.qs + Aty)) / (
                    max(np.max(np.abs(Px)), np.max(np.abs(Aty)), np.max(np.abs(self.qs))) + 1e-10
                )
                rho_new = self.rho * np.sqrt(pr_s / (du_s + 1e-10))
                rho_new = float(min(max(rho_new, RHO_MIN), RHO_MAX))
                info.rho_estimate = rho_new
                tol = s["adaptive_rho_tolerance"]
                if rho_new > self.rho * tol or rho_new < self.rho / tol:
                    self._rho_vectors(rho_new)
                    self._factor()
                    rho_updates += 1

        if status == -2 and pri <= 10 * (s["eps_abs"] + s["eps_rel"] * npri) and dua <= 10 * (s["eps_abs"] + s["eps_rel"] * ndua):
            status = 2  # "solved inaccurate": within 10x of the tolerances at max_iter
        self.x, self.z, self.y = x, z, y

        xu = self.D * x
        yu = self.E * y / self.c
        info.iter = it
        info.pri_res, info.dua_res = float(pri), float(dua)
        info.rho_updates = rho_updates
        if status in (1, 2, -2):
            res.x, res.y = xu, yu
        if status in (1, 2) and (s["polish"] or s["certify"]):
            ok, xp, yp, cert = self._refine(xu, yu, s["certify_tol"])
            if not ok and s["certify"] and self.n <= 1500:
                # truth mode, small problem: the refinement cycled from the ADMM's guess of the active set (several box
                # rows binding); take the guess from the exact active-set QP solver scipy vendors and certify THAT
                try:
                    _, xh, yh = solve_qp_highs(self.P, self.q, self.A, self.l, self.u)
                    ok, xp, yp, cert = self._refine(xh, yh, s["certify_tol"])
                except Exception:
                    ok = False
            info.status_polish = 1 if ok else -1
            info.kkt_certificate = cert
            if ok:
                res.x, res.y = xp, yp
            elif s["certify"]:
                status = 2
        if status in (1, 2, -2):
            info.obj_val = float(0.5 * res.x @ (self.P @ res.x) + self.q @ res.x)
        info.status_val = status
        info.status = _STATUS[status]
        info.setup_time = self.setup_time
        info.solve_time = time.perf_counter() - t0
        info.run_time = info.setup_time + info.solve_time
        STATS.append(
            dict(n=n, m=m, iter=it, status=status, pri=float(pri), dua=float(dua), rho=self.rho,
                 rho_updates=rho_updates, setup_s=self.setup_time, solve_s=info.solve_time,
                 polish=info.status_polish, cert=info.kkt_certificate,
                 refine_rounds=getattr(self, "_refine_rounds", 0), n_active=getattr(self, "_refine_nact", 0))
        )
        return res

    # ------------------------------------------------- infeasibility certificates
    def _primal_infeasible(self, dy):
        # Paper section 3.4: dy certifies primal infeasibility when A'dy = 0 and
        # u'(dy)+ + l'(dy)- < 0.  dy is first projected on the polar of the
        # recession cone of [l, u] (no positive part on rows with u = +inf, no
        # negative part on rows with l = -inf).
        eps = self.s["eps_prim_inf"]
        up_inf = self.u >= OSQP_INFTY
        lo_inf = self.l <= -OSQP_INFTY
        dy = np.where(up_inf, np.minimum(dy, 0.0), dy)
        dy = np.where(lo_inf, np.maximum(dy, 0.0), dy)
        nrm = np.max(np.abs(self.E * dy))
        if nrm <= 1e-10:
            return False
        us = np.where(up_inf, 0.0, self.us)
        ls = np.where(lo_inf, 0.0, self.ls)
        support = us @ np.maximum(dy, 0) + ls @ np.minimum(dy, 0)
        if support >= -eps * nrm:
            return False
        return np.max(np.abs((self.AsT @ dy) / self.D)) <= eps * nrm

    def _dual_infeasible(self, dx):
        eps = self.s["eps_dual_inf"]
        dxu = self.D * dx
        nrm = np.max(np.abs(dxu))
        if nrm <= 1e-12:
            return False
        if self.qs @ dx / self.c >= -eps * nrm:
            return False
        if np.max(np.abs((self.Ps @ dx) / self.D)) / self.c > eps * nrm:
            return False
        Adx = (self.As @ dx) / self.E
        up_ok = np.all((self.u >= OSQP_INFTY) | (Adx <= eps * nrm))
        lo_ok = np.all((self.l <= -OSQP_INFTY) | (Adx >= -eps * nrm))
        return bool(up_ok and lo_ok)

    # ------------------------------------------------- polish / KKT certificate
    def _refine(self, x, y, tol, max_rounds=60):
        """Primal-dual active-set refinement in the unscaled problem.

        Start from OSQP's polish guess of the active rows (paper section 4:
        lower-active where z_i - l_i < -y_i, upper-active where u_i - z_i < y_i),
        solve the equality constrained QP on that set exactly, then add violated
        rows / drop rows whose multiplier has the wrong sign until the KKT
        conditions hold.  Returns (ok, x, y, max_kkt_residual).
        """
        P, A, q, l, u = self.P, self.A, self.q, self.l, self.u  # noqa: E741
        n, m = self.n, self.m
        pd = P.diagonal()
        diagP = (P - sp.diags(pd)).nnz == 0 and np.all(pd > 0)
        Ax = A @ x
        eq = (u - l) < 1e-12
        low = ((Ax - l) < -y) & ~eq
        upp = ((u - Ax) < y) & ~eq
        Acsr = A.tocsr()
        scale_p = 1.0 + max(np.max(np.abs(Ax)), 1.0)
        best = (np.inf, x, y)
        self._refine_rounds = 0
        self._refine_nact = 0
        for _ in range(max_rounds):
            self._refine_rounds += 1
            act = np.flatnonzero(eq | low | upp)
            b = np.where(upp[act], u[act], l[act])
            Aa = Acsr[act]
            self._refine_nact = len(act)
            if diagP:
                Aad = Aa.toarray()
                G = (Aad / pd) @ Aad.T
                rhs = -(b + Aad @ (q / pd))
                # active rows may be linearly dependent: minimum-norm multiplier
                try:
                    ya = sla.lstsq(G, rhs, cond=1e-13, lapack_driver="gelsd", check_finite=False)[0]
                except sla.LinAlgError:
                    return False, x, y, np.inf
                xn = -(q + Aad.T @ ya) / pd
                # one step of refinement on the primal equations
                r = b - Aad @ xn
                if np.max(np.abs(r), initial=0.0) > 1e-13 * scale_p:
                    dya = sla.lstsq(G, -r, cond=1e-13, lapack_driver="gelsd", check_finite=False)[0]
                    ya = ya + dya
                    xn = -(q + Aad.T @ ya) / pd
            else:
                na = len(act)
                KK = sp.bmat([[P + 1e-10 * sp.eye(n), Aa.T], [Aa, -1e-10 * sp.eye(na)]], format="csc")
                sol = spla.splu(KK).solve(np.concatenate([-q, b]))
                xn, ya = sol[:n], sol[n:]
            yn = np.zeros(m)
            yn[act] = ya
            Axn = A @ xn
            viol_l = (l - Axn) > tol * scale_p
            viol_u = (Axn - u) > tol * scale_p
            scale_d = 1.0 + np.max(np.abs(yn), initial=0.0)
            bad_low = low & (yn > tol * scale_d)
            bad_upp = upp & (yn < -tol * scale_d)
            stat = np.max(np.abs(P @ xn + q + A.T @ yn), initial=0.0)
            prim = max(np.max(l - Axn, initial=0.0), np.max(Axn - u, initial=0.0), 0.0)
            sign = max(np.max(np.where(low, yn, 0.0), initial=0.0), np.max(np.where(upp, -yn, 0.0), initial=0.0))
            cert = max(stat / scale_d, prim / scale_p, sign / scale_d)
            if cert < best[0]:
                best = (cert, xn, yn)
            changed = viol_l.any() or viol_u.any() or bad_low.any() or bad_upp.any()
            if not changed:
                return cert <= 10 * tol, xn, yn, float(cert)
            low = (low | (viol_l & ~eq)) & ~bad_low
            upp = (upp | (viol_u & ~eq)) & ~bad_upp
            both = low & upp
            upp &= ~both
        return False, best[1], best[2], float(best[0])


# ---------------------------------------------------------------------------
# Exact QP through the HiGHS active-set QP solver that scipy vendors
# (scipy.optimize._highspy; private API, scipy 1.18.1).  Cross-check for tiny
# cases only; it is slow.
def solve_qp_highs(P, q, A, l, u):  # noqa: E741
    import scipy.optimize._highspy._core as hs

    A = sp.csc_matrix(A, dtype=float)
    m, n = A.shape
    Pl = sp.tril(sp.csc_matrix(P, dtype=float), format="csc")
    inf = hs.kHighsInf
    lp = hs.HighsLp()
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = np.asarray(q, dtype=float)
    lp.col_lower_ = np.full(n, -inf)
    lp.col_upper_ = np.full(n, inf)
    lp.row_lower_ = np.where(np.isfinite(l) & (np.abs(l) < OSQP_INFTY), l, -inf)
    lp.row_upper_ = np.where(np.isfinite(u) & (np.abs(u) < OSQP_INFTY), u, inf)
    lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data
    hess = hs.HighsHessian()
    hess.dim_ = n
    hess.format_ = hs.HessianFormat.kTriangular
    hess.start_ = Pl.indptr.astype(np.int32)
    hess.index_ = Pl.indices.astype(np.int32)
    hess.value_ = Pl.data
    H = hs._Highs()
    H.setOptionValue("output_flag", False)
    model = hs.HighsModel()
    model.lp_ = lp
    model.hessian_ = hess
    H.passModel(model)
    H.run()
    st = H.getModelStatus()
    sol = H.getSolution()
    return st, np.array(sol.col_value), -np.array(sol.row_dual)
