"""Numpy model of the device algorithm (same arithmetic as csrc/, readable form).

NOT the oracle and NOT on the product path: a development/test aid that mirrors
what the CUDA kernels compute (copy-split ADMM with one shared K x K operator),
so kernel intermediates can be compared step by step on CPU.
See DESIGN.md section "QP solver" for the derivation.
"""

from __future__ import annotations

import numpy as np


class Params:
    def __init__(self, **kw):
        self.rho = 1.0          # global multiplier on every row weight
        self.w_jerk = 1.0
        self.w_acc = 1.0
        self.w_vel = 1.0
        self.w_pos = 1.0
        self.w_col = 1.0
        self.sigma = 1e-6
        self.row_normalise = True
        self.copies = None      # None -> N-1 (exact copy splitting)
        self.eps_abs = 1e-6
        self.eps_rel = 1e-6
        self.check_every = 25
        self.max_iter = 20000
        self.adaptive_rho = True
        self.adapt_every = 100
        self.adapt_tol = 5.0
        self.vlim, self.alim, self.jlim = 2.0, 15.0, 20.0
        self.__dict__.update(kw)


def operators(K, h):
    k = np.arange(K)
    d = k[:, None] - k[None, :]
    T = (d >= 0).astype(float)
    S = np.where(d >= 0, h * h * (d + 0.5), 0.0)
    D = np.zeros((K - 1, K))
    D[k[:-1], k[:-1]] = -1.0 / h
    D[k[:-1], k[:-1] + 1] = 1.0 / h
    return D, h * T, S


class SharedFactor:
    """The K x K operator shared by every agent/axis of a scenario (T5 in SURVEY.md)."""

    def __init__(self, K, h, copies, prm: Params):
        D, V, S = operators(K, h)
        self.D, self.V, self.S = D, V, S
        if prm.row_normalise == "inf":
            rn = lambda A: 1.0 / np.max(np.abs(A), axis=1)  # noqa: E731
        elif prm.row_normalise:
            rn = lambda A: 1.0 / np.sum(A * A, axis=1)  # noqa: E731
        else:
            rn = lambda A: np.ones(A.shape[0])  # noqa: E731
        self.r_j = prm.rho * prm.w_jerk * rn(D)
        self.r_a = prm.rho * prm.w_acc * np.ones(K)
        self.r_v = prm.rho * prm.w_vel * rn(V)[: K - 1]
        self.r_p = prm.rho * prm.w_pos * rn(S)[: K - 1]
        self.r_c = prm.rho * prm.w_col * rn(S)[: K - 1]  # per copy, rows 0..K-2 <-> positions 1..K-1
        self.copies = copies
        Vb, Sb = V[: K - 1], S[: K - 1]
        M = (2.0 + prm.sigma) * np.eye(K) + D.T @ (self.r_j[:, None] * D) + np.diag(self.r_a)
        M += Vb.T @ (self.r_v[:, None] * Vb) + Sb.T @ ((self.r_p + copies * self.r_c)[:, None] * Sb)
        self.M = M
        C = np.stack([V[K - 1], S[K - 1]])  # terminal equalities: v[K], p[K]
        Mi = np.linalg.inv(M)
        G = np.linalg.inv(C @ Mi @ C.T)
        self.Q = G @ C @ Mi                 # mu = Q r - G d
        self.G = G
        self.Nmat = Mi - Mi @ C.T @ self.Q  # x = Nmat r + N0 d
        self.N0 = Mi @ C.T @ G
        self.C = C


def solve_qp(p0, v0, pf, vf, h, K, R, space, eta=None, x0=None, prm: Params | None = None, lam0=None):
    """min sum||a||^2 s.t. the reference rows (scp.py:182-257) and, when eta (K,N,N,2)
    is given, eta_ij[k].(p_i[k]-p_j[k]) >= R for k=1..K-1, i<j (scp.py:487-552).
    Returns x (N,K,2) and an info dict."""
    prm = prm or Params()
    N = p0.shape[0]
    copies = (N - 1) if prm.copies is None else prm.copies
    have_col = eta is not None and N > 1
    if not have_col:
        copies = 0
    F = SharedFactor(K, h, copies, prm)
    D, V, S = F.D, F.V, F.S
    lo, hi = np.asarray(space[:2], float), np.asarray(space[2:], float)
    kk = np.arange(1, K + 1, dtype=float)[None, :, None]
    off = p0[:, None, :] + h * kk * v0[:, None, :]            # (N,K,2) offset of pos row k (state k+1)
    l_v = (-prm.vlim - v0)[:, None, :] * np.ones((1, K - 1, 1))
    u_v = (prm.vlim - v0)[:, None, :] * np.ones((1, K - 1, 1))
    l_p = lo[None, None, :] - off[:, : K - 1]
    u_p = hi[None, None, :] - off[:, : K - 1]
    d_eq = np.stack([vf - v0, pf - off[:, K - 1]], axis=1)      # (N,2[v,p],2[axis])

    x = np.zeros((N, K, 2)) if x0 is None else x0.reshape(N, K, 2).copy()
    # row values A x and v = z + y/rho (y = 0 -> v = A x)
    def rows(x):
        return (np.einsum("rk,nkc->nrc", D, x), x, np.einsum("rk,nkc->nrc", V[: K - 1], x),
                np.einsum("rk,nkc->nrc", S[: K - 1], x))
    vj, va, vv, vp = (r.copy() for r in rows(x))
    lam = np.zeros((K - 1, N, N)) if lam0 is None else lam0.copy()   # upper triangle used
    iu, ju = np.triu_indices(N, 1)
    rj, ra, rv, rp, rc = (F.r_j[None, :, None], F.r_a[None, :, None], F.r_v[None, :, None],
                          F.r_p[None, :, None], F.r_c[None, :, None])
    info = dict(iters=0, status=0, rho_updates=0)
    force = np.zeros((N, K - 1, 2))
    pos_rows = rows(x)[3]
    it = 0
    rho_scale = 1.0
    for it in range(1, prm.max_iter + 1):
        zj, za = np.clip(vj, -prm.jlim, prm.jlim), np.clip(va, -prm.alim, prm.alim)
        zv, zp = np.clip(vv, l_v, u_v), np.clip(vp, l_p, u_p)
        wj, wa, wv, wp = 2 * zj - vj, 2 * za - va, 2 * zv - vv, 2 * zp - vp
        rhs = prm.sigma * x + np.einsum("rk,nrc->nkc", D, rj * wj) + ra * wa
        rhs += np.einsum("rk,nrc->nkc", V[: K - 1], rv * wv)
        rhs += np.einsum("rk,nrc->nkc", S[: K - 1], rp * wp + copies * rc * pos_rows + force)
        x_new = np.einsum("kl,nlc->nkc", F.Nmat, rhs) + np.einsum("ke,nec->nkc", F.N0, d_eq)
        mu = np.einsum("ek,nkc->nec", F.Q, rhs) - np.einsum("ef,nfc->nec", F.G, d_eq)
        aj, aa, av, ap = rows(x_new)
        # v' = A x' + (v - z)
        vj, va, vv, vp = aj + (vj - zj), aa + (va - za), av + (vv - zv), ap + (vp - zp)
        pos_rows = ap
        dx = x_new - x
        x = x_new
        pr_col = 0.0
        if have_col:
            p = off[:, : K - 1] + ap                                  # positions 1..K-1, (N,K-1,2)
            g = np.einsum("kpc,pkc->kp", eta[1:, iu, ju], p[iu] - p[ju])  # (K-1,P)
            rck = F.r_c[:, None]
            lam_old = lam[:, iu, ju]
            lam_new = np.maximum(0.0, lam_old + 0.5 * rck * (R - g))
            lam[:, iu, ju] = lam_new
            fpair = (2 * lam_new - lam_old)[..., None] * eta[1:, iu, ju]  # (K-1,P,2) on i, minus on j
            force = np.zeros((N, K - 1, 2))
            np.add.at(force, iu, np.transpose(fpair, (1, 0, 2)))
            np.add.at(force, ju, -np.transpose(fpair, (1, 0, 2)))
            pr_col = np.max(np.abs(lam_new - lam_old) / rck) if lam_new.size else 0.0
        if it % prm.check_every and it != prm.max_iter:
            continue
        # ---- residuals (unscaled, reference units)
        zj2, za2 = np.clip(vj, -prm.jlim, prm.jlim), np.clip(va, -prm.alim, prm.alim)
        zv2, zp2 = np.clip(vv, l_v, u_v), np.clip(vp, l_p, u_p)
        pri = max(np.max(np.abs(aj - zj2)), np.max(np.abs(aa - za2)), np.max(np.abs(av - zv2)),
                  np.max(np.abs(ap - zp2)), pr_col)
        yj, ya, yv, yp = rj * (vj - zj2), ra * (va - za2), rv * (vv - zv2), rp * (vp - zp2)
        Aty = np.einsum("rk,nrc->nkc", D, yj) + ya + np.einsum("rk,nrc->nkc", V[: K - 1], yv)
        ycol = np.zeros((N, K - 1, 2))
        if have_col:
            fy = lam[:, iu, ju][..., None] * eta[1:, iu, ju]
            np.add.at(ycol, iu, -np.transpose(fy, (1, 0, 2)))
            np.add.at(ycol, ju, np.transpose(fy, (1, 0, 2)))
        Aty += np.einsum("rk,nrc->nkc", S[: K - 1], yp + ycol)
        Ctmu = np.einsum("ek,nec->nkc", F.C, mu)
        dua = np.max(np.abs(2 * x + Aty + Ctmu))
        n_pri = max(np.max(np.abs(aj)), np.max(np.abs(aa)), np.max(np.abs(av)), np.max(np.abs(ap)))
        n_dua = max(np.max(np.abs(2 * x)), np.max(np.abs(Aty + Ctmu)))
        info.update(iters=it, pri=float(pri), dua=float(dua))
        if pri <= prm.eps_abs + prm.eps_rel * n_pri and dua <= prm.eps_abs + prm.eps_rel * n_dua:
            info["status"] = 1
            break
        if prm.adaptive_rho and it % prm.adapt_every == 0:
            est = np.sqrt((pri / max(n_pri, 1e-12)) / max(dua / max(n_dua, 1e-12), 1e-12))
            if est > prm.adapt_tol or est < 1.0 / prm.adapt_tol:
                est = float(np.clip(est, 1e-3, 1e3))
                prm2 = Params(**prm.__dict__)
                prm2.rho = prm.rho * est
                prm = prm2
                F = SharedFactor(K, h, copies, prm)
                # keep y: v = z + y/rho -> v = z + (v - z)/est
                vj, va = zj2 + (vj - zj2) / est, za2 + (va - za2) / est
                vv, vp = zv2 + (vv - zv2) / est, zp2 + (vp - zp2) / est
                rj, ra, rv, rp, rc = (F.r_j[None, :, None], F.r_a[None, :, None], F.r_v[None, :, None],
                                      F.r_p[None, :, None], F.r_c[None, :, None])
                # the collision part of the next rhs was formed with the old (2 lam' - lam); keep it
                info["rho_updates"] += 1
    info["iters"] = it
    info["rho"] = prm.rho
    info["lam"] = lam
    return x, info
