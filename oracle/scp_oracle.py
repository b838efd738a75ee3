"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SCP path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker / the timed CPU
baseline.  The product path (ba-path-planning_b200/) never imports it.

What it restates: /root/reference/src/path_planning/solvers/scp.py, class SCP,
compute part (lines 10-615).  Same decision vector, same rows in the same
order, same bounds, same loop and stopping rule; the Python list/loop assembly
is replaced by vectorised numpy that produces the *same matrices* (pinned in
tests/test_oracle_vs_reference.py, container only, and through the committed
fixtures in tests/golden/).  The QP back-end is the osqp shim in
oracle/shims/osqp (see its header for what is and is not verifiable).

Parity status: the reference ships no tests, golden vectors or fixtures
 -- upstream parity is UNPINNED.  The pins used instead:
  (1) this restatement == the verbatim reference run with the same shim
      (matrices bit-equal, trajectories to 1e-9), checked in the container;
  (2) KKT certificates (<= 1e-9) on every subproblem of the golden runs, which
      make the golden iterates the exact minimisers independent of any solver;
  (3) SURVEY.md T1-T7 identities.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy.sparse as sp

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def _osqp():
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    import osqp  # the shim (the real package is not installable offline)

    return osqp


# --------------------------------------------------------------------------- operators
def jerk_matrix(N, K, h):
    """scp.py:10-28 -- first difference per vehicle/axis, rows (agent, k, axis), +-1/h."""
    r = np.arange(2 * N * (K - 1))
    i, rem = np.divmod(r, 2 * (K - 1))
    k, ax = np.divmod(rem, 2)
    c0 = 2 * i * K + 2 * k + ax
    rows = np.concatenate([r, r])
    cols = np.concatenate([c0, c0 + 2])
    vals = np.concatenate([np.full(r.size, -1.0 / h), np.full(r.size, 1.0 / h)])
    return sp.coo_matrix((vals, (rows, cols)), shape=(2 * N * (K - 1), 2 * N * K)).tocsc()


def integrator_blocks(K, h):
    """scp.py:198-203 (T, lower-triangular ones) and scp.py:227-232 (S[k,j] = h^2 (k-j+1/2), j<=k)."""
    k = np.arange(K)
    d = k[:, None] - k[None, :]
    T = (d >= 0).astype(float)
    S = np.where(d >= 0, h * h * (d + 0.5), 0.0)
    return T, S


class ScpOracle:
    """Mirror of reference ``SCP`` (scp.py:31-615), compute part only."""

    def __init__(self, n_vehicles=5, time_horizon=3.0, time_step=0.1, min_distance=0.1, space_dims=None):
        # scp.py:40-74
        self.N = n_vehicles
        self.T = time_horizon
        self.h = time_step
        self.K = int(self.T / self.h)
        self.R = min_distance
        self.space_dims = [0, 0, 20, 20] if space_dims is None else space_dims
        self.convergence_tolerance = 1.5e-2
        self.pos_min = np.array(self.space_dims[:2], dtype=float)
        self.pos_max = np.array(self.space_dims[2:], dtype=float)
        self.vel_min, self.vel_max = -2.0, 2.0
        self.acc_min, self.acc_max = -15.0, 15.0
        self.jerk_min, self.jerk_max = -20.0, 20.0
        self.trajectories = None
        self.record = {}

    # scp.py:99-129
    def set_initial_states(self, positions, velocities=None):
        velocities = np.zeros((self.N, 2)) if velocities is None else velocities
        self.initial_positions = np.asarray(positions, dtype=float).flatten()
        self.initial_velocities = np.asarray(velocities, dtype=float).flatten()
        assert len(self.initial_positions) == len(self.initial_velocities) == 2 * self.N

    def set_final_states(self, positions, velocities=None):
        velocities = np.zeros((self.N, 2)) if velocities is None else velocities
        self.final_positions = np.asarray(positions, dtype=float).flatten()
        self.final_velocities = np.asarray(velocities, dtype=float).flatten()
        assert len(self.final_positions) == len(self.final_velocities) == 2 * self.N

    # ------------------------------------------------------------------ constant rows
    def precompute_constraint_matrices(self):
        """scp.py:182-257.  Row k of the vel/pos blocks constrains state k+1; the last
        row (k = K-1) is the terminal equality."""
        N, K, h = self.N, self.K, self.h
        I2 = sp.eye(2, format="csc")
        IN = sp.eye(N, format="csc")
        self.C_jerk = jerk_matrix(N, K, h)
        self.l_jerk = np.full(2 * N * (K - 1), self.jerk_min)
        self.u_jerk = np.full(2 * N * (K - 1), self.jerk_max)
        self.C_acc = sp.eye(2 * N * K, format="csc")
        self.l_acc = np.full(2 * N * K, self.acc_min)
        self.u_acc = np.full(2 * N * K, self.acc_max)
        T, S = integrator_blocks(K, h)
        self.C_vel = sp.kron(IN, h * sp.kron(sp.csc_matrix(T), I2, format="csc"), format="csc")
        self.C_pos = sp.kron(IN, sp.kron(sp.csc_matrix(S), I2, format="csc"), format="csc")

        p0 = self.initial_positions.reshape(N, 2)
        v0 = self.initial_velocities.reshape(N, 2)
        pf = self.final_positions.reshape(N, 2)
        vf = self.final_velocities.reshape(N, 2)
        # velocity rows: box on v[k+1]-v0 for k<K-1, equality vf-v0 at k=K-1 (scp.py:212-224)
        lv = np.broadcast_to((self.vel_min - v0)[:, None, :], (N, K, 2)).copy()
        uv = np.broadcast_to((self.vel_max - v0)[:, None, :], (N, K, 2)).copy()
        lv[:, K - 1, :] = uv[:, K - 1, :] = vf - v0
        # position rows: offset p0 + h (k+1) v0 (scp.py:242-257)
        kk = np.arange(1, K + 1, dtype=float)[None, :, None]
        off = p0[:, None, :] + h * kk * v0[:, None, :]
        lp = self.pos_min[None, None, :] - off
        up = self.pos_max[None, None, :] - off
        lp[:, K - 1, :] = up[:, K - 1, :] = pf - off[:, K - 1, :]
        self.l_vel, self.u_vel = lv.reshape(-1), uv.reshape(-1)
        self.l_pos, self.u_pos = lp.reshape(-1), up.reshape(-1)

    def _stack_dynamics(self):
        C = sp.vstack([self.C_jerk, self.C_acc, self.C_vel, self.C_pos], format="csc")
        l = np.hstack([self.l_jerk, self.l_acc, self.l_vel, self.l_pos])  # noqa: E741
        u = np.hstack([self.u_jerk, self.u_acc, self.u_vel, self.u_pos])
        return C, l, u

    # ------------------------------------------------------------------ QP #0
    def solve_initial_trajectory(self):
        """scp.py:323-369: min sum||a||^2 (P = 2I, q = 0) over the constant rows, OSQP defaults."""
        osqp = _osqp()
        n = 2 * self.N * self.K
        P = sp.identity(n, format="csc") * 2.0
        q = np.zeros(n)
        C, l, u = self._stack_dynamics()  # noqa: E741
        prob = osqp.OSQP()
        prob.setup(P=P, q=q, A=C, l=l, u=u, verbose=False)
        res = prob.solve()
        self.record.setdefault("qp", []).append(_qp_record(res))
        if res.info.status_val not in (1, 2):
            raise RuntimeError(f"OSQP failed: {res.info.status}")
        return res.x

    # ------------------------------------------------------------------ state map
    def states_from_accelerations(self, a):
        """scp.py:371-397 and scp.py:559-595 (identical maps):
        v[k] = v0 + h sum_{j<k} a[j];  p[k] = p0 + h k v0 + h^2 sum_{j<k} (k-j-1/2) a[j]."""
        N, K, h = self.N, self.K, self.h
        a = np.asarray(a, dtype=float).reshape(N, K, 2)
        p0 = self.initial_positions.reshape(N, 2)
        v0 = self.initial_velocities.reshape(N, 2)
        c1 = np.cumsum(a, axis=1)
        c2 = np.cumsum(c1, axis=1)
        vel = np.empty((N, K, 2))
        pos = np.empty((N, K, 2))
        vel[:, 0] = v0
        pos[:, 0] = p0
        vel[:, 1:] = v0[:, None, :] + h * c1[:, :-1]
        kk = np.arange(1, K, dtype=float)[None, :, None]
        pos[:, 1:] = p0[:, None, :] + h * kk * v0[:, None, :] + h * h * (c2[:, :-1] - 0.5 * c1[:, :-1])
        return pos, vel

    # ------------------------------------------------------------------ feasibility gate
    def fast_check_avoidance(self, positions):
        """scp.py:597-615: first (k, i<j) with distance < R - 0.01 -> False."""
        N = self.N
        iu, ju = np.triu_indices(N, 1)
        d = np.linalg.norm(positions[iu] - positions[ju], axis=-1)  # (P, K)
        bad = d.T < self.R - 0.01  # (K, P) in the reference's scan order
        if bad.any():
            k, p = np.unravel_index(np.argmax(bad), bad.shape)
            self.record["first_violation"] = (int(k), int(iu[p]), int(ju[p]), float(d[p, k]))
            return False
        return True

    # ------------------------------------------------------------------ collision rows
    def collision_rows(self, a_prev):
        """scp.py:453-557.  Row order k-major then i<j lexicographic (scp.py:487-496);
        eta from the previous iterate; coefficients +eta h^2 (k-m-1/2) on agent i's
        a[m], m<k, and the negatives on agent j (scp.py:512-534);
        l = R + (eta.d - dist) - eta.(p0_i-p0_j) - eta.(v0_i-v0_j) k h (scp.py:543-550), u = +inf."""
        N, K, h, Rm = self.N, self.K, self.h, self.R
        prev, _ = self.states_from_accelerations(a_prev)
        iu, ju = np.triu_indices(N, 1)
        npairs = iu.size
        p0 = self.initial_positions.reshape(N, 2)
        v0 = self.initial_velocities.reshape(N, 2)
        diff = prev[iu] - prev[ju]  # (P, K, 2)
        diff = np.transpose(diff, (1, 0, 2))  # (K, P, 2)
        dist = np.hypot(diff[..., 0], diff[..., 1])
        eta = np.empty_like(diff)
        deg = dist < 1e-6
        safe = np.where(deg, 1.0, dist)
        eta[...] = diff / safe[..., None]
        if deg.any():  # scp.py:503-507, random direction, dist := 1.0 (row order draws)
            for k, p in zip(*np.nonzero(deg)):
                ang = np.random.uniform(0.0, 2.0 * np.pi)
                eta[k, p] = (np.cos(ang), np.sin(ang))
            dist = np.where(deg, 1.0, dist)
        lin = np.einsum("kpa,kpa->kp", eta, diff) - dist
        ipc = np.einsum("kpa,pa->kp", eta, p0[iu] - p0[ju])
        ivc = np.einsum("kpa,pa->kp", eta, v0[iu] - v0[ju]) * (np.arange(K) * h)[:, None]
        l_coll = (Rm + lin - (ipc + ivc)).reshape(-1)
        u_coll = np.full(npairs * K, np.inf)

        rows, cols, vals = [], [], []
        base_i = iu * (2 * K)
        base_j = ju * (2 * K)
        for k in range(1, K):
            m = np.arange(k)
            w = (h * h) * (k - m - 0.5)
            r = (k * npairs + np.arange(npairs))[:, None]
            rr = np.broadcast_to(r, (npairs, k))
            ex = eta[k, :, 0][:, None] * w[None, :]
            ey = eta[k, :, 1][:, None] * w[None, :]
            ci = base_i[:, None] + 2 * m[None, :]
            cj = base_j[:, None] + 2 * m[None, :]
            rows += [rr, rr, rr, rr]
            cols += [ci, ci + 1, cj, cj + 1]
            vals += [ex, ey, -ex, -ey]
        if rows:
            rows = np.concatenate([x.ravel() for x in rows])
            cols = np.concatenate([x.ravel() for x in cols])
            vals = np.concatenate([x.ravel() for x in vals])
        A = sp.coo_matrix((vals, (rows, cols)), shape=(npairs * K, 2 * N * K)).tocsc()
        self._last_eta = eta
        return A, l_coll, u_coll

    # ------------------------------------------------------------------ QP #t
    def solve_with_avoidance(self, a_prev):
        """scp.py:399-451: rows [jerk; acc; vel; pos; collision], fresh OSQP setup every
        iteration with warm_start=True, max_iter=10000, x warm start only; a failed solve
        is a printed warning and result.x is used regardless."""
        osqp = _osqp()
        n = 2 * self.N * self.K
        A_c, l_c, u_c = self.collision_rows(a_prev)
        P = sp.identity(n, format="csc") * 2.0
        q = np.zeros(n)
        C, l, u = self._stack_dynamics()  # noqa: E741
        A = sp.vstack([C, A_c], format="csc")
        prob = osqp.OSQP()
        prob.setup(P=P, q=q, A=A, l=np.hstack([l, l_c]), u=np.hstack([u, u_c]), verbose=False,
                   warm_start=True, max_iter=10000)
        prob.warm_start(x=a_prev)
        res = prob.solve()
        self.record.setdefault("qp", []).append(_qp_record(res))
        return res.x

    # ------------------------------------------------------------------ outer loop
    def generate_trajectories(self, max_iterations=15):
        """scp.py:131-180."""
        t0 = time.perf_counter()
        self.record = {"rel_steps": [], "qp": [], "iterates": []}
        self.precompute_constraint_matrices()
        a = self.solve_initial_trajectory()
        pos0, _ = self.states_from_accelerations(a)
        is_feasible = self.fast_check_avoidance(pos0)
        self.record["initial_feasible"] = bool(is_feasible)
        self.record["a_initial"] = a.copy()
        iteration, converged = 0, False
        while iteration < max_iterations and not converged and not is_feasible:
            a_new = self.solve_with_avoidance(a)
            rel = np.linalg.norm(a_new - a) / np.linalg.norm(a)
            self.record["rel_steps"].append(float(rel))
            if rel <= self.convergence_tolerance:
                converged = True
            a = a_new
            self.record["iterates"].append(a.copy())
            iteration += 1
        acc = a.reshape(self.N, self.K, 2)
        pos, vel = self.states_from_accelerations(acc)
        self.trajectories = {"positions": pos, "velocities": vel, "accelerations": acc}
        self.record.update(iterations=iteration, converged=bool(converged), time_sec=time.perf_counter() - t0)
        return self.trajectories


def _qp_record(res):
    i = res.info
    return dict(status_val=int(i.status_val), iter=int(i.iter), obj=float(i.obj_val),
                cert=float(getattr(i, "kkt_certificate", float("nan"))))


# --------------------------------------------------------------------------- checks
def min_separation(positions):
    """min over k in [0,K), i<j of ||p_i[k]-p_j[k]||; pass iff >= R - 0.01 (scp.py:610)."""
    N = positions.shape[0]
    if N < 2:
        return float("inf")
    iu, ju = np.triu_indices(N, 1)
    return float(np.linalg.norm(positions[iu] - positions[ju], axis=-1).min())


def dynamics_residual(acc, p0, v0, pf, vf, h, space_dims, positions=None,
                      vlim=2.0, alim=15.0, jlim=20.0):
    """Max violation of the constant rows of scp.py:182-257 evaluated on ``acc`` (N,K,2):
    jerk/acc/vel/pos boxes on states 1..K-1, the 4N terminal equalities on state K, and
    (optionally) ||positions - reconstruct(acc)||_inf.  Not in the reference; defined in
    SURVEY.md section 8(c) and applied identically to oracle and GPU outputs."""
    acc = np.asarray(acc, dtype=float)
    N, K, _ = acc.shape
    p0, v0, pf, vf = (np.asarray(x, dtype=float).reshape(N, 2) for x in (p0, v0, pf, vf))
    c1 = np.cumsum(acc, axis=1)
    c2 = np.cumsum(c1, axis=1)
    kk = np.arange(1, K + 1, dtype=float)[None, :, None]
    v = v0[:, None, :] + h * c1  # states 1..K
    p = p0[:, None, :] + h * kk * v0[:, None, :] + h * h * (c2 - 0.5 * c1)
    jerk = np.diff(acc, axis=1) / h
    lo = np.asarray(space_dims[:2], dtype=float)
    hi = np.asarray(space_dims[2:], dtype=float)
    viol = [
        np.max(np.abs(jerk) - jlim, initial=0.0),
        np.max(np.abs(acc) - alim, initial=0.0),
        np.max(np.abs(v[:, :-1]) - vlim, initial=0.0),
        np.max(lo - p[:, :-1], initial=0.0),
        np.max(p[:, :-1] - hi, initial=0.0),
        np.max(np.abs(v[:, -1] - vf), initial=0.0),
        np.max(np.abs(p[:, -1] - pf), initial=0.0),
    ]
    if positions is not None:
        rec = np.concatenate([p0[:, None, :], p[:, :-1]], axis=1)
        viol.append(np.max(np.abs(rec - positions), initial=0.0))
    return float(max(max(viol), 0.0))
