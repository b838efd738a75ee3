"""``SCP`` -- same class, constructor, setters and ``generate_trajectories`` contract as the
reference (src/path_planning/solvers/scp.py:31-180); the compute runs on the GPU.

What stays on the host: argument checks, the prints the reference makes, the result dict.
What moved to the device (one call, no host round trip inside): the constant operators
(:182-257), the initial QP (:323-369), state reconstruction (:371-397, :559-595), the
feasibility gate (:597-615), collision linearisation (:453-557), every per-iteration QP
(:399-451) and the convergence loop (:152-166).
"""

from __future__ import annotations

import ctypes as C
import time

import numpy as np

from .. import _capi


def use_stream_engine(engine, n_agents, n_steps, n_scenarios=1):
    """"auto": ONE (or a few) scenarios whose state cannot live in one CTA's shared memory (2NK doubles per array;
    the one-CTA solver then works out of L2 and the whole-grid kernel pays grid barriers) go to the streaming
    solver, which needs K <= 128 (measured: 200 agents, K=100: 2.5 s against 12 s).  Batches that fill the GPU with
    one CTA per scenario stay on the one-CTA solver -- its polish needs ~5x fewer ADMM iterations (measured at
    148 x 50 and 148 x 100 agents: 765 / 129 against 88 / 44 scenarios/s) -- and so do config 1 (K=500) and small
    scenarios."""
    if engine == "stream":
        return True
    if engine == "cta":
        return False
    return n_steps <= 128 and 2 * n_agents * n_steps >= 8192 and n_scenarios <= 8


class SCP:
    def __init__(self, n_vehicles=5, time_horizon=3.0, time_step=0.1, min_distance=0.1, space_dims=None):
        self.N = n_vehicles
        self.T = time_horizon
        self.h = time_step
        self.K = int(self.T / self.h)
        self.R = min_distance
        if space_dims is None:
            space_dims = [0, 0, 20, 20]
        self.space_dims = space_dims
        self.convergence_tolerance = 1.5e-2
        self.trajectories = None
        self.initial_positions = None
        self.initial_velocities = None
        self.final_positions = None
        self.final_velocities = None
        self.pos_min = np.array([space_dims[0], space_dims[1]])
        self.pos_max = np.array([space_dims[2], space_dims[3]])
        self.vel_min, self.vel_max = -2, 2
        self.acc_min, self.acc_max = -15.0, 15.0
        self.jerk_min, self.jerk_max = -20, 20
        # new, additive: solver settings (fields of scp_b200_problem), device index, last record
        self.solver_settings = {}
        self.engine = "auto"   # "cta": one CTA / whole grid per scenario; "stream": streaming multi-kernel solver; "auto"
        self.device = 0
        self.verbose = True
        self.last_record = None
        self._say("---=== SCP Problem initialized ===---")
        self._say(f"Number of timesteps: {self.K}")
        self._say(f"Timestep: {self.h}")
        self._say(f"Minimum distance between vehicles: {self.R}")
        self._say(f"Space dimensions: {self.space_dims}")

    def _say(self, *a):
        if self.verbose:
            print(*a)

    def set_initial_states(self, positions, velocities=None):
        if velocities is None:
            velocities = np.zeros((self.N, 2))
        self.initial_positions = np.asarray(positions, dtype=float).flatten()
        self.initial_velocities = np.asarray(velocities, dtype=float).flatten()
        assert len(self.initial_positions) == len(self.initial_velocities) == 2 * self.N, (
            f"Initial states mismatch positions={len(self.initial_positions)}, "
            f"velocities={len(self.initial_velocities)}, expected={2*self.N}"
        )

    def set_final_states(self, positions, velocities=None):
        if velocities is None:
            velocities = np.zeros((self.N, 2))
        self.final_positions = np.asarray(positions, dtype=float).flatten()
        self.final_velocities = np.asarray(velocities, dtype=float).flatten()
        assert len(self.final_positions) == len(self.final_velocities) == 2 * self.N, (
            f"Final states mismatch positions={len(self.final_positions)}, "
            f"velocities={len(self.final_velocities)}, expected={2*self.N}"
        )

    def _problem(self, max_iterations):
        p = _capi.default_problem(self.N, self.T, self.h, self.R, self.space_dims)
        p.n_steps = self.K
        p.vel_limit = float(self.vel_max)
        p.acc_limit = float(self.acc_max)
        p.jerk_limit = float(self.jerk_max)
        p.scp_tolerance = float(self.convergence_tolerance)
        p.max_scp_iter = int(max_iterations)
        for k, v in self.solver_settings.items():
            setattr(p, k, v)
        return p

    def generate_trajectories(self, max_iterations=15):
        """Main method to generate collision-free trajectories using SCP (device side)."""
        lib = _capi.load()
        start = time.time()
        if max_iterations > _capi.MAX_SCP_ITER:
            raise ValueError(f"max_iterations > {_capi.MAX_SCP_ITER} not supported")
        if any(a is None for a in (self.initial_positions, self.initial_velocities, self.final_positions,
                                   self.final_velocities)):
            raise ValueError("set_initial_states() and set_final_states() must be called before generate_trajectories()")
        p = self._problem(max_iterations)
        N, K = self.N, self.K
        buf = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        p0, v0 = buf(self.initial_positions), buf(self.initial_velocities)
        pf, vf = buf(self.final_positions), buf(self.final_velocities)
        acc = np.empty((N, K, 2))
        pos = np.empty((N, K, 2))
        vel = np.empty((N, K, 2))
        rec = _capi.Record()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        if use_stream_engine(self.engine, N, K):
            r = self._solve_streaming(max_iterations, p0, v0, pf, vf, acc, pos, vel)
        else:
            _capi.check(lib.scp_b200_solve_batch_host(C.byref(p), 1, ptr(p0), ptr(v0), ptr(pf), ptr(vf), ptr(acc),
                                                      ptr(pos), ptr(vel), C.cast(C.byref(rec), C.c_void_p),
                                                      int(self.device)))
            r = _capi.record_to_dict(rec)
        self.last_record = r
        if r["status"] == _capi.STATUS_INITIAL_QP_FAILED:
            print("not feasible")
            raise RuntimeError("OSQP failed: initial QP not solved (device ADMM hit its iteration limit)")
        if not r["initial_feasible"]:
            k, i, j = r["first_violation"]
            self._say(f"Avoidance constraint violation at timestep {k} between vehicles {i} and {j}: "
                      f"distance = {r['first_violation_dist']:.3f}")
        for it, rel in enumerate(r["rel_steps"]):
            self._say(f"SCP Iteration {it+1}")
            self._say(rel)
            if rel <= self.convergence_tolerance:
                self._say(f"Converged after {it+1} iterations.")
        if r["qp_unsolved"]:
            self._say(f"Warning: {r['qp_unsolved']} subproblem(s) hit the ADMM iteration limit")
        if r.get("candidate_overflow"):
            print("Warning: collision rows were dropped (candidate capacity of the streaming solver exceeded)")
        self.trajectories = {"positions": pos, "velocities": vel, "accelerations": acc}
        self._say(f"Trajectory generation completed in {time.time() - start:.3f} seconds")
        return self.trajectories

    def _solve_streaming(self, max_iterations, p0, v0, pf, vf, acc, pos, vel):
        """Large scenarios: the streaming solver (include/scp_b200.h, scp_b200_stream_*), host buffers in and out."""
        import torch

        from .stream import StreamSolver

        with torch.cuda.device(int(self.device)):
            settings = dict(self.solver_settings, max_scp_iter=int(max_iterations), vel_limit=float(self.vel_max),
                            acc_limit=float(self.acc_max), jerk_limit=float(self.jerk_max),
                            scp_tolerance=float(self.convergence_tolerance))
            s = StreamSolver(self.N, self.T, self.h, self.R, self.space_dims, n_scenarios=1, **settings)
            traj, recs = s.solve(p0.reshape(1, self.N, 2), pf.reshape(1, self.N, 2), v0.reshape(1, self.N, 2),
                                 vf.reshape(1, self.N, 2))
            s.close()
        acc[...] = traj["accelerations"][0]
        pos[...] = traj["positions"][0]
        vel[...] = traj["velocities"][0]
        r = recs[0]
        r["device_ms"] = s.last_device_ms
        return r

    # Plotting is out of scope for acceleration; thin pass-throughs keep the entry points alive.
    def visualize_trajectories(self, show_animation=False, save_path="trajectories.pdf"):
        if self.trajectories is None:
            raise ValueError("Trajectories not generated yet")
        from ..viz import _plots

        return _plots.trajectories(self, show_animation, save_path)

    def visualize_time_snapshots(self, num_snapshots=5, save_path=None):
        if self.trajectories is None:
            raise ValueError("Trajectories not generated yet")
        from ..viz import _plots

        return _plots.snapshots(self, num_snapshots, save_path)
