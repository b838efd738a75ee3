import sys, time, random, numpy as np
sys.path.insert(0, "ba-path-planning_b200"); sys.path.insert(0, ".")
from path_planning import SCP, generate_positions
from oracle import scp_oracle
random.seed(0)
p0, pf = generate_positions(10, 0.8)
s = SCP(n_vehicles=10, time_horizon=100, time_step=0.2, min_distance=0.8, space_dims=[0, 0, 200, 200])
s.set_initial_states(p0); s.set_final_states(pf)
t0 = time.time(); tr = s.generate_trajectories(max_iterations=15); dt = time.time() - t0
r = s.last_record
print("C1 time", dt, {k: r[k] for k in ("status","scp_iterations","converged","admm_iterations","qp_unsolved","polish_ok","polish_attempts","min_separation","objective","max_copies")})
z = np.zeros((10,2))
print("dyn residual", scp_oracle.dynamics_residual(tr["accelerations"], p0, z, pf, z, 0.2, [0,0,200,200], positions=tr["positions"]), "minsep", scp_oracle.min_separation(tr["positions"]))
s.solver_settings = {"team_mode": 1}
t0 = time.time(); tr1 = s.generate_trajectories(max_iterations=15); dt1 = time.time() - t0
print("C1 (one CTA) time", dt1, "max diff vs team", np.abs(tr1["positions"]-tr["positions"]).max())
