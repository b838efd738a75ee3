#!/usr/bin/env python
"""CPU side of BASELINE.json config 5: the restated reference (oracle/scp_oracle.py + osqp shim at OSQP's default
eps 1e-3, max_iter 10000 -- the settings scp.py:360,442 pass) on scenarios of the bench's config-5 generator
(`generate_positions_large`, seeds 10000+b, K=50), one process per scenario, timed with the reference's own
definition (perf_counter around generate_trajectories, compute_trajectories_batch.py:46-55).

    python tools/cpu_port_scaling.py N [n_scenarios] >> profiles/cpu_port_scaling_r2.jsonl

TEST / MEASUREMENT INFRASTRUCTURE: imports oracle/, never used by the product.  N >= 100 needs the explicit collision
matrix of scp.py:512-534 (24.5 M nnz at N=100, 394 M at N=200, built from Python lists): not runnable here; bench.py
extrapolates from the measured points with the fitted power law and labels the figure as an extrapolation."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ba-path-planning_b200")):
    sys.path.insert(0, p)


def main():
    N = int(sys.argv[1])
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    from oracle import scp_oracle
    from path_planning.scenarios.position_generator import generate_positions_large

    osqp = scp_oracle._osqp()
    for b in range(count):
        osqp.OVERRIDES.clear()
        osqp.STATS.clear()
        random.seed(10_000 + b)
        np.random.seed(10_000 + b)
        p0, pf, space = generate_positions_large(N, 0.8, time_horizon=10.0)
        o = scp_oracle.ScpOracle(N, 10.0, 0.2, 0.8, space)
        o.set_initial_states(p0)
        o.set_final_states(pf)
        t0 = time.perf_counter()
        status = "success"
        try:
            tr = o.generate_trajectories(max_iterations=15)
            minsep = scp_oracle.min_separation(tr["positions"])
        except Exception as e:
            status, minsep = f"error: {e}", float("nan")
        wall = time.perf_counter() - t0
        print(json.dumps(dict(n_agents=N, K=50, seed=10_000 + b, status=status, time_sec=wall,
                              scp_iterations=o.record.get("iterations", 0), min_separation=minsep,
                              qp_status=[s["status"] for s in osqp.STATS], qp_iter=[s["iter"] for s in osqp.STATS],
                              cores=1, host=os.uname().nodename, cpu_count=os.cpu_count())), flush=True)


if __name__ == "__main__":
    main()
