"""ORACLE / TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
For every (N, seed) case it runs the VERBATIM reference scp.py (loaded by
oracle/ref_loader.py with the osqp/matplotlib shims) in truth mode -- ADMM to
1e-5 then active-set refinement with a KKT certificate -- and, beside it, the
numpy restatement oracle/scp_oracle.py; the two must agree to 1e-9 and every
subproblem must carry a certificate <= 1e-9, otherwise the case is rejected.
Stored: inputs, the certified iterates/outputs, and the same scenario solved at
OSQP's default eps 1e-3 (how far the loose reference itself sits from the minimiser).
"""

from __future__ import annotations

import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, scp_oracle  # noqa: E402

CASES = [  # name, N, T, h, R, space, seed (None: explicit positions)
    ("n5_s0", 5, 10.0, 0.2, 0.8, [0, 0, 20, 20], 0),
    ("n5_s1", 5, 10.0, 0.2, 0.8, [0, 0, 20, 20], 1),
    ("n8_s1", 8, 10.0, 0.2, 0.8, [0, 0, 20, 20], 1),
    ("n10_s1", 10, 10.0, 0.2, 0.8, [0, 0, 20, 20], 1),
    ("n10_s2", 10, 10.0, 0.2, 0.8, [0, 0, 20, 20], 2),
    ("n15_s2", 15, 10.0, 0.2, 0.8, [0, 0, 20, 20], 2),
    ("n25_s3", 25, 10.0, 0.2, 0.8, [0, 0, 20, 20], 3),
    ("n25_s10003", 25, 10.0, 0.2, 0.8, [0, 0, 20, 20], 10003),
    # round 2: BASELINE.json config 1 sized (compute-trajectories defaults: 10 agents, T=100, h=0.2 -> K=500, 200 x 200 m)
    ("n10_k500_s0", 10, 100.0, 0.2, 0.8, [0, 0, 200, 200], 0),
    # 50 agents (config-5 size); the first seed whose subproblems all certify
    ("n50_s1", 50, 10.0, 0.2, 0.8, [0, 0, 20, 20], 1),
]
# Cases in which a BOX row class binds at the optimum (the reference's limits are plain attributes set in
# SCP.__init__, scp.py:67-74, read by _precompute_constraint_matrices, scp.py:188-257): name -> overrides.
#   acc / jerk: limits lowered until the manoeuvre saturates them; pos: arena tightened around the straight paths so
#   that the evasive detour hits the wall.  (Two-agent cases: with several binding box rows on five agents the shim's
#   active-set refinement cycles and no certificate comes out.)
EXTRA = {
    # two agents swapping places along y = 1.0 / 1.1 in a corridor: the lower wall (0.8) stops agent 0's detour after
    # 0.2 m, so its position rows bind while agent 1 takes the rest of the 0.8 m separation
    "n2_corridor_pos": dict(N=2, T=10.0, h=0.2, R=0.8, space=[0.0, 0.8, 10.0, 1.9],
                            p0=[[1.0, 1.0], [9.0, 1.1]], pf=[[9.0, 1.0], [1.0, 1.1]]),
    # 4 m swap in 4 s: the rest-to-rest min-energy move peaks at |a| = 1.43 m/s^2; limit 1.3 binds at both ends
    "n2_swap_acc": dict(N=2, T=4.0, h=0.2, R=0.8, space=[0.0, 0.0, 6.0, 6.0], acc=1.3,
                        p0=[[1.0, 1.0], [5.0, 1.2]], pf=[[5.0, 1.0], [1.0, 1.2]]),
    # perpendicular crossing at (3, 3): one agent hurries, the other waits -- the manoeuvre needs |jerk| = 1.13 m/s^3
    # along the travel axes (the straight moves 0.75); limit 0.9 binds
    "n2_cross_jerk": dict(N=2, T=4.0, h=0.2, R=0.8, space=[0.0, 0.0, 6.0, 6.0], jerk=0.9,
                          p0=[[1.0, 3.0], [3.0, 1.2]], pf=[[5.0, 3.0], [3.0, 5.2]]),
}
CASES += [(k, None, None, None, None, None, None) for k in EXTRA]
OUT = os.path.join(ROOT, "tests", "golden")


def apply_limits(s, limits):
    """Box limits are attributes of the reference's SCP object (scp.py:67-74)."""
    for key, names in (("vel", ("vel_min", "vel_max")), ("acc", ("acc_min", "acc_max")), ("jerk", ("jerk_min", "jerk_max"))):
        if limits and key in limits:
            setattr(s, names[0], -float(limits[key]))
            setattr(s, names[1], float(limits[key]))


def run_reference(ref, N, T, h, R, space, p0, pf, overrides, limits=None):
    import osqp

    osqp.OVERRIDES.clear()
    osqp.OVERRIDES.update(overrides)
    osqp.STATS.clear()
    with ref_loader.quiet() as buf:
        s = ref.solvers.scp.SCP(n_vehicles=N, time_horizon=T, time_step=h, min_distance=R, space_dims=space)
        apply_limits(s, limits)
        s.set_initial_states(p0)
        s.set_final_states(pf)
        tr = s.generate_trajectories(max_iterations=15)
    rels = [float(x) for x in buf.getvalue().splitlines() if x and (x[0].isdigit() or x.startswith("nan"))]
    stats = list(osqp.STATS)
    osqp.OVERRIDES.clear()
    return tr, rels, stats


def main(only=None):
    ref = ref_loader.load_reference()
    gen = ref.scenarios.position_generator.generate_positions
    for name, N, T, h, R, space, seed in CASES:
        limits, explicit = None, None
        if name in EXTRA:
            ex = EXTRA[name]
            if "base" in ex:
                base = [c for c in CASES if c[0] == ex["base"]][0]
                _, N, T, h, R, space, seed = base
            else:
                N, T, h, R, space, seed = ex["N"], ex["T"], ex["h"], ex["R"], ex["space"], -1
                explicit = (np.array(ex["p0"], float), np.array(ex["pf"], float))
            limits = {k: ex[k] for k in ("vel", "acc", "jerk") if k in ex}
        # the certificate, not the ADMM tolerance, makes the result exact; larger cases start the refinement earlier
        eps = 1e-5 if (N < 15 and name not in EXTRA) else 1e-3   # binding box rows: the shim's ADMM needs > 2e5 iterations for 1e-5
        truth = dict(eps_abs=eps, eps_rel=eps, max_iter=200000, certify=True)
        if only and name not in only:
            continue
        path = os.path.join(OUT, f"{name}.npz")
        if os.path.exists(path):
            continue
        t0 = time.time()
        random.seed(max(seed, 0))
        np.random.seed(max(seed, 0))
        p0, pf = explicit if explicit is not None else gen(N, R)
        tr, rels, stats = run_reference(ref, N, T, h, R, space, p0, pf, truth, limits)
        certs = [s["cert"] for s in stats]
        if not all(s["polish"] == 1 and s["cert"] <= 1e-9 for s in stats):
            print(f"{name}: REJECTED (a subproblem did not certify: {certs})", flush=True)
            continue
        osqp = scp_oracle._osqp()
        osqp.OVERRIDES.update(truth)
        o = scp_oracle.ScpOracle(N, T, h, R, space)
        apply_limits(o, limits)
        o.set_initial_states(p0)
        o.set_final_states(pf)
        tro = o.generate_trajectories(15)
        osqp.OVERRIDES.clear()
        for k in ("positions", "velocities", "accelerations"):
            assert np.abs(tr[k] - tro[k]).max() <= 1e-9, (name, k)
        loose, rels_loose, _ = run_reference(ref, N, T, h, R, space, p0, pf, {}, limits)
        lim = dict(vel=2.0, acc=15.0, jerk=20.0)
        lim.update(limits or {})
        np.savez_compressed(
            path, N=N, T=T, h=h, R=R, space=np.array(space, float), seed=seed, p0=p0, pf=pf,
            positions=tr["positions"], velocities=tr["velocities"], accelerations=tr["accelerations"],
            rel_steps=np.array(rels), iterations=len(rels), certs=np.array(certs),
            a_initial=o.record["a_initial"], iterates=np.array(o.record["iterates"]),
            objective=float((tr["accelerations"] ** 2).sum()), min_separation=scp_oracle.min_separation(tr["positions"]),
            loose_positions=loose["positions"], loose_accelerations=loose["accelerations"],
            loose_rel_steps=np.array(rels_loose),
            vel_limit=lim["vel"], acc_limit=lim["acc"], jerk_limit=lim["jerk"],
        )
        print(f"{name}: {len(rels)} SCP iterations, max cert {max(certs):.1e}, {time.time()-t0:.0f}s", flush=True)


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
