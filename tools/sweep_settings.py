"""Profiling aid: times the batch solve for a few solver-setting combinations (config 2 scenarios)."""
import json, subprocess, sys
combos = [
    {},
    {"polish_first": 0, "warm_duals": 0},
    {"polish_first": 0, "warm_duals": 1},
    {"polish_first": 8, "warm_duals": 1},
    {"polish_first": 0, "warm_duals": 0, "check_every": 50},
    {"polish_first": 0, "warm_duals": 0, "polish_rounds": 20},
    {"polish_first": 0, "warm_duals": 0, "cand_margin": 1.0},
    {"polish_first": 0, "warm_duals": 0, "cand_margin": 0.25},
    {"polish_first": 0, "warm_duals": 0, "rho0": 3.0},
    {"polish_first": 0, "warm_duals": 0, "rho0": 0.3},
]
for c in combos:
    cmd = [sys.executable, "bench.py", "--scenarios", "592", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"]
    for k, v in c.items():
        cmd += ["--set", f"{k}={v}"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out); a = d["aux"]
    print(json.dumps(c), f"-> {d['value']:.0f} scen/s; admm/scen {a['admm_iterations_per_scenario']:.0f}; solved {a['scenarios_all_qps_solved']}; pass {a['scenarios_min_separation_pass']}; "
          f"rounds {a['rank0_polish_rounds']}; attempts {a['rank0_polish_attempts']}; ok {a['rank0_polish_ok']}; frac admm {a['rank0_cycles_frac_admm']:.2f} polish {a['rank0_cycles_frac_polish']:.2f}", flush=True)
