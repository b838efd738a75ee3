// scp_defaults.h -- default problem settings (shared by the CUDA library and the test emulation).
#ifndef SCP_DEFAULTS_H
#define SCP_DEFAULTS_H
#include "../../include/scp_b200.h"

static inline void scp_fill_default_problem(scp_b200_problem* p, int n_agents, double time_horizon,
                                            double time_step, double min_distance) {
  p->n_agents = n_agents;
  p->n_steps = (int)(time_horizon / time_step);   /* K = int(T/h), scp.py:43 (same double division) */
  p->time_step = time_step;
  p->min_distance = min_distance;
  p->space[0] = 0; p->space[1] = 0; p->space[2] = 20; p->space[3] = 20;  /* scp.py:47-49 */
  p->vel_limit = 2.0; p->acc_limit = 15.0; p->jerk_limit = 20.0;         /* scp.py:67-74 */
  p->scp_tolerance = 1.5e-2;                                            /* scp.py:52 */
  p->feas_margin = 0.01;                                                /* scp.py:610 */
  p->max_scp_iter = 15;                                                 /* scp.py:131 */
  p->max_admm_iter = 5000;
  p->check_every = 25;
  p->adapt_every = 100;
  p->polish = 1;
  p->eps_abs = 1e-5; p->eps_rel = 1e-5;
  p->rho0 = 1.0; p->sigma = 1e-6;
  p->w_jerk = 0.4; p->w_acc = 2.0; p->w_vel = 10.0; p->w_pos = 1.0; p->w_col = 4.0;
  p->cand_margin = 0.5;
  p->verify_tol = 1e-6;
  p->polish_first_eps = 5e-2;
  p->polish_first = 0;
  p->relax_pct = 0;
  p->stall_window = 500;
  p->warm_duals = 0;
  p->polish_rounds = 40;
  p->team_mode = 0;
  p->lazy_rows = 1;
  p->momentum_pct = 0;
  p->max_admm_iter_qp0 = 0;
  p->cap_halving = 1;
  p->polish_max_failed = 10;
}
#endif
