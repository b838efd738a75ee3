/*
 * scp_b200.h -- C ABI of the B200-native SCP trajectory solver.
 *
 * Drop-in boundary for ONE path of jankammeth/BA-path-planning:
 *   SCP.generate_trajectories  (src/path_planning/solvers/scp.py:131-180)
 * and the private helpers it calls.  The reference has no FFI of its own (it is
 * pure Python calling the third-party `osqp` C core at scp.py:360-362 and
 * scp.py:441-445); each entry point below names the reference code it replaces.
 * The reference-side binding is a ctypes stub, shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - every function returns 0 on success, non-zero on error; the message of the
 *     last error of the calling thread is returned by scp_b200_last_error();
 *   - "d_" pointers are device pointers owned by the caller (PyTorch in the
 *     Python host), "h_" pointers are host pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - all real data is IEEE double, the reference's dtype (numpy float64);
 *   - trajectories use the reference layout (N, K, 2): agent-major, then time,
 *     then x/y (scp.py:141,168,171-175); a batch adds a leading B.
 */
#ifndef SCP_B200_H
#define SCP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCP_B200_ABI_VERSION 4
#define SCP_B200_MAX_SCP_ITER 32

/* Problem definition shared by every scenario of a batch.
 * Mirrors SCP.__init__ (scp.py:32-97) + the solver settings that the reference
 * leaves at OSQP defaults. */
typedef struct scp_b200_problem {
  int32_t n_agents;        /* N           scp.py:40 */
  int32_t n_steps;         /* K=int(T/h)  scp.py:43 */
  double time_step;        /* h           scp.py:42 */
  double min_distance;     /* R           scp.py:44 */
  double space[4];         /* xmin,ymin,xmax,ymax  scp.py:47-49,63-64 */
  double vel_limit;        /* 2.0   scp.py:67-68 */
  double acc_limit;        /* 15.0  scp.py:70-71 */
  double jerk_limit;       /* 20.0  scp.py:73-74 */
  double scp_tolerance;    /* 1.5e-2 scp.py:52 */
  double feas_margin;      /* 0.01: gate distance < R - margin, scp.py:610 */
  int32_t max_scp_iter;    /* 15    scp.py:131 */
  /* QP solver (replaces osqp, scp.py:326-367 and 441-451) */
  int32_t max_admm_iter;   /* per ADMM run */
  int32_t check_every;     /* residual test period */
  int32_t adapt_every;     /* rho adaptation period */
  int32_t polish;          /* 1: active-set polish with KKT certificate */
  double eps_abs, eps_rel; /* residual tolerances */
  double rho0, sigma;      /* step size, x-proximal weight */
  double w_jerk, w_acc, w_vel, w_pos, w_col; /* row-class weights (unit rho) */
  double cand_margin;      /* collision rows kept when prev. distance < R + margin */
  double verify_tol;       /* dropped rows must hold to this tolerance */
  double polish_first_eps; /* residual gate: polish is tried once pri,dua <= gate*(1+norm) and the active set has settled */
  int32_t polish_first;    /* >0: polish attempt with this many rounds before the first ADMM iteration of a subproblem */
  int32_t relax_pct;       /* streaming solver: ADMM over-relaxation alpha in percent (0 or 100: none; OSQP default 160) */
  int32_t stall_window;    /* give up on a subproblem whose primal residual stalls over this many iterations (0: off) */
  int32_t warm_duals;      /* 1: keep multipliers across SCP iterations (reference/OSQP restarts from y = 0) */
  int32_t polish_rounds;   /* add/drop rounds per polish attempt */
  int32_t team_mode;       /* 0 auto, 1 one CTA per scenario, 2 whole cooperative grid per scenario */
  int32_t lazy_rows;       /* streaming solver: 1 = the box-row classes (jerk, acc, vel, pos) start outside the ADMM and a
                              class joins, per scenario, when a converged iterate violates it (verify-and-enlarge, as
                              for the collision rows); 0 = all box rows carried from the start */
  int32_t momentum_pct;    /* streaming solver: heavy-ball extrapolation of the collision state in percent (0: off) */
  /* --- ABI 3 --- */
  int32_t max_admm_iter_qp0; /* iteration cap of the initial QP (OSQP default max_iter 4000, scp.py:360); 0: max_admm_iter */
  int32_t cap_halving;     /* 1: every subproblem of a scenario that ran into the iteration cap halves the cap of its next
                              ones (floor 500); 0: every subproblem gets the full cap, like scp.py:442 */
  int32_t polish_max_failed; /* failed polish attempts per subproblem before it ends on the ADMM residual test only */
} scp_b200_problem;

/* Per-scenario result record (device or host array of B records). */
typedef struct scp_b200_record {
  int32_t status;          /* SCP_B200_STATUS_* */
  int32_t scp_iterations;  /* trips of the loop scp.py:152-166 */
  int32_t converged;       /* rel step <= tol reached */
  int32_t initial_feasible;/* gate scp.py:144 passed (loop skipped) */
  int32_t admm_iterations; /* total over all subproblems */
  int32_t qp_unsolved;     /* subproblems that hit max_admm_iter (reference: warning only, scp.py:446-447) */
  int32_t rebuilds;        /* candidate-set enlargements */
  int32_t max_copies;      /* largest per-(agent,step) candidate count */
  int32_t first_violation[3]; /* k,i,j of the gate's first violation, -1 if none */
  int32_t polish_ok;       /* subproblems that ended with a KKT certificate */
  int32_t qp_infeasible;   /* subproblems stopped by the primal infeasibility certificate (subset of qp_unsolved) */
  int32_t polish_attempts;
  double first_violation_dist;
  double min_separation;   /* of the returned positions, k in [0,K) */
  double objective;        /* sum ||a||^2 of the returned accelerations */
  double pri_res, dua_res; /* of the last subproblem */
  double cand_row_iters;   /* sum over ADMM iterations of collision rows actually carried */
  int64_t cycles_total, cycles_admm, cycles_polish; /* SM clock cycles spent on this scenario */
  int32_t polish_rounds;   /* add/drop rounds over all polish attempts */
  int32_t reserved2;
  int64_t cycles_pbuild, cycles_psolve, cycles_peval, cycles_papply; /* polish breakdown */
  int64_t device_ns;       /* ABI 3: nanoseconds of GPU time spent on THIS scenario (one-CTA solver: sum over its quanta,
                              %globaltimer; streaming solver: batch start to the moment the scenario finished) */
  double rel_step[SCP_B200_MAX_SCP_ITER]; /* scp.py:157-160, one per trip */
} scp_b200_record;

enum {
  SCP_B200_STATUS_OK = 0,
  SCP_B200_STATUS_INITIAL_QP_FAILED = 1, /* reference raises RuntimeError, scp.py:363-365 */
  SCP_B200_STATUS_START_TOO_CLOSE = 2,   /* some ||p0_i-p0_j|| < R: the k=0 rows (scp.py:487-496) are infeasible */
  SCP_B200_STATUS_NOT_RUN = -1
};

int scp_b200_abi_version(void);
/* sizeof(scp_b200_problem) / sizeof(scp_b200_record): lets a foreign binding check its struct layout. */
size_t scp_b200_sizeof_problem(void);
size_t scp_b200_sizeof_record(void);
const char* scp_b200_last_error(void);

/* Measurement aid (bench.py): sustained fp64 FMA throughput of the current device in TFLOP/s -- the roofline
 * denominator of the solver kernels, whose arithmetic is fp64 on the FMA pipe (the reference computes in numpy float64). */
int scp_b200_measure_fp64_peak(double* tflops_out);

/* Fill `prob` with the reference's defaults (scp.py:32-74, OSQP-free settings). */
void scp_b200_default_problem(scp_b200_problem* prob, int n_agents, double time_horizon,
                              double time_step, double min_distance);

/* Constant operator tables for one (K, h, weights): replaces the constant part of
 * _precompute_constraint_matrices (scp.py:182-232).  Host computes, then uploads. */
size_t scp_b200_tables_bytes(const scp_b200_problem* prob);
int scp_b200_build_tables(const scp_b200_problem* prob, void* d_tables, void* stream);

/* Scratch for `slots` concurrently resident scenarios (one CTA each). */
size_t scp_b200_workspace_bytes(const scp_b200_problem* prob, int slots);
int scp_b200_default_slots(const scp_b200_problem* prob);

/* The whole of SCP.generate_trajectories (scp.py:131-180) for B independent
 * scenarios, device buffers, one launch, no host round trip inside.
 *   d_p0,d_v0,d_pf,d_vf : (B,N,2)   set_initial_states / set_final_states, scp.py:99-129
 *   d_acc,d_pos,d_vel   : (B,N,K,2) the result dict, scp.py:171-175
 *   d_records           : B records */
int scp_b200_solve_batch(const scp_b200_problem* prob, int n_scenarios,
                         const double* d_p0, const double* d_v0, const double* d_pf,
                         const double* d_vf, const void* d_tables, void* d_workspace,
                         size_t workspace_bytes, int slots, double* d_acc, double* d_pos,
                         double* d_vel, scp_b200_record* d_records, void* stream);

/* Same, host buffers: allocates device memory, copies in, solves, copies out.
 * This is the call a non-PyTorch host (or the reference's own SCP class through
 * ctypes) makes; it is what bench.py's "e2e" figure times. */
int scp_b200_solve_batch_host(const scp_b200_problem* prob, int n_scenarios, const double* h_p0,
                              const double* h_v0, const double* h_pf, const double* h_vf,
                              double* h_acc, double* h_pos, double* h_vel,
                              scp_b200_record* h_records, int device);

/* _compute_positions_velocities / _accelerations_to_positions_velocities
 * (scp.py:371-397, 559-595): (B,N,K,2) accelerations -> positions, velocities. */
int scp_b200_reconstruct(const double* d_acc, const double* d_p0, const double* d_v0, int n_scenarios,
                         int n_agents, int n_steps, double time_step, double* d_pos, double* d_vel,
                         void* stream);

/* Collision linearisation, _add_collision_constraints (scp.py:453-557), matrix
 * free: for every row (k, i<j) in the reference's order (k-major, then i<j
 * lexicographic) writes eta (2 doubles) and the bound of
 *     eta . (p_i[k] - p_j[k]) >= bound          
 * and reduces the minimum separation per scenario (the quantity tested by
 * _fast_check_avoidance_constraints, scp.py:597-615) plus the first violating
 * row in scan order.
 *   d_pos   : (B,N,K,2)          d_eta : (B,K,P,2)   d_bound : (B,K,P)   P = N(N-1)/2
 *   d_minsep: (B)                d_first: (B,3) int32 (k,i,j or -1)
 * d_eta / d_bound may be NULL (reduction only). */
int scp_b200_linearize(const double* d_pos, int n_scenarios, int n_agents, int n_steps,
                       double min_distance, double feas_margin, double* d_eta, double* d_bound,
                       double* d_minsep, int32_t* d_first, void* stream);

/* Same for the pair indices [pair_begin, pair_end) only (pair index p = i(2N-i-1)/2 + j-i-1, i<j): the share of one
 * rank when the agents of a large scenario are sharded over GPUs.  d_eta : (B,K,Pr,2), d_bound : (B,K,Pr) with
 * Pr = pair_end - pair_begin; d_minsep / d_first reduce over this range only (combine across ranks with MIN). */
int scp_b200_linearize_range(const double* d_pos, int n_scenarios, int n_agents, int n_steps,
                             double min_distance, double feas_margin, int64_t pair_begin, int64_t pair_end,
                             double* d_eta, double* d_bound, double* d_minsep, int32_t* d_first, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The steps either side of the solve.
 *
 * scp_b200_generate_scenarios: batched rejection-sampling scenario generator on the device -- the construction and
 * acceptance rules of reference scenarios/position_generator.py:44-75 (layout 0: starts on the four corner circles,
 * goals on the central diamond / circles of the 20 x 20 m box, pairwise spacing >= min_distance, max_attempts draws
 * per set, default 1000) or of the bounded-travel generator used beyond 50 agents (layout 1: arena side sqrt(16 N),
 * spacing 1.25 R, travel 0.5..1 x 0.4 v_max T; default 200 N attempts).  Scenario b draws from a counter-based
 * stream keyed by (seed, first_scenario + b): reproducible, but NOT the host generator's Mersenne-twister draws.
 *   d_p0, d_pf : (B,N,2) out;  d_status : (B) int32 out, 1 = placed, 0 = attempts exhausted (the reference raises
 *   ValueError, position_generator.py:58-59, 72-73). */
int scp_b200_generate_scenarios(int n_scenarios, int n_agents, int layout, double min_distance, double time_horizon,
                                double vel_limit, uint64_t seed, int first_scenario, int max_attempts, double* d_p0,
                                double* d_pf, int32_t* d_status, void* stream);

/* Post-solve analysis of B trajectories (result dict scp.py:171-175): the minimum separation at the samples (the
 * quantity _fast_check_avoidance_constraints tests, scp.py:597-615, and print_distance_analysis reports,
 * position_generator.py:173-205), the minimum separation in CONTINUOUS time (piecewise-constant acceleration between
 * samples, scp.py:371-397: closed-form minimum of the quartic |d(t)|^2 on every interval), and the dynamics residual
 * of SURVEY.md 8(c): box rows on states 1..K-1 (scp.py:188-257), state recursion between consecutive samples,
 * terminal equalities on state K. */
typedef struct scp_b200_check {
  double min_separation;                 /* min over k in [0,K), i<j of ||p_i[k] - p_j[k]|| */
  double min_separation_step;            /* the k where it is attained */
  double min_separation_continuous;      /* min over t in [0,(K-1)h], i<j */
  double min_separation_continuous_time; /* the t where it is attained (seconds) */
  double box_violation;                  /* max violation of the jerk / acc / vel / pos boxes, >= 0 */
  double dynamics_violation;             /* max |state[k+1] - step(state[k], a[k])| and |state[0] - initial state| */
  double terminal_violation;             /* max |state[K] - final state| */
  double dynamics_residual;              /* max of the three: pass iff <= 1e-3 */
} scp_b200_check;
int scp_b200_check_batch(const scp_b200_problem* prob, int n_scenarios, const double* d_acc, const double* d_pos,
                         const double* d_vel, const double* d_p0, const double* d_v0, const double* d_pf,
                         const double* d_vf, scp_b200_check* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming solver: the same loop (scp.py:131-180) as a fixed sequence of HBM/L2-streaming kernels
 * over all scenarios and agents, per-scenario control flow decided on the device.  Two uses:
 *   world == 1 : a batch of B scenarios that are too large for one CTA's shared memory
 *                (BASELINE.json configs 3 and 5: 50-200 agents);
 *   world  > 1 : ONE scenario (B = 1) whose agents are sharded over the GPUs of a node
 *                (config 4): rank g owns agents [g*ceil(N/world), ...), every ADMM iteration ends
 *                with an NCCL all-gather of the positions (scp.py:463 is where the reference
 *                recomputes them), residual scalars are all-gathered at check iterations.
 * Subproblems end on the ADMM residual test (eps_abs/eps_rel of the problem), there is no polish.
 * n_steps <= 128.  The result records keep their meaning; cycles_* and polish_* stay 0 and
 * reserved2 bit 2 says that the per-(step,agent) candidate capacity was exceeded. */
typedef struct scp_b200_stream scp_b200_stream;

/* scp_b200_default_problem with the ADMM settings that suit the streaming solver (fixed rho, over-relaxation 1.6,
 * eps 1e-4, iteration cap 20000, lazy box rows, rho0 = clamp((50/K)^2, 0.1, 1), stall window 1000). */
void scp_b200_stream_default_problem(scp_b200_problem* prob, int n_agents, double time_horizon,
                                     double time_step, double min_distance);

/* 128-byte NCCL unique id (rank 0 creates it, the host side broadcasts it to the other ranks). */
int scp_b200_nccl_unique_id(void* id128);

/* max_candidates: collision rows carried per (step, agent), <= 48 (0: default 16).
 * nccl_id128 may be NULL when world == 1.  The CUDA device current at this call owns the solver. */
int scp_b200_stream_create(const scp_b200_problem* prob, int n_scenarios, int max_candidates, int rank,
                           int world, const void* nccl_id128, scp_b200_stream** out);
void scp_b200_stream_destroy(scp_b200_stream* solver);

/* Optional, world > 1: replace the per-iteration ncclAllGather of the positions by peer-memory stores over NVLink.
 * Every rank exports 192 bytes (3 CUDA IPC handles), the host all-gathers them into world x 192 bytes and every rank
 * connects; on failure (no peer access) the solver keeps using NCCL. */
int scp_b200_stream_ipc_handles(scp_b200_stream* solver, void* out192);
int scp_b200_stream_ipc_connect(scp_b200_stream* solver, const void* all_handles);

/* Device buffers as in scp_b200_solve_batch; with world > 1 every rank passes the same full-size
 * inputs and receives the full outputs.  Blocks until finished; *device_ms (optional) is the time
 * between the first and the last operation on `stream`, *macro_steps the check periods run. */
int scp_b200_stream_solve(scp_b200_stream* solver, const double* d_p0, const double* d_v0,
                          const double* d_pf, const double* d_vf, double* d_acc, double* d_pos,
                          double* d_vel, scp_b200_record* d_records, void* stream, float* device_ms,
                          int64_t* macro_steps);

/* Same with host buffers. */
int scp_b200_stream_solve_host(scp_b200_stream* solver, const double* h_p0, const double* h_v0,
                               const double* h_pf, const double* h_vf, double* h_acc, double* h_pos,
                               double* h_vel, scp_b200_record* h_records, float* device_ms,
                               int64_t* macro_steps);

#ifdef __cplusplus
}
#endif
#endif /* SCP_B200_H */
