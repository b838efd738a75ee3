"""ctypes binding of the C ABI in include/scp_b200.h (libscp_b200.so).

This is the only door from Python into the solver.  There is no CPU fallback:
if the CUDA library is missing, or no CUDA device is present, every compute
entry point raises.  Struct layouts mirror include/scp_b200.h field by field.
"""

from __future__ import annotations

import ctypes as C
import os

MAX_SCP_ITER = 32
ABI_VERSION = 4

_HERE = os.path.dirname(os.path.abspath(__file__))
# in-tree build (csrc/Makefile -> ../lib); an installed package points SCP_B200_LIB at its copy
LIB_PATH = os.environ.get("SCP_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libscp_b200.so")


class Problem(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32),
        ("n_steps", C.c_int32),
        ("time_step", C.c_double),
        ("min_distance", C.c_double),
        ("space", C.c_double * 4),
        ("vel_limit", C.c_double),
        ("acc_limit", C.c_double),
        ("jerk_limit", C.c_double),
        ("scp_tolerance", C.c_double),
        ("feas_margin", C.c_double),
        ("max_scp_iter", C.c_int32),
        ("max_admm_iter", C.c_int32),
        ("check_every", C.c_int32),
        ("adapt_every", C.c_int32),
        ("polish", C.c_int32),
        ("eps_abs", C.c_double),
        ("eps_rel", C.c_double),
        ("rho0", C.c_double),
        ("sigma", C.c_double),
        ("w_jerk", C.c_double),
        ("w_acc", C.c_double),
        ("w_vel", C.c_double),
        ("w_pos", C.c_double),
        ("w_col", C.c_double),
        ("cand_margin", C.c_double),
        ("verify_tol", C.c_double),
        ("polish_first_eps", C.c_double),
        ("polish_first", C.c_int32),
        ("relax_pct", C.c_int32),
        ("stall_window", C.c_int32),
        ("warm_duals", C.c_int32),
        ("polish_rounds", C.c_int32),
        ("team_mode", C.c_int32),
        ("lazy_rows", C.c_int32),
        ("momentum_pct", C.c_int32),
        ("max_admm_iter_qp0", C.c_int32),
        ("cap_halving", C.c_int32),
        ("polish_max_failed", C.c_int32),
    ]


class Record(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("scp_iterations", C.c_int32),
        ("converged", C.c_int32),
        ("initial_feasible", C.c_int32),
        ("admm_iterations", C.c_int32),
        ("qp_unsolved", C.c_int32),
        ("rebuilds", C.c_int32),
        ("max_copies", C.c_int32),
        ("first_violation", C.c_int32 * 3),
        ("polish_ok", C.c_int32),
        ("qp_infeasible", C.c_int32),
        ("polish_attempts", C.c_int32),
        ("first_violation_dist", C.c_double),
        ("min_separation", C.c_double),
        ("objective", C.c_double),
        ("pri_res", C.c_double),
        ("dua_res", C.c_double),
        ("cand_row_iters", C.c_double),
        ("cycles_total", C.c_int64),
        ("cycles_admm", C.c_int64),
        ("cycles_polish", C.c_int64),
        ("polish_rounds", C.c_int32),
        ("reserved2", C.c_int32),
        ("cycles_pbuild", C.c_int64),
        ("cycles_psolve", C.c_int64),
        ("cycles_peval", C.c_int64),
        ("cycles_papply", C.c_int64),
        ("device_ns", C.c_int64),
        ("rel_step", C.c_double * MAX_SCP_ITER),
    ]


class Check(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "min_separation", "min_separation_step", "min_separation_continuous", "min_separation_continuous_time",
        "box_violation", "dynamics_violation", "terminal_violation", "dynamics_residual")]


STATUS_OK = 0
STATUS_INITIAL_QP_FAILED = 1
STATUS_START_TOO_CLOSE = 2

# name -> (restype, argtypes); every symbol include/scp_b200.h declares
_P = C.POINTER
_SIGNATURES = {
    "scp_b200_abi_version": (C.c_int, []),
    "scp_b200_last_error": (C.c_char_p, []),
    "scp_b200_sizeof_problem": (C.c_size_t, []),
    "scp_b200_sizeof_record": (C.c_size_t, []),
    "scp_b200_measure_fp64_peak": (C.c_int, [_P(C.c_double)]),
    "scp_b200_default_problem": (None, [_P(Problem), C.c_int, C.c_double, C.c_double, C.c_double]),
    "scp_b200_tables_bytes": (C.c_size_t, [_P(Problem)]),
    "scp_b200_build_tables": (C.c_int, [_P(Problem), C.c_void_p, C.c_void_p]),
    "scp_b200_workspace_bytes": (C.c_size_t, [_P(Problem), C.c_int]),
    "scp_b200_default_slots": (C.c_int, [_P(Problem)]),
    "scp_b200_solve_batch": (
        C.c_int,
        [_P(Problem), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "scp_b200_solve_batch_host": (
        C.c_int,
        [_P(Problem), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int],
    ),
    "scp_b200_reconstruct": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
         C.c_void_p],
    ),
    "scp_b200_linearize_range": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "scp_b200_stream_default_problem": (None, [_P(Problem), C.c_int, C.c_double, C.c_double, C.c_double]),
    "scp_b200_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "scp_b200_stream_create": (
        C.c_int, [_P(Problem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, _P(C.c_void_p)]),
    "scp_b200_stream_destroy": (None, [C.c_void_p]),
    "scp_b200_stream_ipc_handles": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scp_b200_stream_ipc_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scp_b200_stream_solve": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, _P(C.c_float), _P(C.c_int64)],
    ),
    "scp_b200_stream_solve_host": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         _P(C.c_float), _P(C.c_int64)],
    ),
    "scp_b200_generate_scenarios": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "scp_b200_check_batch": (
        C.c_int,
        [_P(Problem), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "scp_b200_linearize": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
}

_lib = None


class ScpB200Error(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load libscp_b200.so (built by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ScpB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.scp_b200_abi_version() != ABI_VERSION:
        raise ScpB200Error("ABI version mismatch between _capi.py and libscp_b200.so")
    if lib.scp_b200_sizeof_problem() != C.sizeof(Problem) or lib.scp_b200_sizeof_record() != C.sizeof(Record):
        raise ScpB200Error("struct layout mismatch between _capi.py and include/scp_b200.h")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().scp_b200_last_error()
        raise ScpB200Error(f"scp_b200 error {rc}: {msg.decode() if msg else '?'}")


def default_problem(n_agents, time_horizon, time_step, min_distance, space_dims=None, lib=None, stream=False) -> Problem:
    p = Problem()
    fn = (lib or load()).scp_b200_stream_default_problem if stream else (lib or load()).scp_b200_default_problem
    fn(C.byref(p), int(n_agents), float(time_horizon), float(time_step),
                                             float(min_distance))
    if space_dims is not None:
        for i in range(4):
            p.space[i] = float(space_dims[i])
    return p


def record_to_dict(r: Record) -> dict:
    n = max(0, min(int(r.scp_iterations), MAX_SCP_ITER))
    return dict(
        status=int(r.status), scp_iterations=int(r.scp_iterations), converged=bool(r.converged),
        initial_feasible=bool(r.initial_feasible), admm_iterations=int(r.admm_iterations),
        qp_unsolved=int(r.qp_unsolved), rebuilds=int(r.rebuilds), max_copies=int(r.max_copies),
        first_violation=tuple(int(v) for v in r.first_violation), polish_ok=int(r.polish_ok), qp_infeasible=int(r.qp_infeasible), polish_attempts=int(r.polish_attempts),
        cycles_total=int(r.cycles_total), cycles_admm=int(r.cycles_admm), cycles_polish=int(r.cycles_polish), polish_rounds=int(r.polish_rounds), cycles_pbuild=int(r.cycles_pbuild), cycles_psolve=int(r.cycles_psolve),
        cycles_peval=int(r.cycles_peval), cycles_papply=int(r.cycles_papply),
        first_violation_dist=float(r.first_violation_dist), min_separation=float(r.min_separation),
        objective=float(r.objective), pri_res=float(r.pri_res), dua_res=float(r.dua_res), cand_row_iters=float(r.cand_row_iters),
        reserved2=int(r.reserved2), device_ns=int(r.device_ns),
        candidate_overflow=bool(int(r.reserved2) & 4),   # streaming solver: collision rows dropped for lack of slots
        rel_steps=[float(r.rel_step[i]) for i in range(n)],
    )
