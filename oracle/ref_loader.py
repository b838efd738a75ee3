"""ORACLE / TEST INFRASTRUCTURE ONLY.

Imports the reference package *verbatim* from /root/reference/src under the
alias ``ref_path_planning`` (so it cannot collide with the drop-in package,
which keeps the reference's own name ``path_planning``).  The two modules the
reference needs and this image lacks -- ``osqp`` and ``matplotlib`` -- are
satisfied by the shims in oracle/shims (see their headers).

/root/reference exists only in the build container, never on the GPU box:
this loader is used by oracle/make_golden.py and by the container-only tests
that pin oracle/scp_oracle.py against the reference.  Nothing under tests/
marked ``gpu``, nor bench.py, nor smoke(), may call it.
"""

from __future__ import annotations

import contextlib
import importlib
import importlib.util
import io
import os
import sys

REFERENCE_SRC = os.environ.get("SCP_REFERENCE_SRC", "/root/reference/src")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
ALIAS = "ref_path_planning"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "path_planning", "solvers", "scp.py"))


def install_shims() -> None:
    """Put the osqp / matplotlib shims ahead of site-packages (idempotent)."""
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)


def load_reference():
    """Return the reference package object (alias ``ref_path_planning``)."""
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    if not reference_available():
        raise FileNotFoundError(f"reference sources not found under {REFERENCE_SRC}")
    install_shims()
    pkg_dir = os.path.join(REFERENCE_SRC, "path_planning")
    spec = importlib.util.spec_from_file_location(
        ALIAS, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    importlib.import_module(ALIAS + ".solvers.scp")
    importlib.import_module(ALIAS + ".scenarios.position_generator")
    importlib.import_module(ALIAS + ".cli.compute_trajectories_batch")
    return mod


@contextlib.contextmanager
def quiet():
    """The reference prints a banner and per-iteration lines; swallow them."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf
