"""Drop-in for the SCP path of jankammeth/BA-path-planning, B200-native.

Same import surface as the reference package (src/path_planning/__init__.py:1-5):
``SCP``, ``generate_positions``, ``make_boxplot``; the solve itself runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/scp_b200.h.
"""

from .scenarios.position_generator import generate_positions
from .solvers.scp import SCP
from .viz.plot_runtime_boxplot import make_boxplot

__all__ = ["SCP", "generate_positions", "make_boxplot"]
