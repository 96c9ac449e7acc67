"""pose_graph_initialization_b200 — B200-native (sm_100a) hypothesis-verification path of
danini/pose-graph-initialization behind the C-ABI of include/pgi.h.

(The directory name uses underscores because `pose-graph-initialization_b200` is not importable.)
"""
from . import engine, scene  # noqa: F401
from .engine import Engine, PgiError, VERDICT_DTYPE  # noqa: F401
