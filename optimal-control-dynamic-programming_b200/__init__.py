"""bellman_b200 — B200-native backward Bellman value-iteration sweep behind the reference's
classdef surface (Dynamic_Solver / Solver_position / Solver_attitude / Solver_pos_att).

Import as ``import bellman_b200`` (shim at the repository root; this directory's name is not a
valid Python identifier).
"""
from . import tables  # noqa: F401
from ._lib import (BellmanError, Sweep, SweepGroup, KERNEL_AUTO, KERNEL_DIRECT, KERNEL_WINDOW,  # noqa: F401
                   KERNEL_SPLITC, KERNEL_TILE, EXPORTS, LIB_PATH, load, plan_slabs, query_locate, query_stencil, get_unique_id, rollout_pos_att, dense6_run, rollout_attitude6)
from .solvers import Dynamic_Solver, Solver_position, Solver_attitude, Solver_pos_att  # noqa: F401
