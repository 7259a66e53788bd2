"""Orbital forward simulation on the GPU (bellman_rollout_orbit, one thread per initial state) against
the C restatement of Solver_position.get_optimal_path (Solver_position.m:189-224 + private/*.m).
cos / sin / cosh / sinh / pow come from CUDA's math library on the GPU and from the C library in the
oracle, so the bar here is a tolerance: identical control sequences, states within 1e-9 (the
integrator's own tolerance is 1e-8 per stage)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_orbit_rollout_matches_oracle(bellman, oracle_lib):
    sp = bellman.Solver_position()
    n_sweep, n_steps = 400, 600
    sp.simplified_run(n_stages=n_sweep)
    d = sp._desc
    ora = oracle_lib.sweep(d, n_stages=n_sweep)
    idx = sp._sweep.get_idx()
    assert np.array_equal(idx, ora["idx_last"])
    rng = np.random.default_rng(3)
    y0 = np.concatenate([rng.uniform(-0.45, 0.45, size=(95, 3)), rng.uniform(-0.3, 0.3, size=(95, 3))], axis=1)
    y0 = np.vstack([[-1.0, 0, 0, 0, 0, 0], y0])          # the reference's own start (outside the grid: clamped policy)
    Xg, Ug = sp.get_optimal_path(y0, n_steps=n_steps)
    R0, V0 = sp.get_target_R0V0()
    Xo, Co, Wo = oracle_lib.rollout_orbit(d, idx, sp.U_vector, y0, n_steps, sp.h, R0, V0, mu=sp.mu)
    assert np.all(sp.rkf45_warnings == 0) and np.all(Wo == 0)
    Uo = np.asarray(sp.U_vector)[Co]
    same = np.all(Ug == Uo, axis=(1, 2))
    # a trajectory whose state passes within rounding of a cell midpoint may pick the other node
    assert same.mean() >= 0.98, "control sequences differ on %d of %d trajectories" % ((~same).sum(), len(same))
    np.testing.assert_allclose(Xg[same], Xo[same], rtol=0, atol=1e-9)


def test_orbit_rollout_stride_and_facade_defaults(bellman, oracle_lib):
    sp = bellman.Solver_position()
    sp.simplified_run(n_stages=50)
    X1, U1 = sp.get_optimal_path(n_steps=120)                    # default y0 = [-1 0 0 0 0 0] (Solver_position.m:195-197)
    X4, U4 = sp.get_optimal_path(n_steps=120, stride_out=4)
    assert X1.shape == (1, 121, 6) and U1.shape == (1, 120, 3) and X4.shape == (1, 31, 6) and U4.shape == (1, 30, 3)
    np.testing.assert_array_equal(X4[0], X1[0, ::4])
    np.testing.assert_array_equal(U4[0], U1[0, ::4])
