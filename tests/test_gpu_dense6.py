"""The coupled 6-D attitude sweep on the GPU (bellman_dense6_run, k_stage_dense6) against the C
restatement of the dense stage operator: J and argmin bit-equal.  Reference: Solver_attitude.run,
attitude-control/Solver_attitude.m:521-601 (never executed there; semantics = the code as written)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solver(bellman, nw, nq, N):
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_q = nw, nq
    sa.T_final = N * sa.h
    return sa


@pytest.mark.parametrize("nw,nq,stages", [(10, 6, 12), (7, 9, 5), (16, 4, 3)])
def test_dense6_matches_oracle(bellman, oracle_lib, nw, nq, stages):
    sa = _solver(bellman, nw, nq, 40)
    T = sa.dense6_tables()
    Jg, Ig, ms = bellman.dense6_run(T, stages)
    Jo, Io = oracle_lib.dense6_run(T, stages)
    assert np.array_equal(Ig, Io), "argmin differs on %d states" % np.sum(Ig != Io)
    assert np.array_equal(Jg, Jo)
    assert ms > 0


def test_dense6_rough_terminal_cost_and_other_control_counts(bellman, oracle_lib):
    sa = _solver(bellman, 8, 5, 10)
    rng = np.random.default_rng(1)
    for U in ([-0.11, 0.0, 0.11], [-0.2, -0.05, 0.05, 0.2], [0.0, 0.3]):      # nu = 3 (unrolled kernel), 4 and 2 (generic)
        sa.U_vector = np.array(U)
        T = sa.dense6_tables()
        JN = rng.normal(size=T.S) * 5
        Jg, Ig, _ = bellman.dense6_run(T, 3, J_N=JN)
        Jo, Io = oracle_lib.dense6_run(T, 3, J_N=JN)
        assert np.array_equal(Ig, Io) and np.array_equal(Jg, Jo)
        assert Ig.max() < len(U) ** 3 and len(np.unique(Ig)) > len(U)


def test_dense6_facade_run_and_errors(bellman, oracle_lib):
    sa = _solver(bellman, 9, 5, 8)
    sa.run()
    T = sa.dense6_tables()
    Jo, Io = oracle_lib.dense6_run(T, 7)
    shape = tuple(T.n)
    assert np.array_equal(sa.F_Values, Jo.reshape(shape, order="F"))
    c = ((sa.U_idx6[0] - 1) * 3 + (sa.U_idx6[1] - 1)) * 3 + (sa.U_idx6[2] - 1)
    assert np.array_equal(c.ravel(order="F"), Io)
    assert set(np.unique(sa.U1_Opt)) <= set(sa.U_vector) and sa.U3_Opt.shape == shape
    big = bellman.Solver_attitude()                                   # the reference's default mesh: 1000^3 x 10^3
    with pytest.raises(ValueError):
        big.run()
    T.nu = 9
    with pytest.raises(bellman.BellmanError):
        bellman.dense6_run(T, 1)


def test_attitude6_rollout_matches_oracle(bellman, oracle_lib):
    """run() then get_optimal_path() (Solver_attitude.m:1487-1530) for 200 initial states: identical torque
    sequences, states within 1e-10 of the C restatement (atan2 / asin are CUDA's on the GPU)."""
    sa = _solver(bellman, 12, 6, 60)
    sa.run(n_stages=40)
    T, idx = sa._dense6
    rng = np.random.default_rng(3)
    x0 = np.zeros((200, 7))
    x0[:, 0:3] = rng.uniform(-0.7, 0.7, (200, 3))
    x0[:, 3:6] = rng.uniform(-0.15, 0.15, (200, 3))
    x0[:, 6] = np.sqrt(1 - np.sum(x0[:, 3:6] ** 2, axis=1))
    x0[0] = sa.defaultX0_ode45
    Xg, Ug = sa.get_optimal_path(x0, n_steps=300)
    Xo, Uo = oracle_lib.rollout_attitude6(T, idx, (sa.J1, sa.J2, sa.J3), sa.h, 300, x0)
    same = np.all(Ug == Uo, axis=(1, 2))
    assert same.mean() >= 0.95, "torque sequences differ on %d of %d trajectories" % ((~same).sum(), len(same))
    np.testing.assert_allclose(Xg[same], Xo[same], rtol=0, atol=1e-10)
    assert len(np.unique(Uo)) == 3
    X1, U1 = sa.get_optimal_path(n_steps=10)                           # default X0
    assert X1.shape == (1, 11, 7) and U1.shape == (1, 10, 3) and np.array_equal(X1[0], Xg[0, :11])
    with pytest.raises(RuntimeError):
        _solver(bellman, 6, 4, 10).get_optimal_path()
