"""The coupled 6-D attitude sweep, Solver_attitude.run (attitude-control/Solver_attitude.m:521-601,
SURVEY 8f row 4), on the CPU: the facade's tables and the C restatement of the dense stage operator
against the line-by-line numpy restatement that builds the reference's full 9-D arrays
(oracle/matlab_literal.SolverAttitude6Literal).  Parity unpinned: the reference never ran this path."""
import numpy as np
import pytest


def _literal_and_tables(bellman, nw, nq, N):
    from oracle.matlab_literal import SolverAttitude6Literal
    L = SolverAttitude6Literal(n_mesh_w=nw, n_mesh_q=nq, N_stage=N)
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_q = nw, nq
    sa.T_final = N * sa.h
    T = sa.dense6_tables()
    return L, sa, T


def test_dense6_tables_match_the_literal_arrays(bellman):
    """grids and the J_N = 0 first stage: min over the 27 combinations of J_current_state_fix is attained
    at U = 0 everywhere, so stage N-1 equals the state cost — which pins gs and r against the 9-D array."""
    L, sa, T = _literal_and_tables(bellman, 5, 4, 3)
    for g, want in zip(T.grid, (L.sr_1, L.sr_2, L.sr_3, L.s_yaw, L.s_pitch, L.s_roll)):
        assert np.array_equal(g, want)
    F, U1, U2, U3 = L.run(n_stages=1)
    assert np.array_equal(T.gs, F.ravel(order="F")) and np.all(U1 == 2) and np.all(U2 == 2) and np.all(U3 == 2)
    assert T.w_next[0].shape == (3, 125) and T.a_next[2].shape == (T.S,) and T.S == 5 ** 3 * 4 ** 3
    # next states stay near the node they start from (h = 5 ms): w within h*(|k w w| + U/J), angles within h*|w|
    w1 = np.tile(T.grid[0], 25)
    assert np.max(np.abs(T.w_next[0] - w1[None])) < 0.005 * (0.2 * 0.88 ** 2 + 0.11 / 0.0245) * 1.01
    yaw = np.broadcast_to(T.grid[3].reshape(1, 1, 1, -1, 1, 1), T.n).ravel(order="F")
    assert np.max(np.abs(T.a_next[0] - yaw)) < 0.005 * 0.88 * 2.5


@pytest.mark.parametrize("nw,nq,N", [(5, 4, 8), (4, 3, 12)])
def test_dense6_oracle_vs_literal(bellman, oracle_lib, nw, nq, N):
    L, sa, T = _literal_and_tables(bellman, nw, nq, N)
    F, U1, U2, U3 = L.run()
    J, idx = oracle_lib.dense6_run(T, N - 1)
    c = (((U1 - 1) * 3 + (U2 - 1)) * 3 + (U3 - 1)).ravel(order="F")
    assert np.array_equal(idx, c)                       # nested min over U3, U2, U1 = first flat minimiser
    np.testing.assert_allclose(J, F.ravel(order="F"), rtol=1e-13, atol=0)
    assert len(np.unique(idx)) > 1


def test_dense6_oracle_rough_terminal_cost(bellman, oracle_lib):
    """A rough J_N makes every corner and weight matter (64-corner interpolation, extrapolation at the
    edges of all six dimensions)."""
    L, sa, T = _literal_and_tables(bellman, 4, 3, 3)
    JN = np.random.default_rng(0).normal(size=T.S) * 5
    F, U1, U2, U3 = L.run(n_stages=2, J_N=JN)
    J, idx = oracle_lib.dense6_run(T, 2, J_N=JN)
    c = (((U1 - 1) * 3 + (U2 - 1)) * 3 + (U3 - 1)).ravel(order="F")
    Fl = F.ravel(order="F")
    np.testing.assert_allclose(J, Fl, rtol=0, atol=1e-12)
    bad = idx != c
    assert bad.mean() < 0.01                            # (1-t)*a + t*b vs fma(t, b-a, a): only near-ties may flip


def test_attitude6_rollout_oracle_vs_literal(bellman, oracle_lib):
    """Solver_attitude.get_optimal_path (:1487-1530): the C restatement against the line-by-line numpy
    restatement (quat2angle, three 6-D 'nearest' interpolants, 'taylor' step) under a swept 6-D policy."""
    from oracle.matlab_literal import attitude6_get_optimal_path_literal
    L, sa, T = _literal_and_tables(bellman, 8, 5, 40)
    J, idx = oracle_lib.dense6_run(T, 30)
    rng = np.random.default_rng(0)
    x0 = np.zeros((4, 7))
    x0[:, 0:3] = rng.uniform(-0.6, 0.6, (4, 3))
    x0[:, 3:6] = rng.uniform(-0.15, 0.15, (4, 3))
    x0[:, 6] = np.sqrt(1 - np.sum(x0[:, 3:6] ** 2, axis=1))
    x0[0] = sa.defaultX0_ode45
    X, U = oracle_lib.rollout_attitude6(T, idx, (sa.J1, sa.J2, sa.J3), sa.h, 120, x0)
    shape = tuple(T.n)
    vals = [T.U_vector[k].reshape(shape, order="F") for k in (idx // 9, (idx // 3) % 3, idx % 3)]
    seen = set()
    for b in range(4):
        Xl, Ul = attitude6_get_optimal_path_literal(T.grid, vals, (sa.J1, sa.J2, sa.J3), sa.h, x0[b], 120)
        assert np.array_equal(U[b], Ul.T)
        np.testing.assert_allclose(X[b], Xl.T, rtol=0, atol=1e-13)
        seen |= set(np.unique(U[b]))
    assert seen == {-0.11, 0.0, 0.11}
    np.testing.assert_allclose(np.linalg.norm(X[:, -1, 3:7], axis=1), 1.0, atol=1e-14)    # renormalised every step


@pytest.mark.parametrize("U", [[0.0, 0.3], [-0.2, -0.05, 0.05, 0.2]])
def test_dense6_oracle_vs_literal_other_control_counts(bellman, oracle_lib, U):
    """nu = 2 and 4 torque levels (the reference fixes nu = 3): the flat argmin c = (u1*nu + u2)*nu + u3 of
    the C restatement against the three nested min calls of the literal."""
    from oracle.matlab_literal import SolverAttitude6Literal
    L = SolverAttitude6Literal(n_mesh_w=4, n_mesh_q=3, N_stage=6, U_vector=np.array(U))
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_q, sa.U_vector = 4, 3, np.array(U)
    sa.T_final = 6 * sa.h
    T = sa.dense6_tables()
    JN = np.random.default_rng(2).normal(size=T.S)
    F, U1, U2, U3 = L.run(n_stages=3, J_N=JN)
    J, idx = oracle_lib.dense6_run(T, 3, J_N=JN)
    nu = len(U)
    c = (((U1 - 1) * nu + (U2 - 1)) * nu + (U3 - 1)).ravel(order="F")
    np.testing.assert_allclose(J, F.ravel(order="F"), rtol=0, atol=1e-12)
    assert (idx != c).mean() < 0.01 and idx.max() < nu ** 3
