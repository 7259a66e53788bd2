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
