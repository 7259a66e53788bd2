"""The reference's own test scripts replayed on the CPU oracle (no GPU):
test/test_u_star_M.m (Dynamic_Solver at its default size, rollouts from [2;1] under the steady-state
policy of six stages and under the time-varying policy — the author's comments are the known answers)
and test/test_griddedInterp.m (a bilinear surface is reproduced exactly by 'linear' interpolation and
by its extrapolation outside the grid)."""
import numpy as np

from _ref_scripts import SSU_STAGES, QUERY_POINTS, check_u_star_M_verdicts, interp_surface_desc


def test_script_test_u_star_M_on_oracle(bellman, oracle_lib):
    obj = bellman.Dynamic_Solver()                                    # 100 x 100 states, 1000 controls, N = 200
    d = obj._build()
    ora = oracle_lib.sweep(d, keep_all=True)
    x0 = np.array([[2.0, 1.0]])
    rollouts = {}
    for key, mode, ssu in [(("ssu", k), 1, k) for k in SSU_STAGES] + [(("Nssu", 1), 0, 1)]:
        X, U = oracle_lib.rollout(d, ora["idx_all"][:, 0, :], obj.A, obj.B, d.meta["U_mesh"], x0, mode=mode, ssu_stage=ssu)
        rollouts[key] = (X[0], U[0])
    cost = check_u_star_M_verdicts(obj, rollouts)
    # the closed-loop cost from a grid-adjacent start tracks the value function J_1 (interpolated) closely
    assert 70.0 < cost[("Nssu", 1)] < 85.0


def test_script_test_griddedInterp_on_oracle(bellman, oracle_lib):
    for pt in QUERY_POINTS:
        d, JN = interp_surface_desc(bellman, pt)
        out = oracle_lib.sweep(d, n_stages=1, J_N=JN)
        want = 2.0 * pt[0] * pt[1] + pt[1]
        np.testing.assert_allclose(out["J_last"][0], want, rtol=1e-12, atol=1e-12)     # every state sees F(point)
        assert np.all(out["idx_last"] == 0)
