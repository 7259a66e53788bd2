"""'Next' rows on the CPU: the oracle's nearest-policy lookup and simplified-plant rollout against the
literal numpy restatement of attitude-control/test/test_simplified.m (10x10 grids, J = 2 / 2.5 / 3,
h = 0.01, U = [-0.01 0 0.01], time-varying U*_Opt(:,:,k))."""
import numpy as np

from oracle import matlab_literal as ml


def simplified_test_desc(bellman, axis, N):
    t = bellman.tables
    J = (2.0, 2.5, 3.0)[axis]
    ang = ((-5.0, 5.0), (-4.0, 4.0), (-5.5, 5.5))[axis]
    d = t.attitude_axis_desc(t.deg2rad(-0.7), t.deg2rad(0.7), 10, ang[0], ang[1], 10, [-0.01, 0.0, 0.01], J,
                             6.0, 6.0, 0.1, 0.01, N)
    return d, J


def test_oracle_rollout_matches_literal_test_simplified(bellman, oracle_lib):
    N, h = 80, 0.01
    U_vector = np.array([-0.01, 0.0, 0.01])
    for axis in range(3):
        d, J = simplified_test_desc(bellman, axis, N)
        out = oracle_lib.sweep(d, keep_all=True)
        idx_all = out["idx_all"][:, 0, :]                                   # [N, S], row k-1 = stage k
        s_w, s_t = d.grid[0][0], d.grid[1][0]
        u_inc = (h * (((U_vector / J + 2 * (U_vector / J)) + 2 * (U_vector / J)) + U_vector / J)) / 6
        U_opt = U_vector[idx_all[:N - 1].reshape(N - 1, 10, 10).transpose(2, 1, 0)]   # [n_w, n_t, stage]
        for x0 in ([0.004, 0.03], [-0.011, -0.06], [0.02, 0.2]):             # last one starts off-grid
            Xl, Ul = ml.simplified_axis_rollout_literal(s_w, s_t, U_opt, 0, lambda u: u / J, h, x0, N - 1)
            Xo, Co = oracle_lib.rollout_axis(d, idx_all, u_inc, [x0], N - 1, h, 0, time_varying=True)
            assert np.array_equal(U_vector[Co[0]], Ul)
            assert np.array_equal(Xo[0], Xl)
        # fixed (stage-1) policy, as Solver_attitude keeps it (Solver_attitude.m:249-251)
        Xl, Ul = ml.simplified_axis_rollout_literal(s_w, s_t, U_opt[:, :, 0], 0, lambda u: u / J, h, [0.004, 0.03], 200)
        Xo, Co = oracle_lib.rollout_axis(d, idx_all[0], u_inc, [[0.004, 0.03]], 200, h, 0)
        assert np.array_equal(Xo[0], Xl) and np.array_equal(U_vector[Co[0]], Ul)


def test_oracle_policy_lookup_matches_nearest_interpolant(bellman, oracle_lib):
    rng = np.random.default_rng(3)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 8, 7, 6, 5
    d = sp.channel_desc(0)
    out = oracle_lib.sweep(d, n_stages=5)
    idx = out["idx_last"][0]
    grids = [d.grid[k][0] for k in range(4)]
    lo = np.array([g[0] for g in grids]); hi = np.array([g[-1] for g in grids])
    x = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(500, 4))
    x[:8] = [[g[i] for g in grids] for i in range(5)] + [lo, hi, 0.5 * (grids[0][:1].repeat(4) + lo)]
    F = ml.GriddedInterpolantNearest(grids, idx.reshape([len(g) for g in grids], order="F"))
    want = F(*[x[:, k] for k in range(4)])
    got = oracle_lib.policy_lookup(d, idx, x)
    assert np.array_equal(got, want)
