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


def test_controller_file_round_trip_cpu(bellman, oracle_lib, tmp_path):
    """Solver_pos_att.save_controller / load_controller / set_controller / get_thruster_on_off_optimal
    (Solver_pos_att.m:291, :849-882, :404-449) with a controller produced by the oracle: no GPU involved."""
    rng = np.random.default_rng(7)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 8, 7, 6, 5
    ctls = {}
    for ci, ch in enumerate("xyz"):
        d = sp.channel_desc(ci)
        out = oracle_lib.sweep(d, n_stages=5)
        shape = tuple(d.n)
        ctl = {"GridVectors": [d.grid[k][0] for k in range(4)],
               "F_gI_Values": out["J_last"][0].reshape(shape, order="F"),
               "U_Optimal_id": out["idx_last"][0].reshape(shape, order="F") + 1,
               "f0_allcomb": d.meta["f0_allcomb"], "f1_allcomb": d.meta["f1_allcomb"],
               "f6_allcomb": d.meta["f6_allcomb"], "f7_allcomb": d.meta["f7_allcomb"], "stop_stage": d.N - 5}
        f = str(tmp_path / ("channel_%s_controller_1.mat" % ch))
        sp.save_controller(f, ctl)
        back = sp.load_controller(f)
        assert np.array_equal(back["U_Optimal_id"], ctl["U_Optimal_id"])
        assert np.array_equal(back["F_gI_Values"], ctl["F_gI_Values"]) and back["stop_stage"] == ctl["stop_stage"]
        sp.set_controller(f, ch)
        ctls[ch] = (d, ctl)
    # one state, all twelve thrusters, against the oracle's nearest lookup per channel
    for _ in range(50):
        x = rng.uniform(-0.25, 0.25, 3); v = rng.uniform(-0.12, 0.12, 3)
        t = rng.uniform(-0.1, 0.1, 3); w = rng.uniform(-0.04, 0.04, 3)
        f = sp.get_thruster_on_off_optimal(x, v, t, w)
        ang = {"x": 1, "y": 2, "z": 0}
        for ci, ch in enumerate("xyz"):
            d, ctl = ctls[ch]
            q = np.array([[x[ci], v[ci], t[ang[ch]], w[ang[ch]]]])
            c = oracle_lib.policy_lookup(d, (ctl["U_Optimal_id"] - 1).ravel(order="F"), q)[0]
            for name, thr in zip(("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"), sp._channel_thrusters[ch]):
                assert f[thr] == ctl[name][c]
    # frame change (:411-415): identity attitude, RSW axes aligned with ECI for R0 = e1, V0 = e2
    f1 = sp.get_thruster_on_off_optimal([0.01, -0.02, 0.03], [0.0, 0.01, 0.0], [0, 0, 0], [0, 0, 0],
                                        R0=[7000.0, 0, 0], V0=[0, 7.5, 0], q=[0, 0, 0, 1])
    f2 = sp.get_thruster_on_off_optimal([0.01, -0.02, 0.03], [0.0, 0.01, 0.0], [0, 0, 0], [0, 0, 0])
    assert np.array_equal(f1, f2)


def test_dynamic_solver_archive_cpu(bellman, golden, tmp_path):
    """Dynamic_Solver.save / load / compare_data (Dynamic_Solver.m:266-280) on the reference's stored run."""
    a = bellman.Dynamic_Solver()
    a.N, a.dx, a.du = golden["N"], golden["dx"], golden["du"]
    a.J_star, a.u_star = golden["J_star"], golden["u_star"]
    a.s_r = golden["X1_mesh"][:, 0]
    f = str(tmp_path / "obj_1_copy.mat")
    a.save(f)
    b = bellman.Dynamic_Solver.load(f)
    assert b.N == a.N and np.array_equal(b.s_r, a.s_r) and np.array_equal(b.u_star, a.u_star)
    assert bellman.Dynamic_Solver.compare_data(a, b)
    b.J_star = b.J_star.copy(); b.J_star[0, 0, 0] += 1e-9
    assert not bellman.Dynamic_Solver.compare_data(a, b)
