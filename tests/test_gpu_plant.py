"""Full-plant forward simulations on the GPU (bellman_rollout_pos_att / bellman_rollout_attitude, one
thread per initial state, ode45 restated) against the C restatement of
Solver_pos_att.get_optimal_path (pos-att/Solver_pos_att.m:452-500, :692-757) and
Solver_attitude.get_optimal_path_simplified_testode45 (attitude-control/Solver_attitude.m:1669-1705).
pow / asin / cos / sin come from CUDA's math library on the GPU and from the C library in the oracle,
so the bar is a tolerance: identical thruster / torque sequences, states within 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctl_from_idx(d, idx):
    fv = [d.meta[k] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")]
    return {"GridVectors": [d.grid[k][0] for k in range(4)], "U_Optimal_id": idx.reshape(tuple(d.n), order="F") + 1,
            "f0_allcomb": fv[0], "f1_allcomb": fv[1], "f6_allcomb": fv[2], "f7_allcomb": fv[3], "stop_stage": 1}


def _random_states(rng, batch):
    y0 = np.zeros((batch, 13))
    y0[:, 0:3] = rng.uniform(-0.18, 0.18, size=(batch, 3))
    y0[:, 3:6] = rng.uniform(-0.08, 0.08, size=(batch, 3))
    y0[:, 6:9] = rng.uniform(-0.04, 0.04, size=(batch, 3))
    y0[:, 9] = np.sqrt(1 - np.sum(y0[:, 6:9] ** 2, axis=1))
    y0[:, 10:13] = rng.uniform(-0.03, 0.03, size=(batch, 3))
    return y0


def test_pos_att_plant_rollout_matches_oracle(bellman, oracle_lib):
    """Controllers swept on the GPU at a reduced mesh (150 stages: the thrusters fire), installed with
    set_controller, then 200 stages of the 13-state plant for 64 initial states."""
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 12, 10, 8, 7
    sp.check_period = 0                                                 # no early stop: exactly 150 stages
    descs, idxs, fvals = [], [], []
    for ci, ch in enumerate("xyz"):
        ctl = sp.calculate_one_channel_U_Opt(ci, n_stages=150)
        d = sp.channel_desc(ci)
        idx = (np.asarray(ctl["U_Optimal_id"]) - 1).astype(np.int32).ravel(order="F")
        assert np.array_equal(idx, oracle_lib.sweep(d, n_stages=150)["idx_last"][0])
        sp.set_controller(ctl, ch)
        descs.append(d); idxs.append(idx)
        fvals.append(np.stack([d.meta[k] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")]))
    rng = np.random.default_rng(5)
    y0 = np.vstack([sp.default_X0(), _random_states(rng, 63)])
    n_steps = 200
    Xg, Fg, FMg = sp.get_optimal_path(y0, n_steps=n_steps)
    R0, V0 = sp.get_target_R0V0()
    Xo, Fo, FMo, Wo = oracle_lib.rollout_pos_att(descs, idxs, fvals, y0, n_steps, sp.h, R0, V0, sp.InertiaM, sp.Mass, sp.T_dist)
    assert np.all(sp.ode45_warnings == 0) and np.all(Wo == 0)
    assert len(np.unique(Fo.reshape(-1, 12), axis=0)) > 8
    same = np.all(Fg == Fo, axis=(1, 2))
    assert same.mean() >= 0.95, "thruster sequences differ on %d of %d trajectories" % ((~same).sum(), len(same))
    np.testing.assert_allclose(Xg[same], Xo[same], rtol=0, atol=1e-9)
    np.testing.assert_allclose(FMg[same], FMo[same], rtol=1e-12, atol=1e-15)
    # stride_out and the facade default (the reference's X0)
    X1, F1, FM1 = sp.get_optimal_path(n_steps=40)
    X4, F4, FM4 = sp.get_optimal_path(n_steps=40, stride_out=4)
    assert X1.shape == (1, 41, 13) and F1.shape == (1, 40, 12) and X4.shape == (1, 11, 13) and FM4.shape == (1, 10, 6)
    np.testing.assert_array_equal(X4[0], X1[0, ::4])
    np.testing.assert_array_equal(F4[0], F1[0, ::4])
    np.testing.assert_array_equal(X1[0], Xg[0, :41])


def test_pos_att_plant_rollout_random_policy_and_failure_channel(bellman, oracle_lib):
    """A random policy switches thrusters every few stages (the harshest lookup test); the x channel uses
    the 6-combination failure-mode controller, so the channels have different C."""
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 9, 8, 7, 6
    rng = np.random.default_rng(11)
    descs, idxs, fvals = [], [], []
    for ci, ch in enumerate("xyz"):
        d = sp.channel_desc(ci, failure=(ci == 0))
        idx = rng.integers(0, d.C, size=d.S).astype(np.int32)
        sp.set_controller(_ctl_from_idx(d, idx), ch)
        descs.append(d); idxs.append(idx)
        fvals.append(np.stack([d.meta[k] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")]))
    assert descs[0].C == 6 and descs[1].C == 9
    y0 = _random_states(rng, 40)
    Xg, Fg, FMg = sp.get_optimal_path(y0, n_steps=120)
    R0, V0 = sp.get_target_R0V0()
    Xo, Fo, FMo, Wo = oracle_lib.rollout_pos_att(descs, idxs, fvals, y0, 120, sp.h, R0, V0, sp.InertiaM, sp.Mass, sp.T_dist)
    same = np.all(Fg == Fo, axis=(1, 2))
    assert same.mean() >= 0.9
    np.testing.assert_allclose(Xg[same], Xo[same], rtol=0, atol=1e-9)
    assert np.all(Fg[..., 0] == 0)                                      # thruster 0 failed (:236-240)


def test_attitude_plant_rollout_matches_oracle(bellman, oracle_lib):
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 120, 60
    sa.simplified_run(n_stages=200)
    d = sa._desc
    idx = sa._sweep.get_idx()
    assert np.array_equal(idx, oracle_lib.sweep(d, n_stages=200)["idx_last"])
    rng = np.random.default_rng(9)
    y0 = np.zeros((48, 7))
    y0[:, 0:3] = rng.uniform(-0.5, 0.5, size=(48, 3))
    y0[:, 3:6] = rng.uniform(-0.12, 0.12, size=(48, 3))
    y0[:, 6] = np.sqrt(1 - np.sum(y0[:, 3:6] ** 2, axis=1))
    y0[0] = sa.defaultX0_ode45
    n_steps = 400
    Xg, Ug = sa.get_optimal_path_simplified_testode45(y0, n_steps=n_steps)
    Xo, Co, Wo = oracle_lib.rollout_attitude(d, idx, sa.U_vector, y0, n_steps, sa.h, sa.InertiaM)
    assert np.all(sa.ode45_warnings == 0) and np.all(Wo == 0)
    Uo = np.asarray(sa.U_vector)[Co]
    assert len(np.unique(Co)) == 3
    same = np.all(Ug == Uo, axis=(1, 2))
    assert same.mean() >= 0.95, "torque sequences differ on %d of %d trajectories" % ((~same).sum(), len(same))
    np.testing.assert_allclose(Xg[same], Xo[same], rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(Xg[:, -1, 3:7], axis=1), 1.0, atol=1e-6)


def test_plant_rollout_argument_errors(bellman):
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 6, 6, 5, 5
    rng = np.random.default_rng(2)
    for ci, ch in enumerate("xyz"):
        d = sp.channel_desc(ci)
        sp.set_controller(_ctl_from_idx(d, rng.integers(0, d.C, size=d.S).astype(np.int32)), ch)
    with pytest.raises(bellman.BellmanError):
        sp.get_optimal_path(n_steps=10, stride_out=3)                   # n_steps not a multiple of stride_out
    bad = sp.InertiaM.copy()
    bad[0, 1] += 1e-3
    sp.InertiaM = bad
    with pytest.raises(bellman.BellmanError):
        sp.get_optimal_path(n_steps=4)                                  # inertia must be symmetric
