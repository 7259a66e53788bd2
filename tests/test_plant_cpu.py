"""Full-plant forward simulations that the reference integrates with ode45 (second
half): Solver_pos_att.get_optimal_path (pos-att/Solver_pos_att.m:452-500, :692-757) and
Solver_attitude.get_optimal_path_simplified_testode45 (attitude-control/Solver_attitude.m:1669-1705).
The C restatement (oracle/bellman_oracle.c) against the independent numpy restatement
(oracle/plant_literal.py) and against a high-accuracy integration of the same plant.  Parity unpinned:
the reference stores no output of these paths; ode45's step control is restated from its published
algorithm (see the module docstrings)."""
import numpy as np
import pytest


def test_ode45_restatement_known_behaviour(oracle_lib):
    """y' = -y on [0, 1]: MaxStep = 0.1*(tf - t0) binds, so ode45 takes exactly 10 steps, no rejection,
    61 evaluations; a decaying mode faster than the step limit triggers the rejection branch.  C and
    numpy statements take identical step sequences."""
    from oracle import plant_literal as pl
    y, n, nf, w = oracle_lib.ode45_linear([-1.0], 0.0, 1.0, [1.0])
    assert (n, nf, w) == (10, 0, 0) and abs(y[0] - np.exp(-1.0)) < 1e-7
    for lam, tf in (([-1.0, -50.0, 3.0], 2.0), ([-400.0], 1.0), ([2.0, -0.5], 5.0)):
        st = {}
        y0 = np.ones(len(lam))
        yl = pl.ode45_last(lambda t, y: np.array(lam) * y, (0.0, tf), y0, stats=st)
        yc, n, nf, w = oracle_lib.ode45_linear(lam, 0.0, tf, y0)
        assert (n, nf, w) == (st["nsteps"], st["nfailed"], 0) and nf >= (1 if min(lam) < -20 else 0)
        np.testing.assert_allclose(yc, yl, rtol=1e-9, atol=1e-300)   # summation order of f*hB differs (BLAS vs loop)
        slow = [i for i, l in enumerate(lam) if l > -10]
        np.testing.assert_allclose(yc[slow], np.exp(np.array(lam)[slow] * tf), rtol=2e-3)    # RelTol 1e-3


def _pos_att_controllers(bellman, oracle_lib, stages=150, mesh=(8, 7, 6, 5), random_policy=None):
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = mesh
    descs, idxs, fvals, ctls = [], [], [], {}
    for ci, ch in enumerate("xyz"):
        d = sp.channel_desc(ci)
        if random_policy is None:
            idx = oracle_lib.sweep(d, n_stages=stages)["idx_last"][0]
        else:
            idx = random_policy.integers(0, d.C, size=d.S).astype(np.int32)
        fv = np.stack([d.meta[k] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")])
        descs.append(d); idxs.append(idx); fvals.append(fv)
        ctls[ch] = {"GridVectors": [d.grid[k][0] for k in range(4)],
                    "U_Optimal_id": idx.reshape(tuple(d.n), order="F") + 1,
                    "f0_allcomb": fv[0], "f1_allcomb": fv[1], "f6_allcomb": fv[2], "f7_allcomb": fv[3]}
    return sp, descs, idxs, fvals, ctls


def test_pos_att_plant_oracle_vs_literal(bellman, oracle_lib):
    from oracle import plant_literal as pl
    sp, descs, idxs, fvals, ctls = _pos_att_controllers(bellman, oracle_lib)
    R0, V0 = oracle_lib.target_R0V0()
    n_steps = 40
    y0 = np.stack([pl.default_X0_pos_att(),                                # the reference's X0 (:458-468)
                   np.array([0.05, -0.08, 0.02, 0.01, -0.02, 0.005, 0.02, -0.03, 0.01, 0.0, 0.01, -0.02, 0.015])])
    y0[1, 9] = np.sqrt(1 - np.sum(y0[1, 6:9] ** 2))
    np.testing.assert_allclose(y0[0, 6:10], [0, np.sin(np.deg2rad(1.5)), 0, np.cos(np.deg2rad(1.5))], atol=1e-16)
    X, F, FM, W = oracle_lib.rollout_pos_att(descs, idxs, fvals, y0, n_steps, sp.h, R0, V0, sp.InertiaM, sp.Mass, sp.T_dist)
    assert np.all(W == 0)
    lit = pl.PosAttPlantLiteral(ctls, sp.InertiaM, sp.Mass, sp.T_dist, sp.h, R0, V0)
    for b in range(2):
        st = {}
        Xl, Fl, FMl = lit.get_optimal_path(y0[b], n_steps, stats=st)
        assert np.array_equal(F[b], Fl)                                     # identical thruster sequences
        np.testing.assert_allclose(FM[b], FMl, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(X[b], Xl, rtol=0, atol=1e-12)
        assert set(st["nsteps"]) == {10} and set(st["nfailed"]) == {0}      # MaxStep binds: 10 steps per stage
    assert len(np.unique(F.reshape(-1, 12), axis=0)) > 3                    # the policy actually switches
    # the thruster levels are the channel's on/off values and the moments follow :806-809
    assert set(np.unique(np.abs(F))) <= {0.0, 0.13}
    np.testing.assert_allclose(FM[..., 4], (F[..., 0] - F[..., 1] + F[..., 6] - F[..., 7]) * sp.T_dist, atol=1e-18)


def test_pos_att_plant_against_high_accuracy_integration(bellman, oracle_lib):
    """Piecewise-constant thrusters from the restated run, plant integrated by DOP853 at 1e-12: the ode45
    restatement stays within its own tolerance class (RelTol 1e-3 per step; far tighter here because the
    step limit, not the error estimate, sets the step)."""
    from scipy.integrate import solve_ivp
    from oracle import plant_literal as pl
    sp, descs, idxs, fvals, ctls = _pos_att_controllers(bellman, oracle_lib, random_policy=np.random.default_rng(3))
    R0, V0 = oracle_lib.target_R0V0()
    n_steps = 30
    y0 = pl.default_X0_pos_att()[None]
    X, F, FM, W = oracle_lib.rollout_pos_att(descs, idxs, fvals, y0, n_steps, sp.h, R0, V0, sp.InertiaM, sp.Mass, sp.T_dist)
    lit = pl.PosAttPlantLiteral(ctls, sp.InertiaM, sp.Mass, sp.T_dist, sp.h, R0, V0)
    y = y0[0].copy()
    for k in range(n_steps):
        U_M, acc = FM[0, k, 3:], FM[0, k, :3]
        sol = solve_ivp(lambda t, s: lit.rates(t, s, U_M, acc), [k * sp.h, (k + 1) * sp.h], y, method="DOP853", rtol=1e-12, atol=1e-15)
        y = sol.y[:, -1]
    np.testing.assert_allclose(X[0, -1], y, rtol=0, atol=1e-9)
    assert abs(np.linalg.norm(X[0, -1, 6:10]) - 1) < 1e-9                  # quaternion norm is an invariant of :745-748


def test_attitude_plant_oracle_vs_literal(bellman, oracle_lib):
    from oracle import plant_literal as pl
    from oracle.matlab_literal import GriddedInterpolantNearest
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 40, 30
    d = bellman.tables.stack_problems(sa._axis_descs())
    out = oracle_lib.sweep(d, n_stages=30)
    idx = out["idx_last"]
    n_steps = 60
    y0 = np.array([[0.1, -0.2, 0.15, 0.05, -0.04, 0.08, 0.0], [-0.3, 0.25, 0.0, -0.1, 0.02, 0.12, 0.0]])
    y0[:, 6] = np.sqrt(1 - np.sum(y0[:, 3:6] ** 2, axis=1))
    X, Cc, W = oracle_lib.rollout_attitude(d, idx, sa.U_vector, y0, n_steps, sa.h, sa.InertiaM)
    assert np.all(W == 0) and len(np.unique(Cc)) > 1
    FU = [GriddedInterpolantNearest([d.grid[0][a], d.grid[1][a]], sa.U_vector[idx[a].reshape(tuple(d.n), order="F")]) for a in range(3)]
    lit = pl.AttitudePlantLiteral(FU, sa.InertiaM, sa.h)
    for b in range(2):
        Xl, Ul = lit.run(y0[b], n_steps)
        assert np.array_equal(sa.U_vector[Cc[b]], Ul)
        np.testing.assert_allclose(X[b], Xl, rtol=0, atol=1e-12)


def test_pos_att_facade_constants_match_the_restatements(bellman, oracle_lib):
    """Host-side pieces of Solver_pos_att.get_optimal_path that need no GPU: the target orbit
    (get_target_R0V0, :759-777) and the default initial state (:458-468)."""
    from oracle import plant_literal as pl
    sp = bellman.Solver_pos_att()
    R0, V0 = sp.get_target_R0V0()
    Ro, Vo = oracle_lib.target_R0V0()
    np.testing.assert_allclose(R0, Ro, rtol=1e-15)
    np.testing.assert_allclose(V0, Vo, rtol=1e-15)
    np.testing.assert_allclose(sp.default_X0(), pl.default_X0_pos_att(), rtol=0, atol=1e-17)
    assert abs(np.linalg.norm(sp.default_X0()[6:10]) - 1) < 1e-15
    with pytest.raises(KeyError):
        sp.get_optimal_path(n_steps=4)                     # no controller installed yet (set_controller)
