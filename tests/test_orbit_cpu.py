"""Orbital forward simulation: the C restatement in oracle/bellman_oracle.c of
position-control/private/{kepler_U,f_and_g,fDot_and_gDot,sv_from_coe,stumpC,stumpS,rkf45}.m and
Solver_position.m:189-361, pinned against the worked examples of the textbook those files come
from (H. D. Curtis, Orbital Mechanics for Engineering Students, Examples 3.6, 3.7, 4.7 — the
reference stores no output of this path) and against an independent high-accuracy integration."""
import numpy as np


def test_kepler_U_textbook_example_3_6(oracle_lib):
    # ro = 10000 km, vro = 3.0752 km/s, dt = 3600 s, a = -5.0878e-5 1/km  ->  x = 128.511 km^0.5
    x, n = oracle_lib.kepler_U(398600.0, 3600.0, 10000.0, 3.0752, -5.0878e-5)
    assert abs(x - 128.511) < 5e-4 and 1 <= n <= 10


def test_update_RV_target_textbook_example_3_7(oracle_lib):
    R, V = oracle_lib.update_RV_target(398600.0, [7000.0, -12124.0, 0.0], [2.6679, 4.6210, 0.0], 3600.0)
    np.testing.assert_allclose(R, [-3297.77, 7413.40, 0.0], atol=6e-3)
    np.testing.assert_allclose(V, [-8.2976, -0.964045, 0.0], atol=6e-5)


def test_sv_from_coe_textbook_example_4_7(oracle_lib, bellman):
    d = np.pi / 180
    coe = [80000.0, 1.4, 40 * d, 30 * d, 60 * d, 30 * d]
    r, v = oracle_lib.sv_from_coe(coe, 398600.0)
    np.testing.assert_allclose(r, [-4039.9, 4814.56, 3628.62], atol=6e-3)
    np.testing.assert_allclose(v, [-10.386, -4.77192, 1.74388], atol=6e-5)
    # the facade's own sv_from_coe (numpy) agrees with the C restatement
    rf, vf = bellman.Solver_position.sv_from_coe(coe, 398600.0)
    np.testing.assert_allclose(rf, r, rtol=1e-14)
    np.testing.assert_allclose(vf, v, rtol=1e-14)
    R0, V0 = oracle_lib.target_R0V0()
    Rf, Vf = bellman.Solver_position().get_target_R0V0()
    np.testing.assert_allclose(Rf, R0, rtol=1e-15)
    np.testing.assert_allclose(Vf, V0, rtol=1e-15)
    np.testing.assert_allclose(R0, [6678.0, 0.0, 0.0], atol=1e-9)      # perigee of a 300 km x e = 0.1 orbit


def _rkf45_literal(ode, tspan, y0, tol=1.e-8, truncate_before_eval=False):
    """position-control/private/rkf45.m:49-118, line by line (f is the 6-column matrix of the .m file).
    truncate_before_eval=True is NOT the reference: it clips h to tf - t before the six evaluations."""
    import math
    a = np.array([0, 1 / 4, 3 / 8, 12 / 13, 1, 1 / 2])
    b = np.array([[0, 0, 0, 0, 0], [1 / 4, 0, 0, 0, 0], [3 / 32, 9 / 32, 0, 0, 0],
                  [1932 / 2197, -7200 / 2197, 7296 / 2197, 0, 0], [439 / 216, -8, 3680 / 513, -845 / 4104, 0],
                  [-8 / 27, 2, -3544 / 2565, 1859 / 4104, -11 / 40]])
    c4 = np.array([25 / 216, 0, 1408 / 2565, 2197 / 4104, -1 / 5, 0])
    c5 = np.array([16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55])
    t0, tf = tspan
    t, y = t0, np.array(y0, dtype=float)
    h = (tf - t0) / 100
    f = np.zeros((len(y), 6))
    while t < tf:
        hmin = 16 * np.spacing(abs(t)) if t != 0 else 16 * 5e-324
        ti, yi = t, y.copy()
        if truncate_before_eval:
            h = min(h, tf - t)
        for i in range(6):
            t_inner = ti + a[i] * h
            y_inner = yi.copy()
            for j in range(i):
                y_inner = y_inner + h * b[i, j] * f[:, j]
            f[:, i] = ode(t_inner, y_inner)
        te = np.array([sum((h * f[k, i]) * (c4[i] - c5[i]) for i in range(6)) for k in range(len(y))])
        te_max = max(abs(te))
        ymax = max(abs(y))
        te_allowed = tol * max(ymax, 1.0)
        delta = math.pow(te_allowed / (te_max + np.finfo(float).eps), 1 / 5)
        if te_max <= te_allowed:
            h = min(h, tf - t)
            t = t + h
            y = yi + np.array([sum((h * f[k, i]) * c5[i] for i in range(6)) for k in range(len(y))])
        h = min(delta * h, 4 * h)
        if h < hmin:
            break
    return y


def test_rkf45_stage_loop_restatement(bellman, oracle_lib):
    """The C stage loop (policy lookup + one rkf45 call per stage) against a second, line-by-line Python
    restatement of rkf45.m driving the same dynamics; and a sanity check of the dynamics themselves.
    rkf45.m clips the LAST step of every call to tf - t AFTER its six slope evaluations (rkf45.m:104),
    so the reference's own result is only first-order accurate in the stage length: restated as is
    (agreement with the literal to 1e-12), and shown to be the only source of the 2e-4 gap to a
    high-accuracy integration (clipping before the evaluations closes it to 1e-7)."""
    from scipy.integrate import solve_ivp
    sp = bellman.Solver_position()
    d = bellman.tables.stack_problems(sp._axis_descs())
    idx = np.stack([np.full(d.S, c, dtype=np.int32) for c in (2, 1, 0)])     # a = (+0.26, 0, -0.26)
    R0, V0 = oracle_lib.target_R0V0()
    mu, n_steps = 398600.0, 200
    y0 = np.array([[-1.0, 0.0, 0.0, 0.0, 0.0, 0.0], [0.2, -0.1, 0.05, 0.01, 0.02, -0.03]])
    X, Cc, W = oracle_lib.rollout_orbit(d, idx, sp.U_vector, y0, n_steps, sp.h, R0, V0, mu=mu)
    assert np.all(W == 0) and np.all(Cc == [2, 1, 0])
    acc = sp.U_vector[[2, 1, 0]]

    def rhs(t, y):
        R, V = oracle_lib.update_RV_target(mu, R0, V0, t)
        nR = np.linalg.norm(R)
        H = np.linalg.norm(np.cross(R, V))
        RdV = R @ V
        ax = (2 * mu / nR ** 3 + H ** 2 / nR ** 4) * y[0] - 2 * RdV / nR ** 4 * H * y[1] + 2 * H / nR ** 2 * y[4] + acc[0]
        ay = -(mu / nR ** 3 - H ** 2 / nR ** 4) * y[1] + 2 * RdV / nR ** 4 * H * y[0] - 2 * H / nR ** 2 * y[3] + acc[1]
        az = -mu / nR ** 3 * y[2] + acc[2]
        return np.array([y[3], y[4], y[5], ax, ay, az])

    for b in range(2):
        y_lit, y_fix = y0[b].copy(), y0[b].copy()
        for k in range(n_steps):
            y_lit = _rkf45_literal(rhs, (k * sp.h, (k + 1) * sp.h), y_lit)
            y_fix = _rkf45_literal(rhs, (k * sp.h, (k + 1) * sp.h), y_fix, truncate_before_eval=True)
            if k in (0, 9, n_steps - 1):
                np.testing.assert_allclose(X[b, k + 1], y_lit, rtol=0, atol=1e-12)
        sol = solve_ivp(rhs, [0.0, n_steps * sp.h], y0[b], method="DOP853", rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(y_fix, sol.y[:, -1], rtol=0, atol=1e-7)
        assert 1e-5 < np.max(np.abs(y_lit - sol.y[:, -1])) < 1e-3          # the reference's own first-order error
