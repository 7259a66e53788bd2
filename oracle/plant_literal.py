"""TEST INFRASTRUCTURE (oracle) — array-at-a-time Python restatement of the two forward simulations of
the reference that integrate the full plant with ode45:

  * Solver_pos_att.get_optimal_path            pos-att/Solver_pos_att.m:452-500, ode_eq :692-757,
    get_thruster_on_off_optimal :404-449, to_Moments_Forces :805-823, ECI2body :825-829,
    RSW2ECI :831-847, update_RV_target :779-803 (+ private/kepler_U.m, f_and_g.m, fDot_and_gDot.m)
  * Solver_attitude.get_optimal_path_simplified_testode45
                                                attitude-control/Solver_attitude.m:1669-1705, :1795-1851

It is the second, independent statement the C oracle (oracle/bellman_oracle.c) is compared with on the
CPU: matrices are numpy arrays, the two backslashes are numpy.linalg.solve (LAPACK), the 'nearest'
interpolants are matlab_literal.GriddedInterpolantNearest.

ode45 is MathWorks' driver for the Dormand-Prince 5(4) pair; it is not in the reference repository.
`ode45_last` restates its published algorithm (Shampine & Reichelt, "The MATLAB ODE Suite", 1997; the
structure of the readable ode45.m / odearguments.m) for the DEFAULT options the two call sites use:
RelTol 1e-3, AbsTol 1e-6, MaxStep = 0.1*|tf - t0|, InitialStep from y'(t0), component-wise control.
Parity unpinned: the reference stores no output of these paths and MATLAB cannot run here.
Only tests/ may import this module."""
import math

import numpy as np

from .matlab_literal import GriddedInterpolantNearest

# Dormand-Prince tableau as ode45.m lays it out: column k of B is the argument of stage k + 2
A = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1, 1])
B = np.array([
    [1 / 5, 3 / 40, 44 / 45, 19372 / 6561, 9017 / 3168, 35 / 384],
    [0, 9 / 40, -56 / 15, -25360 / 2187, -355 / 33, 0],
    [0, 0, 32 / 9, 64448 / 6561, 46732 / 5247, 500 / 1113],
    [0, 0, 0, -212 / 729, 49 / 176, 125 / 192],
    [0, 0, 0, 0, -5103 / 18656, -2187 / 6784],
    [0, 0, 0, 0, 0, 11 / 84],
    [0, 0, 0, 0, 0, 0]])
E = np.array([71 / 57600, 0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40])


def _eps(t):
    return float(np.spacing(abs(t))) if t != 0 else 5e-324


def ode45_last(ode, tspan, y0, rtol=1e-3, atol=1e-6, stats=None):
    """[~, Y] = ode45(ode, [t0 tf], y0); returns Y(end, :).  stats (dict) receives nsteps / nfailed /
    nfevals / warned."""
    t0, tfinal = float(tspan[0]), float(tspan[1])
    y = np.array(y0, dtype=np.float64)
    neq = len(y)
    htspan = abs(tfinal - t0)
    hmax = abs(0.1 * (tfinal - t0))
    threshold = atol / rtol
    power = 1 / 5
    t = t0
    f = np.zeros((neq, 7))
    f0 = np.asarray(ode(t, y), dtype=np.float64)
    nfevals, nsteps, nfailed, warned = 1, 0, 0, False
    hmin = 16 * _eps(t)
    absh = min(hmax, htspan)
    rh = np.max(np.abs(f0 / np.maximum(np.abs(y), threshold))) / (0.8 * rtol ** power)
    if absh * rh > 1:
        absh = 1 / rh
    absh = max(absh, hmin)
    f[:, 0] = f0
    done = False
    while not done:
        hmin = 16 * _eps(t)
        absh = min(hmax, max(hmin, absh))
        h = absh
        if 1.1 * absh >= abs(tfinal - t):
            h = tfinal - t
            absh = abs(h)
            done = True
        nofailed = True
        while True:
            hA = h * A
            hB = h * B
            for k in range(5):
                f[:, k + 1] = ode(t + hA[k], y + f[:, :k + 1] @ hB[:k + 1, k])
            tnew = t + hA[5]
            if done:
                tnew = tfinal
            ynew = y + f[:, :6] @ hB[:6, 5]
            f[:, 6] = ode(tnew, ynew)
            nfevals += 6
            err = absh * np.max(np.abs((f @ E) / np.maximum(np.maximum(np.abs(y), np.abs(ynew)), threshold)))
            if err > rtol:
                nfailed += 1
                if absh <= hmin:
                    warned = True
                    if stats is not None:
                        stats.update(nsteps=nsteps, nfailed=nfailed, nfevals=nfevals, warned=warned)
                    return y
                if nofailed:
                    nofailed = False
                    absh = max(hmin, absh * max(0.1, 0.8 * (rtol / err) ** power))
                else:
                    absh = max(hmin, 0.5 * absh)
                h = absh
                done = False
            else:
                break
        nsteps += 1
        if done:
            y = ynew
            break
        if nofailed:
            temp = 1.25 * (err / rtol) ** power
            if temp > 0.2:
                absh = absh / temp
            else:
                absh = 5.0 * absh
        t = tnew
        y = ynew
        f[:, 0] = f[:, 6]
    if stats is not None:
        stats.update(nsteps=nsteps, nfailed=nfailed, nfevals=nfevals, warned=warned)
    return y


# ---- Curtis helpers (pos-att/private/*.m, bare-CR files) ---------------------------------------
def stumpC(z):                                                           # stumpC.m
    if z > 0:
        return (1 - math.cos(math.sqrt(z))) / z
    if z < 0:
        return (math.cosh(math.sqrt(-z)) - 1) / (-z)
    return 1 / 2


def stumpS(z):                                                           # stumpS.m
    if z > 0:
        return (math.sqrt(z) - math.sin(math.sqrt(z))) / math.sqrt(z) ** 3
    if z < 0:
        return (math.sinh(math.sqrt(-z)) - math.sqrt(-z)) / math.sqrt(-z) ** 3
    return 1 / 6


def kepler_U(dt, ro, vro, a, mu):                                        # kepler_U.m:25-44
    error, nMax = 1.e-8, 1000
    x = math.sqrt(mu) * abs(a) * dt
    n, ratio = 0, 1.0
    while abs(ratio) > error and n <= nMax:
        n += 1
        C = stumpC(a * x ** 2)
        S = stumpS(a * x ** 2)
        F = ro * vro / math.sqrt(mu) * x ** 2 * C + (1 - a * ro) * x ** 3 * S + ro * x - math.sqrt(mu) * dt
        dFdx = ro * vro / math.sqrt(mu) * x * (1 - a * x ** 2 * S) + (1 - a * ro) * x ** 2 * C + ro
        ratio = F / dFdx
        x = x - ratio
    return x


def update_RV_target(R0, V0, t, mu):                                     # Solver_pos_att.m:779-803
    r0 = np.linalg.norm(R0)
    v0 = np.linalg.norm(V0)
    vr0 = np.dot(R0, V0) / r0
    alpha = 2 / r0 - v0 ** 2 / mu
    x = kepler_U(t, r0, vr0, alpha, mu)
    z = alpha * x ** 2
    f = 1 - x ** 2 / r0 * stumpC(z)                                      # f_and_g.m
    g = t - 1 / math.sqrt(mu) * x ** 3 * stumpS(z)
    R2 = f * R0 + g * V0
    r2 = np.linalg.norm(R2)
    fdot = math.sqrt(mu) / r2 / r0 * (z * stumpS(z) - 1) * x             # fDot_and_gDot.m
    gdot = 1 - x ** 2 / r2 * stumpC(z)
    V2 = fdot * R0 + gdot * V0
    return R2, V2


def ECI2body(q):                                                         # :825-829
    return np.array([
        [1 - 2 * (q[1] ** 2 + q[2] ** 2), 2 * (q[0] * q[1] + q[2] * q[3]), 2 * (q[0] * q[2] - q[1] * q[3])],
        [2 * (q[1] * q[0] - q[2] * q[3]), 1 - 2 * (q[0] ** 2 + q[2] ** 2), 2 * (q[1] * q[2] + q[0] * q[3])],
        [2 * (q[2] * q[0] + q[1] * q[3]), 2 * (q[2] * q[1] - q[0] * q[3]), 1 - 2 * (q[0] ** 2 + q[1] ** 2)]])


def RSW2ECI(pos, vel):                                                   # :831-847
    R = pos / np.linalg.norm(pos)
    W = np.cross(pos, vel) / np.linalg.norm(np.cross(pos, vel))
    S = np.cross(W, R)
    return np.column_stack([R, S, W])


def angle2quat_zyx(r1, r2, r3):
    """Aerospace Toolbox angle2quat, default 'ZYX' sequence (scalar first)."""
    c = np.cos(np.array([r1, r2, r3]) / 2)
    s = np.sin(np.array([r1, r2, r3]) / 2)
    return np.array([c[0] * c[1] * c[2] + s[0] * s[1] * s[2], c[0] * c[1] * s[2] - s[0] * s[1] * c[2],
                     c[0] * s[1] * c[2] + s[0] * c[1] * s[2], s[0] * c[1] * c[2] - c[0] * s[1] * s[2]])


def default_X0_pos_att():                                                # :458-468
    dr0 = [-0.1, 0, 0]
    dv0 = [0, 0, 0]
    q0 = angle2quat_zyx(np.deg2rad(0), np.deg2rad(3), np.deg2rad(0))[::-1]
    w0 = [0, 0, 0]
    return np.concatenate([dr0, dv0, q0, w0])


class PosAttPlantLiteral:
    """controllers: {'x','y','z'} -> dict with GridVectors, U_Optimal_id (1-based, grid shaped),
    f0_allcomb, f1_allcomb, f6_allcomb, f7_allcomb (what calculate_one_channel_U_Opt saves, :291)."""

    def __init__(self, controllers, InertiaM, Mass, T_dist, h, R0, V0, mu=398600.0):
        self.I = np.asarray(InertiaM, dtype=np.float64)
        self.Mass, self.T_dist, self.h, self.mu = Mass, T_dist, h, mu
        self.R0, self.V0 = np.asarray(R0, dtype=np.float64), np.asarray(V0, dtype=np.float64)
        self.F = {}
        thr = {"x": (0, 1, 6, 7), "y": (2, 3, 8, 9), "z": (4, 5, 10, 11)}
        for ch, ctl in controllers.items():                              # set_controller :849-882
            uid = np.asarray(ctl["U_Optimal_id"]).astype(np.int64) - 1
            for name, k in zip(("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"), thr[ch]):
                self.F[k] = GriddedInterpolantNearest(ctl["GridVectors"], np.asarray(ctl[name]).ravel()[uid])

    def thrusters(self, x, v, t, w, q):                                  # :404-449
        x = ECI2body(q) @ (RSW2ECI(self.R0, self.V0) @ x)
        v = ECI2body(q) @ (RSW2ECI(self.R0, self.V0) @ v)
        f = np.zeros(12)
        for k in (0, 1, 6, 7):
            f[k] = self.F[k](x[0], v[0], t[1], w[1])
        for k in (2, 3, 8, 9):
            f[k] = self.F[k](x[1], v[1], t[2], w[2])
        for k in (4, 5, 10, 11):
            f[k] = self.F[k](x[2], v[2], t[0], w[0])
        return f

    def moments_forces(self, f, q):                                      # :805-823
        U_M_y = (f[0] - f[1] + f[6] - f[7]) * self.T_dist
        U_M_z = (f[2] - f[3] + f[8] - f[9]) * self.T_dist
        U_M_x = (f[4] - f[5] + f[10] - f[11]) * self.T_dist
        a_body = np.array([(f[0] + f[1] + f[6] + f[7]) / self.Mass, (f[2] + f[3] + f[8] + f[9]) / self.Mass,
                           (f[4] + f[5] + f[10] + f[11]) / self.Mass])
        acc = np.linalg.solve(RSW2ECI(self.R0, self.V0), np.linalg.solve(ECI2body(q), a_body))
        return np.array([U_M_x, U_M_y, U_M_z]), acc

    def rates(self, t, X, U_M, acc):                                     # :696-754
        mu = self.mu
        R, V = update_RV_target(self.R0, self.V0, t, mu)
        norm_R = (R @ R) ** .5
        RdotV = np.sum(R * V)
        crossRV = np.cross(R, V)
        H = (crossRV @ crossRV) ** .5
        x1, x2, x3, v1, v2, v3, q1, q2, q3, q4, w1, w2, w3 = X
        Xd = np.zeros(13)
        Xd[0:3] = v1, v2, v3
        Xd[3] = (2 * mu / norm_R ** 3 + H ** 2 / norm_R ** 4) * x1 - 2 * RdotV / norm_R ** 4 * H * x2 + 2 * H / norm_R ** 2 * v2 + acc[0]
        Xd[4] = -(mu / norm_R ** 3 - H ** 2 / norm_R ** 4) * x2 + 2 * RdotV / norm_R ** 4 * H * x1 - 2 * H / norm_R ** 2 * v1 + acc[1]
        Xd[5] = -mu / norm_R ** 3 * x3 + acc[2]
        Xd[6] = 0.5 * (w3 * q2 - w2 * q3 + w1 * q4)
        Xd[7] = 0.5 * (-w3 * q1 + w1 * q3 + w2 * q4)
        Xd[8] = 0.5 * (w2 * q1 - w1 * q2 + w3 * q4)
        Xd[9] = 0.5 * (-w1 * q1 - w2 * q2 - w3 * q3)
        wv = X[10:13]
        Xd[10:13] = np.linalg.solve(self.I, U_M - np.cross(wv, self.I @ wv))
        return Xd

    def get_optimal_path(self, X0, n_steps, stats=None):                 # :452-500
        X = np.zeros((n_steps + 1, 13))
        F_Th = np.zeros((n_steps, 12))
        FM = np.zeros((n_steps, 6))
        X[0] = X0
        for k in range(n_steps):
            Xs = X[k]
            t_stage = np.array([2 * math.asin(Xs[6]), 2 * math.asin(Xs[7]), 2 * math.asin(Xs[8])])
            q = Xs[6:10]
            f = self.thrusters(Xs[0:3], Xs[3:6], t_stage, Xs[10:13], q)
            U_M, acc = self.moments_forces(f, q)
            F_Th[k] = f
            FM[k] = np.concatenate([acc, U_M])
            st = {}
            X[k + 1] = ode45_last(lambda t, y: self.rates(t, y, U_M, acc), (k * self.h, (k + 1) * self.h), Xs, stats=st)
            if stats is not None:
                stats.setdefault("nsteps", []).append(st["nsteps"])
                stats.setdefault("nfailed", []).append(st["nfailed"])
        return X, F_Th, FM


class AttitudePlantLiteral:
    """Solver_attitude.get_optimal_path_simplified_testode45 (:1669-1705): FU = three 'nearest'
    interpolants over (w, theta)."""

    def __init__(self, FU, InertiaM, h):
        self.FU, self.I, self.h = FU, np.asarray(InertiaM, dtype=np.float64), h

    def rates(self, X, U):                                               # :1803-1849
        x1, x2, x3, x4, x5, x6, x7 = X
        w = X[0:3]
        Xd = np.zeros(7)
        Xd[0:3] = np.linalg.solve(self.I, U - np.cross(w, self.I @ w))
        Xd[3] = 0.5 * (x3 * x5 - x2 * x6 + x1 * x7)
        Xd[4] = 0.5 * (-x3 * x4 + x1 * x6 + x2 * x7)
        Xd[5] = 0.5 * (x2 * x4 - x1 * x5 + x3 * x7)
        Xd[6] = 0.5 * (-x1 * x4 - x2 * x5 - x3 * x6)
        return Xd

    def run(self, X0, n_steps):
        X = np.zeros((n_steps + 1, 7))
        Uh = np.zeros((n_steps, 3))
        X[0] = X0
        for k in range(n_steps):
            Xs = X[k]
            U = np.array([self.FU[a](Xs[a], 2 * math.asin(Xs[3 + a])) for a in range(3)])
            Uh[k] = U
            X[k + 1] = ode45_last(lambda t, y: self.rates(y, U), (k * self.h, (k + 1) * self.h), Xs)
        return X, Uh
