"""Array-at-a-time numpy restatement of the reference's four sweeps.  TEST INFRASTRUCTURE.

This module follows the MATLAB sources *literally*: it materialises the same S x C arrays
(``X_next_M1``, ``J_current_state``, ...), evaluates one N-linear interpolation over all S*C
queries, one add and one ``min(..., [], ctrl_dim)`` per stage — the shape of the reference's own
CPU path (it is also bench.py's "reference-shaped" CPU baseline).  It deliberately shares no code
with the product's table builders, so tests can check that the separable tables reproduce these
arrays bit for bit.

Reference lines:
  Dynamic_Solver   test/Dynamic_Solver.m:66-105,184-220 with the fp64 conventions of
                   test/test_coder.m:19-36,91-117 (J_star stored per stage)
  Solver_position  position-control/Solver_position.m:94-186,363-371
  Solver_attitude  attitude-control/Solver_attitude.m:196-259,625-667
  Solver_pos_att   pos-att/Solver_pos_att.m:197-402,784-802,886-918 (fp64; the reference's
                   single casts at :265,:800 are the unpinned 'single mode')

MATLAB runtime semantics restated (closed source; pinned only by test/obj_1.mat, R2016b/R2017a):
  linspace, ndgrid, griddedInterpolant(...,'linear') incl. linear extrapolation, min(X,[],dim)
  with first-index ties.  Interpolation formula used here: bin rule s[i] <= x < s[i+1],
  t = (x - s[i])/(s[i+1]-s[i]), dimension 1 first, (1-t)*lo + t*hi  — the variant SURVEY 4.4 found
  closest to MATLAB (83 % of entries bit-equal per stage, max 9.6e-16 relative).
"""
import numpy as np


# --- MATLAB built-ins -------------------------------------------------------------------------
def ml_linspace(d1, d2, n):
    n = int(n)
    n1 = n - 1
    y = d1 + (np.arange(0, n1 + 1, dtype=np.float64) * (d2 - d1)) / n1
    y[0] = d1
    y[-1] = d2
    return y


def ml_deg2rad(a):
    return (np.pi / 180) * a


def ml_ndgrid(*vecs):
    return np.meshgrid(*vecs, indexing="ij")


def ml_min_last(X):
    """[m, i] = min(X, [], ndims(X)); i is 1-based, first index wins ties (np.argmin does too)."""
    i = np.argmin(X, axis=-1)
    m = np.take_along_axis(X, i[..., None], axis=-1)[..., 0]
    return m, i + 1


class GriddedInterpolantLinear:
    """griddedInterpolant({s1,..,sD}, V, 'linear') with the default (linear) extrapolation."""

    def __init__(self, gridvecs, values):
        self.GridVectors = [np.asarray(g, dtype=np.float64) for g in gridvecs]
        self.Values = np.asarray(values, dtype=np.float64)

    def __call__(self, *queries):
        V = self.Values
        D = len(self.GridVectors)
        cells, ts = [], []
        for s, x in zip(self.GridVectors, queries):
            i = np.searchsorted(s, x, side="right") - 1        # bin rule s[i] <= x < s[i+1]
            i = np.clip(i, 0, len(s) - 2)
            t = (x - s[i]) / (s[i + 1] - s[i])
            cells.append(i)
            ts.append(t)
        # gather the 2^D corners, corner bit d = offset in dimension d
        vals = []
        for m in range(1 << D):
            idx = tuple(cells[d] + ((m >> d) & 1) for d in range(D))
            vals.append(V[idx])
        for d in range(D):                                      # dimension 1 first
            t = ts[d]
            vals = [(1 - t) * vals[2 * m] + t * vals[2 * m + 1] for m in range(len(vals) // 2)]
        return vals[0]


class GriddedInterpolantNearest:
    """griddedInterpolant(..., 'nearest'): nearest node, clamped outside; an exact midpoint goes to
    the upper node (MATLAB's tie side is undocumented — SURVEY Appendix A marks it unpinned)."""

    def __init__(self, gridvecs, values):
        self.GridVectors = [np.asarray(g, dtype=np.float64) for g in gridvecs]
        self.Values = np.asarray(values)

    def __call__(self, *queries):
        idx = []
        for s, x in zip(self.GridVectors, queries):
            i = np.clip(np.searchsorted(s, x, side="right") - 1, 0, len(s) - 2)
            up = (x - s[i]) >= (s[i + 1] - x)
            idx.append(i + up.astype(np.int64))
        return self.Values[tuple(idx)]


# --- Dynamic_Solver ---------------------------------------------------------------------------
class DynamicSolverLiteral:
    """test/Dynamic_Solver.m in fp64 (test/test_coder.m lineage, which test/obj_1.mat embodies)."""

    def __init__(self, **kw):
        self.Q = np.array([[0.25, 0.0], [0.0, 0.05]])
        self.A = np.array([[0.9974, 0.0539], [-0.1078, 1.1591]])
        self.B = np.array([[0.0013], [0.0539]])
        self.R = 0.05
        self.N = 200
        self.dx, self.du = 100, 1000
        self.x_max, self.x_min = 3.0, -2.5
        self.u_max, self.u_min = 10.0, -40.0
        for k, v in kw.items():
            setattr(self, k, v)

    def setup(self):
        A = np.asarray(self.A, dtype=np.float64).ravel(order="F")   # A(1..4) linear indexing
        B = np.asarray(self.B, dtype=np.float64).ravel(order="F")
        Q = np.asarray(self.Q, dtype=np.float64).ravel(order="F")
        self.s_r = ml_linspace(self.x_min, self.x_max, self.dx)                 # :69 (fp64)
        self.X1_mesh, self.X2_mesh = ml_ndgrid(self.s_r, self.s_r)              # :70
        self.U_mesh = ml_linspace(self.u_min, self.u_max, self.du)              # :72
        X1, X2, U = ml_ndgrid(self.s_r, self.s_r, self.U_mesh)                  # :75
        self.X_next_M1 = A[0] * X1 + A[2] * X2 + B[0] * U                       # :186
        self.X_next_M2 = A[1] * X1 + A[3] * X2 + B[1] * U                       # :187
        self.J_current_state = Q[0] * X1 ** 2 + Q[3] * X2 ** 2 + self.R * U ** 2  # :198-199
        self.F = GriddedInterpolantLinear([self.s_r, self.s_r], np.zeros(self.X1_mesh.shape))  # :83

    def run(self, n_stages=None):
        self.setup()
        N = int(self.N)
        self.J_star = np.zeros(self.X1_mesh.shape + (N,))
        self.u_star = np.zeros(self.X1_mesh.shape + (N,))
        self.u_star_idx_all = np.zeros(self.X1_mesh.shape + (N,), dtype=np.int32)
        todo = N - 1 if n_stages is None else n_stages
        for k in range(1, todo + 1):                                            # :86
            k_s = N - k                                                         # :88
            J_F_next = self.F(self.X_next_M1, self.X_next_M2)                   # :207
            self.F.Values, idx = ml_min_last(J_F_next + self.J_current_state)   # :209-210
            self.J_star[:, :, k_s - 1] = self.F.Values                          # test_coder.m:32
            self.u_star[:, :, k_s - 1] = self.U_mesh[idx - 1]                   # :100
            self.u_star_idx_all[:, :, k_s - 1] = idx
        return self

    def get_optimal_path(self, X0=(2.0, 1.0), mode="Nssu", ssu_num=1):
        """:108-145 — returns (X[2,N], U[N])."""
        N = int(self.N)
        A = np.asarray(self.A, dtype=np.float64)
        B = np.asarray(self.B, dtype=np.float64).ravel()
        X = np.zeros((2, N))
        U = np.zeros(N)
        X[:, 0] = X0
        for k in range(1, N):
            USM = self.u_star[:, :, (ssu_num if mode == "ssu" else k) - 1]
            Fu = GriddedInterpolantLinear([self.s_r, self.s_r], USM)
            U[k - 1] = Fu(np.array(X[0, k - 1]), np.array(X[1, k - 1]))
            X[:, k] = A @ X[:, k - 1] + B * U[k - 1]                            # :193
        return X, U


# --- Solver_position --------------------------------------------------------------------------
def _sym_linspace_position(a, b, n):                                            # :363-371
    m = int(np.ceil(n / 2)) + 1
    v1 = ml_linspace(a, 0.0, m)
    v2 = ml_linspace(0.0, b, m)
    return np.concatenate([v1, v2[1:]])


class SolverPositionLiteral:
    def __init__(self, **kw):
        self.v_min, self.v_max = -0.5, 0.5
        self.x_min, self.x_max = -0.5, 0.5
        self.n_mesh_v = self.n_mesh_x = 200
        self.Mass = 4.16
        self.Qx = [6.0, 6.0, 6.0]
        self.Qv = [6.0, 6.0, 6.0]
        self.R = [0.1, 0.1, 0.1]
        self.T_final, self.h = 30.0, 0.005
        for k, v in kw.items():
            setattr(self, k, v)
        self.N_stage = int(np.ceil(self.T_final / self.h))                      # :75-78
        self.U_vector = np.array([-0.13, 0.0, 0.13]) * 2                        # :84

    def axis_arrays(self, axis):
        h = self.h
        s_x = _sym_linspace_position(self.x_min, self.x_max, self.n_mesh_x)     # :97
        s_v = _sym_linspace_position(self.v_min, self.v_max, self.n_mesh_v)     # :101
        X, V, U = ml_ndgrid(s_x, s_v, self.U_vector)                            # :109
        J_current = self.Qx[axis] * X ** 2 + self.Qv[axis] * V ** 2 + self.R[axis] * U ** 2  # :113
        k1 = V                                                                  # RK4_x :157-167
        k2 = V + k1 * h / 2
        k3 = V + k2 * h / 2
        k4 = V + k3 * h
        x_next = X + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        k1 = U / self.Mass                                                      # RK4_v :173-186
        k2 = k3 = k4 = k1
        v_next = V + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        return (s_x, s_v), (x_next, v_next), J_current

    def simplified_run(self, axis=0, n_stages=None):
        """One axis of :94-150.  Returns (J at the last computed stage, idx (1-based), grids)."""
        grids, nxt, J_current = self.axis_arrays(axis)
        F = GriddedInterpolantLinear(grids, np.zeros((len(grids[0]), len(grids[1]))))   # :116
        todo = self.N_stage - 1 if n_stages is None else n_stages
        idx = None
        for _ in range(todo):                                                   # :132
            F.Values, idx = ml_min_last(J_current + F(*nxt))                    # :135
        return F.Values, idx, grids


# --- Solver_attitude.simplified_run -------------------------------------------------------------
class SolverAttitudeLiteral:
    def __init__(self, **kw):
        self.w_min = -ml_deg2rad(50.0)                                          # :106-107
        self.w_max = -ml_deg2rad(-50.0)
        self.n_mesh_w = 1000
        self.angle_min = [-30.0, -20.0, -35.0]                                  # yaw, pitch, roll
        self.angle_max = [30.0, 20.0, 35.0]
        self.n_mesh_t = 300
        inertia = [0.02836 + 0.00016, 0.026817 + 0.00150, 0.023 + 0.00150]      # :118-120
        self.J = inertia                                                        # J1..J3 :167-169
        self.Qw = [6.0, 6.0, 6.0]
        self.Qt = [6.0, 6.0, 6.0]
        self.R = [4.0, 4.0, 4.0]
        self.T_final, self.h = 30.0, 0.005
        self.U_vector = np.array([-0.11, 0.0, 0.11])                            # :174
        for k, v in kw.items():
            setattr(self, k, v)
        self.N_stage = int(np.ceil(self.T_final / self.h))

    def axis_arrays(self, axis):
        h = self.h
        s_w = ml_linspace(self.w_min, self.w_max, self.n_mesh_w)                # :199
        s_t = ml_linspace(ml_deg2rad(self.angle_min[axis]), ml_deg2rad(self.angle_max[axis]),
                          self.n_mesh_t)                                        # :203
        W, T, U = ml_ndgrid(s_w, s_t, self.U_vector)                            # :209
        J_current = self.Qw[axis] * W ** 2 + self.Qt[axis] * T ** 2 + self.R[axis] * U ** 2  # :220
        k1 = U / self.J[axis]                                                   # RK4_w :630-644
        k2 = k3 = k4 = k1
        w_next = W + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        k1 = W                                                                  # RK4_t :646-660
        k2 = W + k1 * h / 2
        k3 = W + k2 * h / 2
        k4 = W + k3 * h
        t_next = T + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        return (s_w, s_t), (w_next, t_next), J_current

    def simplified_run(self, axis=0, n_stages=None):
        grids, nxt, J_current = self.axis_arrays(axis)
        F = GriddedInterpolantLinear(grids, np.zeros((len(grids[0]), len(grids[1]))))   # :222
        todo = self.N_stage - 1 if n_stages is None else n_stages
        idx = None
        for _ in range(todo):                                                   # :236
            F.Values, idx = ml_min_last(J_current + F(*nxt))                    # :239
        return F.Values, idx, grids


# --- Solver_pos_att ---------------------------------------------------------------------------
def _sym_linspace_pos_att(a, b, n):                                             # :906-918
    half = int(np.ceil(n / 2))
    v1 = ml_linspace(a, 0.0, half + 1) if n % 2 == 0 else ml_linspace(a, 0.0, half)
    v2 = ml_linspace(0.0, b, half)
    return np.concatenate([v1, v2[1:]])


def _vectors_allcomb(f1, f2, f3, f4):                                           # :886-904
    g = ml_ndgrid(np.atleast_1d(f1).astype(float), np.atleast_1d(f2).astype(float),
                  np.atleast_1d(f3).astype(float), np.atleast_1d(f4).astype(float))
    f1, f2, f3, f4 = (x.ravel(order="F") for x in g)
    idrm1 = np.flatnonzero((f1 > 0) & (f3 < 0))
    idrm2 = np.flatnonzero((f2 > 0) & (f4 < 0))
    idrm = np.unique(np.concatenate([idrm1, idrm2]))
    keep = np.setdiff1d(np.arange(len(f1)), idrm)
    return f1[keep], f2[keep], f3[keep], f4[keep]


class SolverPosAttLiteral:
    def __init__(self, **kw):
        self.v_min, self.v_max, self.n_mesh_v = -0.1, 0.1, 30                   # :100-106
        self.x_min, self.x_max, self.n_mesh_x = -0.2, 0.2, 30
        self.w_min, self.w_max, self.n_mesh_w = ml_deg2rad(-2.0), ml_deg2rad(2.0), 15
        self.theta_min = [-5.0, -6.0, -7.0]
        self.theta_max = [5.0, 6.0, 7.0]
        self.n_mesh_t = 20
        self.Mass = 4.16
        inertia = [0.02836 + 0.00016, 0.026817 + 0.00150, 0.023 + 0.00150]
        self.J1, self.J2, self.J3 = inertia
        self.Qx, self.Qv = [6.0] * 3, [6.0] * 3
        self.Qt, self.Qw = [0.5] * 3, [0.5] * 3
        self.R = [0.1] * 3
        self.T_final, self.h = 10.0, 0.005
        self.T_dist = 9.65e-2
        self.Thruster_max_F = 0.13
        for k, v in kw.items():
            setattr(self, k, v)
        self.N_stage = int(np.ceil(self.T_final / self.h))

    def channel_arrays(self, ch):
        """Channel 0/1/2 = x/y/z of simplified_run :217-233 (inertia J2/J3/J1)."""
        h, d, T = self.h, self.T_dist, self.Thruster_max_F
        J = [self.J2, self.J3, self.J1][ch]
        s_x = _sym_linspace_pos_att(self.x_min, self.x_max, self.n_mesh_x)
        s_v = _sym_linspace_pos_att(self.v_min, self.v_max, self.n_mesh_v)
        s_t = _sym_linspace_pos_att(ml_deg2rad(self.theta_min[ch]), ml_deg2rad(self.theta_max[ch]),
                                    self.n_mesh_t)
        s_w = _sym_linspace_pos_att(self.w_min, self.w_max, self.n_mesh_w)
        fa = np.array([0.0, T])
        fb = -np.array([0.0, T])
        f1, f2, f6, f7 = _vectors_allcomb(fa, fa, fb, fb)                       # :253
        X = s_x.reshape(-1, 1, 1, 1, 1)                                         # :307-314
        V = s_v.reshape(1, -1, 1, 1, 1)
        Tt = s_t.reshape(1, 1, -1, 1, 1)
        W = s_w.reshape(1, 1, 1, -1, 1)
        F1, F2, F6, F7 = (f.reshape(1, 1, 1, 1, -1) for f in (f1, f2, f6, f7))
        x_next = X + h * V                                                      # :338
        v_next = V + h * ((F1 + F2 + F6 + F7) / self.Mass)                      # :354,:358
        t_next = Tt + h * W                                                     # :374
        w_next = W + h * ((F1 * d + F2 * (-d) + F6 * d + F7 * (-d)) / J)        # :395,:400-401
        full = (len(s_x), len(s_v), len(s_t), len(s_w), len(f1))
        nxt = tuple(np.broadcast_to(a, full) for a in (x_next, v_next, t_next, w_next))  # :324-327
        R = self.R[ch]
        J_current = (self.Qx[ch] * X ** 2 + self.Qv[ch] * V ** 2 + self.Qw[ch] * W ** 2
                     + self.Qt[ch] * Tt ** 2
                     + (R * F1 ** 2 + R * F2 ** 2 + R * F6 ** 2 + R * F7 ** 2))  # :800-801 (fp64)
        return (s_x, s_v, s_t, s_w), nxt, np.broadcast_to(J_current, full), (f1, f2, f6, f7)

    def calculate_one_channel(self, ch, n_stages=None, tol=1e-2, check=True):
        """:244-297.  idsum50_prev is defined as 0 initially (the reference reads it undefined)."""
        grids, nxt, J_current, _ = self.channel_arrays(ch)
        F = GriddedInterpolantLinear(grids, np.zeros(tuple(len(g) for g in grids)))   # :264
        fsum50_prev = 0.0
        idx = None
        k_stop = None
        log = []
        last = 1 if n_stages is None else self.N_stage - n_stages
        for k_s in range(self.N_stage - 1, last - 1, -1):                       # :270
            F.Values, idx = ml_min_last(J_current + F(*nxt))                    # :272
            k_stop = k_s
            if check and k_s % 50 == 0:                                         # :273-285
                fsum50 = float(np.sum(F.Values.ravel(order="F")))
                idsum50 = float(np.sum(idx))
                log.append((k_s, fsum50, idsum50))
                e = fsum50 - fsum50_prev
                fsum50_prev = fsum50
                if abs(e) < tol:
                    break
        return F.Values, idx, grids, k_stop, log


# --- simplified-plant simulation under the nearest policy ---------------------------------------
def simplified_axis_rollout_literal(s_a, s_b, U_opt_values, rate_dim, k_of_u, h, x0, n_steps):
    """attitude-control/test/test_simplified.m:129-151 for ONE axis, literally: every step builds
    griddedInterpolant({s_a, s_b}, U_Opt(:,:,k), 'nearest'), evaluates it at the state, then applies
    next_stage_states (:273-310).  ``U_opt_values`` is [n_a, n_b] (fixed policy) or
    [n_a, n_b, n_steps] (time varying, control VALUES like U*_Opt); ``rate_dim`` is the state the
    control drives (0 for attitude's (w, theta), 1 for position's (x, v)); ``k_of_u`` maps the
    control value to the slope of the rate state (U/J or U/Mass).  Returns X [n_steps+1, 2], U [n_steps]."""
    X = np.zeros((n_steps + 1, 2))
    U = np.zeros(n_steps)
    X[0] = x0
    r, o = rate_dim, 1 - rate_dim
    tv = np.ndim(U_opt_values) == 3
    for k in range(n_steps):
        FU = GriddedInterpolantNearest([s_a, s_b], U_opt_values[:, :, k] if tv else U_opt_values)
        u = float(FU(np.array(X[k, 0]), np.array(X[k, 1])))
        U[k] = u
        w = X[k, r]
        kk = k_of_u(u)                                                          # RK4_w: four equal slopes
        w_new = w + h * (kk + 2 * kk + 2 * kk + kk) / 6
        k1 = w                                                                  # RK4_t
        k2 = w + k1 * h / 2
        k3 = w + k2 * h / 2
        k4 = w + k3 * h
        t_new = X[k, o] + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        X[k + 1, r] = w_new
        X[k + 1, o] = t_new
    return X, U


# --- Solver_attitude.run: the coupled 6-D sweep -------------------------------------------------
class SolverAttitude6Literal:
    """attitude-control/Solver_attitude.m:521-601 (run), :1433-1485 (reshape_states), :629-685
    (calculate_J_current_state_fix_shaped), :825-925 (spacecraft_dynamics_taylor_estimate), :767-823
    (calculate_J_U_opt_state_M), line by line with the full 9-D arrays (w1 w2 w3 yaw pitch roll U1 U2 U3)
    the reference builds with repmat — usable at tiny meshes only.  The reference never ran this path
    (run passes k_s to a one-argument method, :282 vs :384; the default mesh needs 2.7e13-element
    arrays): what is restated is the code as written with that call fixed, fp64 instead of the single
    interpolant values.  Parity unpinned."""

    def __init__(self, n_mesh_w=4, n_mesh_q=3, h=0.005, N_stage=6, **kw):
        self.w_min, self.w_max = -ml_deg2rad(50), -ml_deg2rad(-50)
        self.n_mesh_w, self.n_mesh_q = n_mesh_w, n_mesh_q
        self.yaw_min, self.yaw_max, self.pitch_min, self.pitch_max, self.roll_min, self.roll_max = -30, 30, -20, 20, -35, 35
        self.J1, self.J2, self.J3 = 0.02836 + 0.00016, 0.026817 + 0.00150, 0.023 + 0.00150
        self.Q1 = self.Q2 = self.Q3 = self.Q4 = self.Q5 = self.Q6 = 6
        self.R1 = self.R2 = self.R3 = 4
        self.h, self.N_stage = h, N_stage
        self.U_vector = np.array([-0.11, 0, 0.11])
        for k, v in kw.items():
            setattr(self, k, v)
        nw, nq = self.n_mesh_w, self.n_mesh_q
        self.sr_1 = ml_linspace(self.w_min, self.w_max, nw)
        self.sr_2 = ml_linspace(self.w_min, self.w_max, nw)
        self.sr_3 = ml_linspace(self.w_min, self.w_max, nw)
        self.s_yaw = ml_linspace(ml_deg2rad(self.yaw_min), ml_deg2rad(self.yaw_max), nq)
        self.s_pitch = ml_linspace(ml_deg2rad(self.pitch_min), ml_deg2rad(self.pitch_max), nq)
        self.s_roll = ml_linspace(ml_deg2rad(self.roll_min), ml_deg2rad(self.roll_max), nq)

    def run(self, n_stages=None, J_N=None):
        nw, nq, nu = self.n_mesh_w, self.n_mesh_q, len(self.U_vector)
        sh = lambda v, ax: np.asarray(v, dtype=np.float64).reshape([-1 if k == ax else 1 for k in range(9)])
        X1V, X2V, X3V = sh(self.sr_1, 0), sh(self.sr_2, 1), sh(self.sr_3, 2)                    # reshape_states
        c4, s4 = sh(np.cos(self.s_yaw / 2), 3), sh(np.sin(self.s_yaw / 2), 3)
        c5, s5 = sh(np.cos(self.s_pitch / 2), 4), sh(np.sin(self.s_pitch / 2), 4)
        c6, s6 = sh(np.cos(self.s_roll / 2), 5), sh(np.sin(self.s_roll / 2), 5)
        U1V, U2V, U3V = sh(self.U_vector, 6), sh(self.U_vector, 7), sh(self.U_vector, 8)
        h, J1, J2, J3 = self.h, self.J1, self.J2, self.J3
        J_fix = (self.Q1 * X1V ** 2 + self.Q2 * X2V ** 2 + self.Q3 * X3V ** 2 +                 # :629-685
                 self.Q4 * (s4 * c5 * c6 - c4 * s5 * s6) ** 2 +
                 self.Q5 * (c4 * s5 * c6 + s4 * c5 * s6) ** 2 +
                 self.Q6 * (c4 * c5 * s6 - s4 * s5 * c6) ** 2 +
                 self.R1 * U1V ** 2 + self.R2 * U2V ** 2 + self.R3 * U3V ** 2)
        # spacecraft_dynamics_taylor_estimate :825-925
        x7 = (1 - ((s4 * c5 * c6 - c4 * s5 * s6) ** 2 + (c4 * s5 * c6 + s4 * c5 * s6) ** 2 +
                   (c4 * c5 * s6 - s4 * s5 * c6) ** 2)) ** 0.5
        X1n = X1V + h * ((J2 - J3) / J1 * X2V * X3V + U1V / J1)
        X2n = X2V + h * ((J3 - J1) / J2 * X3V * X1V + U2V / J2)
        X3n = X3V + h * ((J1 - J2) / J3 * X1V * X2V + U3V / J3)
        X4n = (s4 * c5 * c6 - c4 * s5 * s6) + h * (0.5 * (X3V * (c4 * s5 * c6 + s4 * c5 * s6)
                                                          - X2V * (c4 * c5 * s6 - s4 * s5 * c6) + X1V * x7))
        X5n = (c4 * s5 * c6 + s4 * c5 * s6) + h * (0.5 * (-X3V * (s4 * c5 * c6 - c4 * s5 * s6)
                                                          + X1V * (c4 * c5 * s6 - s4 * s5 * c6) + X2V * x7))
        X6n = (c4 * c5 * s6 - s4 * s5 * c6) + h * (0.5 * (X2V * (s4 * c5 * c6 - c4 * s5 * s6)
                                                          - X1V * (c4 * s5 * c6 + s4 * c5 * s6) + X3V * x7))
        x7 = x7 + h * (0.5 * (-X1V * (s4 * c5 * c6 - c4 * s5 * s6) - X2V * (c4 * s5 * c6 + s4 * c5 * s6)
                              - X3V * (c4 * c5 * s6 - s4 * s5 * c6)))
        Qs = np.sqrt(X4n ** 2 + X5n ** 2 + X6n ** 2 + x7 ** 2)
        X4n, X5n, X6n, x7 = X4n / Qs, X5n / Qs, X6n / Qs, x7 / Qs
        r1 = np.arctan2(2. * (X6n * X5n + x7 * X4n), x7 ** 2 + X6n ** 2 - X5n ** 2 - X4n ** 2)
        r2 = np.arcsin(-2. * (X6n * X4n - x7 * X5n))
        r3 = np.arctan2(2. * (X5n * X4n + x7 * X6n), x7 ** 2 - X6n ** 2 - X5n ** 2 + X4n ** 2)
        full = (nw, nw, nw, nq, nq, nq, nu, nu, nu)
        rep = lambda a: np.broadcast_to(a, full)                                               # the repmat calls :905-921
        Xq = [rep(X1n), rep(X2n), rep(X3n), rep(r1), rep(r2), rep(r3)]
        grids = [self.sr_1, self.sr_2, self.sr_3, self.s_yaw, self.s_pitch, self.s_roll]
        F_values = np.zeros(full[:6]) if J_N is None else np.asarray(J_N, dtype=np.float64).reshape(full[:6], order="F")
        todo = self.N_stage - 1 if n_stages is None else n_stages
        for _ in range(todo):                                                                   # :549-555
            F = GriddedInterpolantLinear(grids, F_values)
            tot = J_fix + F(*Xq)
            val, U3 = ml_min_last(tot)                                                          # min(.., [], dim_U3)
            val, U2 = ml_min_last(val)                                                          # dim_U2
            F_values, U1 = ml_min_last(val)                                                     # dim_U1
        # intended composition of :557-559 (the reference indexes without the state subscripts)
        U2s = np.take_along_axis(U2, (U1 - 1)[..., None], axis=-1)[..., 0]
        U3s = np.take_along_axis(np.take_along_axis(U3, (U1 - 1)[..., None, None], axis=-2)[..., 0, :],
                                 (U2s - 1)[..., None], axis=-1)[..., 0]
        return F_values, U1, U2s, U3s


def attitude6_get_optimal_path_literal(grids, U_opt_values, J123, h, X0, n_steps):
    """Solver_attitude.get_optimal_path (:1487-1530) with method 'nearest': U_opt_values = the three
    U{1,2,3}_Opt value arrays (grid shaped).  quat2angle = Aerospace Toolbox, 'ZYX', input normalised."""
    J1, J2, J3 = J123
    FU = [GriddedInterpolantNearest(grids, np.asarray(v)) for v in U_opt_values]
    X = np.zeros((7, n_steps + 1))
    U = np.zeros((3, n_steps))
    X[:, 0] = X0
    for k in range(n_steps):
        q = np.array([X[6, k], X[5, k], X[4, k], X[3, k]])
        q = q / np.sqrt(np.sum(q ** 2))                                     # quatnormalize
        yaw = np.arctan2(2 * (q[1] * q[2] + q[0] * q[3]), q[0] ** 2 + q[1] ** 2 - q[2] ** 2 - q[3] ** 2)
        pitch = np.arcsin(-2 * (q[1] * q[3] - q[0] * q[2]))
        roll = np.arctan2(2 * (q[2] * q[3] + q[0] * q[1]), q[0] ** 2 - q[1] ** 2 - q[2] ** 2 + q[3] ** 2)
        for a in range(3):
            U[a, k] = FU[a](X[0, k], X[1, k], X[2, k], yaw, pitch, roll)
        x1, x2, x3, x4, x5, x6, x7 = X[:, k]
        u1, u2, u3 = U[:, k]
        Xd = np.array([(J2 - J3) / J1 * x2 * x3 + u1 / J1, (J3 - J1) / J2 * x3 * x1 + u2 / J2,     # :1199-1245
                       (J1 - J2) / J3 * x1 * x2 + u3 / J3,
                       0.5 * (x3 * x5 - x2 * x6 + x1 * x7), 0.5 * (-x3 * x4 + x1 * x6 + x2 * x7),
                       0.5 * (x2 * x4 - x1 * x5 + x3 * x7), 0.5 * (-x1 * x4 - x2 * x5 - x3 * x6)])
        X2 = X[:, k] + h * Xd                                                # 'taylor' :1366-1367
        X2[3:7] = X2[3:7] / np.sqrt(X2[3] ** 2 + X2[4] ** 2 + X2[5] ** 2 + X2[6] ** 2)
        X[:, k + 1] = X2
    return X, U
