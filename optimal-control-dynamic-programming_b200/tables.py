"""Host-side table builders: turn each reference class's settings into the separable descriptor
the stage kernel consumes (include/bellman.h, ``bellman_desc``).

Every S x C array the reference precomputes is a sum of 1-D tables.  The 1-D tables are evaluated
here with the reference's own operation order (MATLAB evaluates ``a + b + c`` as ``(a+b)+c``,
``h*(...)/6`` as ``(h*(...))/6``, ``x.^2`` as ``x.*x``, each elementwise op rounded separately),
so the sums the kernel forms are bit-identical to the reference's array elements.

Reference lines followed:
  Dynamic_Solver   test/Dynamic_Solver.m:69-84 (grids), :184-188 (a_D_M), :196-200 (g_D)
  Solver_position  position-control/Solver_position.m:97-128, :152-186, :363-371
  Solver_attitude  attitude-control/Solver_attitude.m:199-233, :625-667
  Solver_pos_att   pos-att/Solver_pos_att.m:200-263, :299-402, :784-802, :886-918
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAX_DIM = 4


# ----------------------------------------------------------------------------------------------
# MATLAB built-in semantics needed by the builders
# ----------------------------------------------------------------------------------------------
def linspace(a, b, n):
    """MATLAB ``linspace`` (R2016b lineage): ``a + ((0:n-1)*(b-a))/(n-1)`` with forced end points.
    Reproduces ``X1_mesh(:,1)`` of test/obj_1.mat bit for bit."""
    a = float(a)
    b = float(b)
    n = int(n)
    if n < 2:
        return np.array([b], dtype=np.float64)
    i = np.arange(n, dtype=np.float64)
    y = a + (i * (b - a)) / float(n - 1)
    y[0] = a
    y[-1] = b
    return y


def deg2rad(x):
    """MATLAB ``deg2rad``: ``(pi/180) * x``."""
    return (np.pi / 180.0) * np.asarray(x, dtype=np.float64)


def sym_linspace_position(a, b, n):
    """Solver_position.sym_linspace (position-control/Solver_position.m:363-371)."""
    if a > 0:
        raise ValueError("minimum states are not negative, use normal linspace")
    m = int(np.ceil(n / 2)) + 1
    v1 = linspace(a, 0.0, m)
    v2 = linspace(0.0, b, m)[1:]
    return np.concatenate([v1, v2])


def sym_linspace_pos_att(a, b, n):
    """Solver_pos_att.sym_linspace (pos-att/Solver_pos_att.m:906-918): even n gives two spacings."""
    if a > 0:
        raise ValueError("minimum states are not negative, use normal linspace")
    h = int(np.ceil(n / 2))
    v1 = linspace(a, 0.0, h + 1 if n % 2 == 0 else h)
    v2 = linspace(0.0, b, h)[1:]
    return np.concatenate([v1, v2])


def vectors_allcomb(f1, f2, f3, f4):
    """Solver_pos_att.vectors_allcomb (pos-att/Solver_pos_att.m:886-904): ndgrid of the four
    on/off vectors (first fastest), minus combinations firing opposing pairs (f1>0 & f3<0, or
    f2>0 & f4<0)."""
    f1 = np.atleast_1d(np.asarray(f1, dtype=np.float64))
    f2 = np.atleast_1d(np.asarray(f2, dtype=np.float64))
    f3 = np.atleast_1d(np.asarray(f3, dtype=np.float64))
    f4 = np.atleast_1d(np.asarray(f4, dtype=np.float64))
    g1, g2, g3, g4 = np.meshgrid(f1, f2, f3, f4, indexing="ij")
    g1, g2, g3, g4 = (g.ravel(order="F") for g in (g1, g2, g3, g4))
    drop = ((g1 > 0) & (g3 < 0)) | ((g2 > 0) & (g4 < 0))
    keep = ~drop
    return g1[keep], g2[keep], g3[keep], g4[keep]


# ----------------------------------------------------------------------------------------------
# descriptor
# ----------------------------------------------------------------------------------------------
@dataclass
class Desc:
    """Python image of ``bellman_desc``: every table is a float64 array with leading axis P."""
    n: List[int]
    C: int
    N: int
    grid: List[np.ndarray]
    src_a: List[int]
    src_b: List[int]
    Ta: List[np.ndarray]
    Tb: List[Optional[np.ndarray]]
    Tc: List[Optional[np.ndarray]]
    q_order: List[int]
    q: List[np.ndarray]
    r: np.ndarray
    P: int = 1
    store_J_all: bool = False
    store_idx_all: bool = False
    meta: dict = field(default_factory=dict)

    @property
    def D(self):
        return len(self.n)

    @property
    def S(self):
        return int(np.prod(self.n))

    def validate(self):
        D, P, C = self.D, self.P, self.C
        assert 2 <= D <= MAX_DIM
        for d in range(D):
            assert self.grid[d].shape == (P, self.n[d]), (d, self.grid[d].shape)
            assert np.all(np.diff(self.grid[d], axis=1) > 0), "grid must be strictly increasing"
            assert self.Ta[d].shape == (P, self.n[self.src_a[d]])
            if self.Tb[d] is not None:
                assert self.src_b[d] >= 0 and self.Tb[d].shape == (P, self.n[self.src_b[d]])
            if self.Tc[d] is not None:
                assert self.Tc[d].shape == (P, C)
            assert self.q[d].shape == (P, self.n[d])
        assert sorted(self.q_order) == list(range(D))
        assert self.r.shape == (P, C)
        return self


def _row(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(1, -1))


def stack_problems(descs):
    """Stack single-problem descriptors of identical shape into one batched descriptor."""
    d0 = descs[0]
    for d in descs[1:]:
        assert d.n == d0.n and d.C == d0.C and d.N == d0.N
        assert d.src_a == d0.src_a and d.src_b == d0.src_b and d.q_order == d0.q_order

    def cat(get):
        xs = [get(d) for d in descs]
        if xs[0] is None:
            assert all(x is None for x in xs)
            return None
        return np.ascontiguousarray(np.concatenate(xs, axis=0))

    D = d0.D
    return Desc(
        n=list(d0.n), C=d0.C, N=d0.N,
        grid=[cat(lambda d, k=k: d.grid[k]) for k in range(D)],
        src_a=list(d0.src_a), src_b=list(d0.src_b),
        Ta=[cat(lambda d, k=k: d.Ta[k]) for k in range(D)],
        Tb=[cat(lambda d, k=k: d.Tb[k]) for k in range(D)],
        Tc=[cat(lambda d, k=k: d.Tc[k]) for k in range(D)],
        q_order=list(d0.q_order),
        q=[cat(lambda d, k=k: d.q[k]) for k in range(D)],
        r=cat(lambda d: d.r),
        P=sum(d.P for d in descs),
        store_J_all=d0.store_J_all, store_idx_all=d0.store_idx_all,
        meta={"problems": [d.meta for d in descs]},
    ).validate()


# ----------------------------------------------------------------------------------------------
# Dynamic_Solver (Kirk ch.3 two-state linear regulator)
# ----------------------------------------------------------------------------------------------
def kirk_desc(A, B, Q, R, N, x_min, x_max, dx, u_min, u_max, du,
              store_J_all=True, store_idx_all=True):
    """test/Dynamic_Solver.m:69-84,184-200 in fp64 (the test/test_coder.m:19-36 convention):
       Xn1 = A(1)*X1 + A(3)*X2 + B(1)*U ; Xn2 = A(2)*X1 + A(4)*X2 + B(2)*U   (column-major A)
       J_current = Q(1)*X1.^2 + Q(4)*X2.^2 + R*U.^2 ;  J_N = 0 (H is unused)."""
    A = np.asarray(A, dtype=np.float64).reshape(2, 2)
    B = np.asarray(B, dtype=np.float64).reshape(2)
    Q = np.asarray(Q, dtype=np.float64).reshape(2, 2)
    R = float(np.asarray(R).reshape(()))
    s = linspace(x_min, x_max, dx)
    u = linspace(u_min, u_max, du)
    # MATLAB linear indexing: A(1)=A11, A(2)=A21, A(3)=A12, A(4)=A22; Q(1)=Q11, Q(4)=Q22
    return Desc(
        n=[int(dx), int(dx)], C=int(du), N=int(N),
        grid=[_row(s), _row(s)],
        src_a=[0, 0], src_b=[1, 1],
        Ta=[_row(A[0, 0] * s), _row(A[1, 0] * s)],
        Tb=[_row(A[0, 1] * s), _row(A[1, 1] * s)],
        Tc=[_row(B[0] * u), _row(B[1] * u)],
        q_order=[0, 1],
        q=[_row(Q[0, 0] * (s * s)), _row(Q[1, 1] * (s * s))],
        r=_row(R * (u * u)),
        store_J_all=store_J_all, store_idx_all=store_idx_all,
        meta={"class": "Dynamic_Solver", "s_r": s, "U_mesh": u, "A": A, "B": B},
    ).validate()


# ----------------------------------------------------------------------------------------------
# Solver_position (one (x, v) axis)
# ----------------------------------------------------------------------------------------------
def _rk4_increment_of_rate(rate, h):
    """``h*(k1 + 2*k2 + 2*k3 + k4)/6`` with k1 = rate, k2 = rate + k1*h/2, k3 = rate + k2*h/2,
    k4 = rate + k3*h  (Solver_position.m:157-167, Solver_attitude.m:646-660)."""
    k1 = rate
    k2 = rate + (k1 * h) / 2
    k3 = rate + (k2 * h) / 2
    k4 = rate + k3 * h
    return (h * (((k1 + 2 * k2) + 2 * k3) + k4)) / 6


def _rk4_increment_const(k, h):
    """Same formula when all four slopes equal ``k`` (Solver_position.m:173-186,
    Solver_attitude.m:630-644)."""
    return (h * (((k + 2 * k) + 2 * k) + k)) / 6


def position_axis_desc(x_min, x_max, n_mesh_x, v_min, v_max, n_mesh_v, U_vector, Mass,
                       Qx, Qv, R, h, N_stage):
    """position-control/Solver_position.m:97-128: dims (x, v);
       x_next = X + h*(k1+2k2+2k3+k4)/6 from V (control independent);  v_next = V + incr(U/Mass);
       J_current = Qx*x.^2 + Qv*v.^2 + R*U.^2."""
    s_x = sym_linspace_position(x_min, x_max, n_mesh_x)
    s_v = sym_linspace_position(v_min, v_max, n_mesh_v)
    U = np.asarray(U_vector, dtype=np.float64).ravel()
    return Desc(
        n=[len(s_x), len(s_v)], C=len(U), N=int(N_stage),
        grid=[_row(s_x), _row(s_v)],
        src_a=[0, 1], src_b=[1, -1],
        Ta=[_row(s_x), _row(s_v)],
        Tb=[_row(_rk4_increment_of_rate(s_v, h)), None],
        Tc=[None, _row(_rk4_increment_const(U / Mass, h))],
        q_order=[0, 1],
        q=[_row(Qx * (s_x * s_x)), _row(Qv * (s_v * s_v))],
        r=_row(R * (U * U)),
        meta={"class": "Solver_position", "s_x": s_x, "s_v": s_v, "U_vector": U},
    ).validate()


# ----------------------------------------------------------------------------------------------
# Solver_attitude.simplified_run (one (w, theta) axis)
# ----------------------------------------------------------------------------------------------
def attitude_axis_desc(w_min, w_max, n_mesh_w, t_min_deg, t_max_deg, n_mesh_t, U_vector, J_axis,
                       Qw, Qt, R, h, N_stage):
    """attitude-control/Solver_attitude.m:199-233,625-667: dims (w, theta);
       w_next = w + incr(U/J);  t_next = T + h*(k1+2k2+2k3+k4)/6 from W;
       J_current = Qw*w.^2 + Qt*theta.^2 + R*U.^2."""
    s_w = linspace(w_min, w_max, n_mesh_w)
    s_t = linspace(deg2rad(t_min_deg), deg2rad(t_max_deg), n_mesh_t)
    U = np.asarray(U_vector, dtype=np.float64).ravel()
    return Desc(
        n=[len(s_w), len(s_t)], C=len(U), N=int(N_stage),
        grid=[_row(s_w), _row(s_t)],
        src_a=[0, 1], src_b=[-1, 0],
        Ta=[_row(s_w), _row(s_t)],
        Tb=[None, _row(_rk4_increment_of_rate(s_w, h))],
        Tc=[_row(_rk4_increment_const(U / J_axis, h)), None],
        q_order=[0, 1],
        q=[_row(Qw * (s_w * s_w)), _row(Qt * (s_t * s_t))],
        r=_row(R * (U * U)),
        meta={"class": "Solver_attitude", "s_w": s_w, "s_t": s_t, "U_vector": U},
    ).validate()


# ----------------------------------------------------------------------------------------------
# Solver_pos_att (one (x, v, theta, w) channel)
# ----------------------------------------------------------------------------------------------
def pos_att_channel_desc(s_x, s_v, s_t, s_w, f0, f1, f6, f7, Qx, Qv, Qt, Qw, R, J_axis,
                         Mass, T_dist, h, N_stage):
    """pos-att/Solver_pos_att.m:244-265,299-402,784-802 in fp64 (the reference casts J_current
    and F.Values to single, SURVEY 'single mode' = unpinned; fp64 is the canonical build):
       x' = X + h*V ; v' = V + h*((f1+f2+f6+f7)/Mass) ; t' = T + h*W ;
       w' = W + h*((f1*d + f2*(-d) + f6*d + f7*(-d))/J)
       J_current = Qx x^2 + Qv v^2 + Qw w^2 + Qt t^2 + (R f1^2 + R f2^2 + R f3^2 + R f4^2)."""
    s_x, s_v, s_t, s_w = (np.asarray(a, dtype=np.float64).ravel() for a in (s_x, s_v, s_t, s_w))
    c0, c1, c6, c7 = vectors_allcomb(f0, f1, f6, f7)
    d = float(T_dist)
    v_dot = (((c0 + c1) + c6) + c7) / Mass
    w_dot = (((c0 * d + c1 * (-d)) + c6 * d) + c7 * (-d)) / J_axis
    r = ((R * (c0 * c0) + R * (c1 * c1)) + R * (c6 * c6)) + R * (c7 * c7)
    return Desc(
        n=[len(s_x), len(s_v), len(s_t), len(s_w)], C=len(c0), N=int(N_stage),
        grid=[_row(s_x), _row(s_v), _row(s_t), _row(s_w)],
        src_a=[0, 1, 2, 3], src_b=[1, -1, 3, -1],
        Ta=[_row(s_x), _row(s_v), _row(s_t), _row(s_w)],
        Tb=[_row(h * s_v), None, _row(h * s_w), None],
        Tc=[None, _row(h * v_dot), None, _row(h * w_dot)],
        q_order=[0, 1, 3, 2],          # Qx x^2 + Qv v^2 + Qw w^2 + Qt t^2  (Solver_pos_att.m:800)
        q=[_row(Qx * (s_x * s_x)), _row(Qv * (s_v * s_v)),
           _row(Qt * (s_t * s_t)), _row(Qw * (s_w * s_w))],
        r=_row(r),
        meta={"class": "Solver_pos_att", "f0_allcomb": c0, "f1_allcomb": c1,
              "f6_allcomb": c6, "f7_allcomb": c7},
    ).validate()


# ----------------------------------------------------------------------------------------------
# Solver_attitude.run — the coupled 6-D problem (w1 w2 w3 yaw pitch roll) x 3 controls
# ----------------------------------------------------------------------------------------------
@dataclass
class Dense6:
    """Tables of the 6-D attitude sweep (attitude-control/Solver_attitude.m:521-601): the next state is
    NOT a sum of 1-D tables (Euler's equations couple the three rates, the angle update goes through a
    quaternion), so the next-state arrays the reference precomputes are kept — but only once each,
    without the repmat over the dimensions they do not depend on:
      w_next[d]  [nu, n0*n1*n2]   X{d}_next: depends on (w1, w2, w3) and on control d only   (:829-833)
      a_next[d]  [S]              X{4,5,6}_next: depends on the state only                    (:835-897)
      gs         [S]              state part of J_current_state_fix                           (:629-685)
      r[d]       [nu]             R_d * U_d^2
    S = prod(n), dimension 0 (w1) fastest."""
    n: List[int]
    nu: int
    N: int
    grid: List[np.ndarray]
    w_next: List[np.ndarray]
    a_next: List[np.ndarray]
    gs: np.ndarray
    r: List[np.ndarray]
    U_vector: np.ndarray

    @property
    def S(self):
        return int(np.prod(self.n))


def attitude6_tables(sr, s_yaw, s_pitch, s_roll, U_vector, J1, J2, J3, Q, R, h, N):
    """reshape_states (:1433-1485), calculate_J_current_state_fix_shaped (:629-685) and
    spacecraft_dynamics_taylor_estimate (:825-925) with the reference's operation order, array at a time
    (numpy broadcasting = MATLAB implicit expansion); sr = (sr_1, sr_2, sr_3), Q = (Q1..Q6), R = (R1, R2, R3)."""
    f = lambda a: np.asarray(a, dtype=np.float64)
    sr = [f(a) for a in sr]
    s_yaw, s_pitch, s_roll, U = f(s_yaw), f(s_pitch), f(s_roll), f(U_vector)
    X1V = sr[0].reshape(-1, 1, 1, 1, 1, 1)
    X2V = sr[1].reshape(1, -1, 1, 1, 1, 1)
    X3V = sr[2].reshape(1, 1, -1, 1, 1, 1)
    c4, s4 = np.cos(s_yaw / 2).reshape(1, 1, 1, -1, 1, 1), np.sin(s_yaw / 2).reshape(1, 1, 1, -1, 1, 1)
    c5, s5 = np.cos(s_pitch / 2).reshape(1, 1, 1, 1, -1, 1), np.sin(s_pitch / 2).reshape(1, 1, 1, 1, -1, 1)
    c6, s6 = np.cos(s_roll / 2).reshape(1, 1, 1, 1, 1, -1), np.sin(s_roll / 2).reshape(1, 1, 1, 1, 1, -1)
    qa = s4 * c5 * c6 - c4 * s5 * s6
    qb = c4 * s5 * c6 + s4 * c5 * s6
    qc = c4 * c5 * s6 - s4 * s5 * c6
    Q1, Q2, Q3, Q4, Q5, Q6 = Q
    n = [len(sr[0]), len(sr[1]), len(sr[2]), len(s_yaw), len(s_pitch), len(s_roll)]
    gs = Q1 * X1V ** 2 + Q2 * X2V ** 2 + Q3 * X3V ** 2 + Q4 * qa ** 2 + Q5 * qb ** 2 + Q6 * qc ** 2       # :629-640
    gs = np.broadcast_to(gs, n)
    # :826-828
    x7 = (1 - (qa ** 2 + qb ** 2 + qc ** 2)) ** 0.5
    U1 = U.reshape(1, 1, 1, -1)
    w3d = (X1V[..., 0, 0, 0], X2V[..., 0, 0, 0], X3V[..., 0, 0, 0])           # [n0,1,1], [1,n1,1], [1,1,n2]
    Xa, Xb, Xc = (a[..., None] for a in w3d)
    w_next = [np.broadcast_to(Xa + h * ((J2 - J3) / J1 * Xb * Xc + U1 / J1), n[:3] + [len(U)]),              # :829-833
              np.broadcast_to(Xb + h * ((J3 - J1) / J2 * Xc * Xa + U1 / J2), n[:3] + [len(U)]),
              np.broadcast_to(Xc + h * ((J1 - J2) / J3 * Xa * Xb + U1 / J3), n[:3] + [len(U)])]
    X4n = qa + h * (0.5 * (X3V * qb - X2V * qc + X1V * x7))                                                   # :835-849
    X5n = qb + h * (0.5 * (-X3V * qa + X1V * qc + X2V * x7))
    X6n = qc + h * (0.5 * (X2V * qa - X1V * qb + X3V * x7))
    x7n = x7 + h * (0.5 * (-X1V * qa - X2V * qb - X3V * qc))
    Qs = np.sqrt(X4n ** 2 + X5n ** 2 + X6n ** 2 + x7n ** 2)                                                  # :865-873
    X4n, X5n, X6n, x7n = X4n / Qs, X5n / Qs, X6n / Qs, x7n / Qs
    yaw = np.arctan2(2. * (X6n * X5n + x7n * X4n), x7n ** 2 + X6n ** 2 - X5n ** 2 - X4n ** 2)                # :877-885
    pitch = np.arcsin(-2. * (X6n * X4n - x7n * X5n))
    roll = np.arctan2(2. * (X5n * X4n + x7n * X6n), x7n ** 2 - X6n ** 2 - X5n ** 2 + X4n ** 2)
    flatF = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))
    R1, R2, R3 = R
    return Dense6(
        n=n, nu=len(U), N=int(N), grid=[sr[0], sr[1], sr[2], s_yaw, s_pitch, s_roll],
        w_next=[np.ascontiguousarray(np.moveaxis(w, 3, 0).reshape(len(U), -1, order="F")) for w in w_next],
        a_next=[flatF(yaw), flatF(pitch), flatF(roll)], gs=flatF(gs),
        r=[f(R1 * U ** 2), f(R2 * U ** 2), f(R3 * U ** 2)], U_vector=U)
