"""ctypes binding of oracle/libbellman_oracle.so (the C restatement).  TEST INFRASTRUCTURE.

Takes any object with the attributes of the product's ``tables.Desc`` (n, C, P, N, grid, src_a,
src_b, Ta, Tb, Tc, q_order, q, r) — duck-typed so this package does not import the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAXD = 4
_dp = C.POINTER(C.c_double)


class CDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("D", C.c_int32), ("n", C.c_int32 * MAXD), ("C", C.c_int32),
        ("P", C.c_int32), ("N", C.c_int32), ("grid", _dp * MAXD),
        ("src_a", C.c_int32 * MAXD), ("src_b", C.c_int32 * MAXD),
        ("Ta", _dp * MAXD), ("Tb", _dp * MAXD), ("Tc", _dp * MAXD),
        ("q_order", C.c_int32 * MAXD), ("q", _dp * MAXD), ("r", _dp),
        ("store_J_all", C.c_int32), ("store_idx_all", C.c_int32), ("device", C.c_int32),
        ("part_dim", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
        ("part_cuts", C.POINTER(C.c_int32)), ("idx_bytes", C.c_int32),
    ]


def build(force=False):
    so = os.path.join(HERE, "libbellman_oracle.so")
    src = os.path.join(HERE, "bellman_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libbellman_oracle.so")
        if not os.path.exists(so):
            build()
        _lib = C.CDLL(so)
    return _lib


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def to_cdesc(d):
    """Returns (CDesc, keepalive list)."""
    cd = CDesc()
    keep = []
    D = len(d.n)
    cd.struct_size = C.sizeof(CDesc)
    cd.D, cd.C, cd.P, cd.N = D, int(d.C), int(d.P), int(d.N)
    for k in range(D):
        cd.n[k] = int(d.n[k])
        cd.src_a[k] = int(d.src_a[k])
        cd.src_b[k] = int(d.src_b[k])
        cd.q_order[k] = int(d.q_order[k])
        for name in ("grid", "Ta", "Tb", "Tc", "q"):
            a = getattr(d, name)[k]
            if a is None:
                getattr(cd, name)[k] = _dp()
            else:
                a = _arr(a)
                keep.append(a)
                getattr(cd, name)[k] = a.ctypes.data_as(_dp)
    r = _arr(d.r)
    keep.append(r)
    cd.r = r.ctypes.data_as(_dp)
    cd.device, cd.part_dim, cd.rank, cd.nranks = -1, -1, 0, 1
    return cd, keep


def locate_modes(d):
    cd, keep = to_cdesc(d)
    modes = np.zeros(d.P * len(d.n), dtype=np.int32)
    lib().oracle_locate_modes(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)))
    return modes


def stage(d, J_next, modes=None, part_dim=-1, own_lo=0, own_hi=0, nthreads=0):
    """One backward stage.  J_next: [P, S] (S column-major flattened).  Returns (J, idx0)."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    S = int(np.prod(d.n))
    Jn = _arr(J_next).reshape(d.P, S)
    Jo = np.zeros_like(Jn)
    Io = np.zeros((d.P, S), dtype=np.int32)
    rc = lib().oracle_stage(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)),
                            Jn.ctypes.data_as(_dp), Jo.ctypes.data_as(_dp),
                            Io.ctypes.data_as(C.POINTER(C.c_int32)),
                            C.c_int(part_dim), C.c_int(own_lo), C.c_int(own_hi), C.c_int(nthreads))
    assert rc == 0
    return Jo, Io


def sweep(d, n_stages=None, J_N=None, keep_all=False, check_period=0, check_tol=0.0, nthreads=0,
          modes=None):
    """Backward sweep from stage N.  Returns dict(J_last, idx_last, stage[, J_all, idx_all])."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    S = int(np.prod(d.n))
    n_stages = d.N - 1 if n_stages is None else int(n_stages)
    J_last = np.zeros((d.P, S))
    idx_last = np.zeros((d.P, S), dtype=np.int32)
    J_all = np.zeros((d.N, d.P, S)) if keep_all else None
    idx_all = np.zeros((d.N, d.P, S), dtype=np.int32) if keep_all else None
    JN = None if J_N is None else _arr(J_N).reshape(d.P, S)
    fn = lib().oracle_sweep
    fn.restype = C.c_int
    stage_no = fn(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)),
                  JN.ctypes.data_as(_dp) if JN is not None else _dp(), C.c_int(n_stages),
                  J_all.ctypes.data_as(_dp) if keep_all else _dp(),
                  idx_all.ctypes.data_as(C.POINTER(C.c_int32)) if keep_all else C.POINTER(C.c_int32)(),
                  J_last.ctypes.data_as(_dp), idx_last.ctypes.data_as(C.POINTER(C.c_int32)),
                  C.c_int(check_period), C.c_double(check_tol), C.c_int(nthreads))
    out = {"J_last": J_last, "idx_last": idx_last, "stage": stage_no}
    if keep_all:
        out["J_all"], out["idx_all"] = J_all, idx_all
    return out


def rollout(d, idx_all, A, B, u_values, x0, mode=0, ssu_stage=1, modes=None):
    """idx_all [N, S]; x0 [batch, 2].  Returns X [batch, N, 2], U [batch, N]."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    x0 = _arr(x0).reshape(-1, 2)
    batch = x0.shape[0]
    X = np.zeros((batch, d.N, 2))
    U = np.zeros((batch, d.N))
    ia = np.ascontiguousarray(idx_all, dtype=np.int32)
    A = _arr(np.asarray(A).ravel(order="F"))
    B = _arr(np.asarray(B).ravel())
    uv = _arr(u_values)
    rc = lib().oracle_rollout(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)),
                              ia.ctypes.data_as(C.POINTER(C.c_int32)), A.ctypes.data_as(_dp),
                              B.ctypes.data_as(_dp), uv.ctypes.data_as(_dp), x0.ctypes.data_as(_dp),
                              C.c_int(batch), C.c_int(mode), C.c_int(ssu_stage),
                              X.ctypes.data_as(_dp), U.ctypes.data_as(_dp))
    assert rc == 0
    return X, U


def num_threads():
    return int(lib().oracle_num_threads())


def set_threads(n):
    lib().oracle_set_threads(C.c_int(int(n)))


def stage_points(d, J_next_p, states, p=0, modes=None):
    """Evaluate only the listed linear state indices of problem ``p``.  Returns (J, idx0).
    J_next_p = None: J_{k+1} is the first stage from a zero terminal cost, evaluated on the fly (a
    spot check of the SECOND stage at grid sizes whose arrays do not fit the host)."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    Jn = None if J_next_p is None else _arr(J_next_p).ravel()
    st = np.ascontiguousarray(states, dtype=np.int64)
    Jo = np.zeros(len(st))
    Io = np.zeros(len(st), dtype=np.int32)
    rc = lib().oracle_stage_points(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(p),
                                   Jn.ctypes.data_as(_dp) if Jn is not None else _dp(),
                                   st.ctypes.data_as(C.POINTER(C.c_int64)),
                                   C.c_int64(len(st)), Jo.ctypes.data_as(_dp),
                                   Io.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return Jo, Io


def policy_lookup(d, idx_p, x, p=0, modes=None):
    """Nearest-policy lookup: idx_p [S] policy of problem p, x [batch, D].  Returns idx0 [batch]."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    x = _arr(x).reshape(-1, len(d.n))
    ia = np.ascontiguousarray(idx_p, dtype=np.int32).ravel()
    out = np.zeros(len(x), dtype=np.int32)
    rc = lib().oracle_policy_lookup(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(p),
                                    ia.ctypes.data_as(C.POINTER(C.c_int32)), x.ctypes.data_as(_dp),
                                    C.c_int(len(x)), out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return out


def rollout_axis(d, idx, u_inc, x0, n_steps, h, rate_dim, p=0, time_varying=False, modes=None):
    """Simplified-plant rollout.  idx: [S] or, time varying, [n_steps(+), S] (row k-1 = stage k).
    x0 [batch, 2].  Returns X [batch, n_steps+1, 2], Cidx [batch, n_steps]."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    x0 = _arr(x0).reshape(-1, 2)
    batch = len(x0)
    ia = np.ascontiguousarray(idx, dtype=np.int32)
    stride = ia.shape[-1] if time_varying else 0
    X = np.zeros((batch, n_steps + 1, 2))
    Cc = np.zeros((batch, n_steps), dtype=np.int32)
    ui = _arr(u_inc)
    rc = lib().oracle_rollout_axis(C.byref(cd), modes.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(p),
                                   ia.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(stride),
                                   C.c_int(int(time_varying)), C.c_int(rate_dim), C.c_double(h),
                                   ui.ctypes.data_as(_dp), x0.ctypes.data_as(_dp), C.c_int(batch),
                                   C.c_int(n_steps), X.ctypes.data_as(_dp),
                                   Cc.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return X, Cc


def target_R0V0(mu=398600.0):
    """Solver_position.get_target_R0V0 (position-control/Solver_position.m:313-331)."""
    R0, V0 = (C.c_double * 3)(), (C.c_double * 3)()
    lib().oracle_get_target_R0V0(C.c_double(mu), R0, V0)
    return np.array(R0[:]), np.array(V0[:])


def kepler_U(mu, dt, ro, vro, a):
    fn = lib().oracle_kepler_U
    fn.restype = C.c_double
    n = C.c_int()
    x = fn(C.c_double(mu), C.c_double(dt), C.c_double(ro), C.c_double(vro), C.c_double(a), C.byref(n))
    return x, n.value


def update_RV_target(mu, R0, V0, t):
    R, V = (C.c_double * 3)(), (C.c_double * 3)()
    lib().oracle_update_RV_target(C.c_double(mu), (C.c_double * 3)(*R0), (C.c_double * 3)(*V0), C.c_double(t), R, V)
    return np.array(R[:]), np.array(V[:])


def sv_from_coe(coe, mu):
    r, v = (C.c_double * 3)(), (C.c_double * 3)()
    lib().oracle_sv_from_coe((C.c_double * 6)(*coe), C.c_double(mu), r, v)
    return np.array(r[:]), np.array(v[:])


def rollout_orbit(d, idx, u_values, y0, n_steps, h, R0, V0, mu=398600.0, tol=1e-8, modes=None):
    """Solver_position.get_optimal_path's stage loop (Solver_position.m:206-224) for a batch.
    idx [3, S] 0-based policies of the x, y, z axes; y0 [batch, 6].
    Returns X [batch, n_steps+1, 6], control indices [batch, n_steps, 3], rkf45 warnings [batch]."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    y0 = _arr(y0).reshape(-1, 6)
    batch = len(y0)
    ia = np.ascontiguousarray(idx, dtype=np.int32)
    X = np.zeros((batch, n_steps + 1, 6))
    Cc = np.zeros((batch, n_steps, 3), dtype=np.int32)
    W = np.zeros(batch, dtype=np.int32)
    uv = _arr(u_values)
    ip = C.POINTER(C.c_int32)
    rc = lib().oracle_rollout_orbit(C.byref(cd), modes.ctypes.data_as(ip), ia.ctypes.data_as(ip), uv.ctypes.data_as(_dp),
                                    C.c_double(mu), (C.c_double * 3)(*R0), (C.c_double * 3)(*V0), C.c_double(h),
                                    C.c_int(n_steps), C.c_double(tol), y0.ctypes.data_as(_dp), C.c_int(batch),
                                    X.ctypes.data_as(_dp), Cc.ctypes.data_as(ip), W.ctypes.data_as(ip))
    assert rc == 0
    return X, Cc, W


def ode45_linear(lam, t0, tf, y0, rtol=1e-3, atol=1e-6):
    """ode45 restatement on y' = lam .* y (test hook).  Returns (y(tf), accepted steps, failed attempts, warned)."""
    lam, y = _arr(lam), _arr(y0).copy()
    nf, w = C.c_int(), C.c_int()
    fn = lib().oracle_ode45_linear
    fn.restype = C.c_int
    n = fn(C.c_int(len(lam)), lam.ctypes.data_as(_dp), C.c_double(t0), C.c_double(tf), y.ctypes.data_as(_dp),
           C.c_double(rtol), C.c_double(atol), C.byref(nf), C.byref(w))
    return y, n, nf.value, w.value


def rollout_pos_att(descs, idxs, fvals, y0, n_steps, h, R0, V0, InertiaM, Mass, T_dist, mu=398600.0,
                    rtol=1e-3, atol=1e-6):
    """Solver_pos_att.get_optimal_path's stage loop (Solver_pos_att.m:452-500) for a batch.
    descs / idxs / fvals: the x, y, z channel descriptors, [S_ch] 0-based policies and [4, C_ch] thruster
    levels (f0/f1/f6/f7_allcomb).  y0 [batch, 13].  Returns X [batch, n_steps+1, 13], F [batch, n_steps, 12],
    FM [batch, n_steps, 6], warnings [batch]."""
    ip = C.POINTER(C.c_int32)
    cds, keeps, modes, ias, fvs = [], [], [], [], []
    for d, ix, fv in zip(descs, idxs, fvals):
        cd, keep = to_cdesc(d)
        cds.append(cd); keeps.append(keep)
        modes.append(locate_modes(d))
        ias.append(np.ascontiguousarray(ix, dtype=np.int32).ravel())
        fvs.append(_arr(fv).reshape(4, int(d.C)))
    y0 = _arr(y0).reshape(-1, 13)
    batch = len(y0)
    par = _arr(np.concatenate([[mu, h, rtol, atol, Mass, T_dist], np.ravel(R0), np.ravel(V0),
                               np.asarray(InertiaM, dtype=np.float64).ravel(order="F")]))
    X = np.zeros((batch, n_steps + 1, 13))
    F = np.zeros((batch, n_steps, 12))
    FM = np.zeros((batch, n_steps, 6))
    W = np.zeros(batch, dtype=np.int32)
    rc = lib().oracle_rollout_pos_att(C.byref(cds[0]), C.byref(cds[1]), C.byref(cds[2]),
                                      modes[0].ctypes.data_as(ip), modes[1].ctypes.data_as(ip), modes[2].ctypes.data_as(ip),
                                      ias[0].ctypes.data_as(ip), ias[1].ctypes.data_as(ip), ias[2].ctypes.data_as(ip),
                                      fvs[0].ctypes.data_as(_dp), fvs[1].ctypes.data_as(_dp), fvs[2].ctypes.data_as(_dp),
                                      par.ctypes.data_as(_dp), C.c_int(n_steps), y0.ctypes.data_as(_dp), C.c_int(batch),
                                      X.ctypes.data_as(_dp), F.ctypes.data_as(_dp), FM.ctypes.data_as(_dp), W.ctypes.data_as(ip))
    assert rc == 0
    return X, F, FM, W


def rollout_attitude(d, idx, u_values, y0, n_steps, h, InertiaM, rtol=1e-3, atol=1e-6, modes=None):
    """Solver_attitude.get_optimal_path_simplified_testode45 (Solver_attitude.m:1669-1705) for a batch.
    idx [3, S]; y0 [batch, 7] = (w1 w2 w3 q1 q2 q3 q4).  Returns X [batch, n_steps+1, 7], control indices
    [batch, n_steps, 3], warnings [batch]."""
    cd, keep = to_cdesc(d)
    modes = locate_modes(d) if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    y0 = _arr(y0).reshape(-1, 7)
    batch = len(y0)
    ia = np.ascontiguousarray(idx, dtype=np.int32)
    uv = _arr(u_values)
    par = _arr(np.concatenate([[h, rtol, atol], np.asarray(InertiaM, dtype=np.float64).ravel(order="F")]))
    X = np.zeros((batch, n_steps + 1, 7))
    Cc = np.zeros((batch, n_steps, 3), dtype=np.int32)
    W = np.zeros(batch, dtype=np.int32)
    rc = lib().oracle_rollout_attitude(C.byref(cd), modes.ctypes.data_as(ip), ia.ctypes.data_as(ip), uv.ctypes.data_as(_dp),
                                       par.ctypes.data_as(_dp), C.c_int(n_steps), y0.ctypes.data_as(_dp), C.c_int(batch),
                                       X.ctypes.data_as(_dp), Cc.ctypes.data_as(ip), W.ctypes.data_as(ip))
    assert rc == 0
    return X, Cc, W


def dense6_run(T, n_stages, J_N=None):
    """The coupled 6-D attitude sweep (Solver_attitude.run) on the tables of tables.attitude6_tables.
    Returns (J [S], idx [S]) of the last stage computed; idx = (u1*nu + u2)*nu + u3, 0-based."""
    P6, P3 = _dp * 6, _dp * 3
    keep = [_arr(g) for g in T.grid] + [_arr(w) for w in T.w_next] + [_arr(a) for a in T.a_next] + [_arr(x) for x in T.r]
    grid = P6(*[a.ctypes.data_as(_dp) for a in keep[0:6]])
    wn = P3(*[a.ctypes.data_as(_dp) for a in keep[6:9]])
    an = P3(*[a.ctypes.data_as(_dp) for a in keep[9:12]])
    rr = P3(*[a.ctypes.data_as(_dp) for a in keep[12:15]])
    gs = _arr(T.gs)
    n = np.asarray(T.n, dtype=np.int32)
    J = np.zeros(T.S)
    idx = np.zeros(T.S, dtype=np.int32)
    JN = None if J_N is None else _arr(J_N).ravel()
    rc = lib().oracle_dense6_run(n.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(T.nu), grid, wn, an, gs.ctypes.data_as(_dp), rr,
                                 C.c_int(int(n_stages)), JN.ctypes.data_as(_dp) if JN is not None else _dp(),
                                 J.ctypes.data_as(_dp), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return J, idx


def rollout_attitude6(T, idx, Jd, h, n_steps, x0):
    """Solver_attitude.get_optimal_path (:1487-1530) under the 6-D policy idx [S] (c = (u1*nu + u2)*nu + u3).
    x0 [batch, 7].  Returns X [batch, n_steps+1, 7], U [batch, n_steps, 3]."""
    P6 = _dp * 6
    keep = [_arr(g) for g in T.grid]
    grid = P6(*[a.ctypes.data_as(_dp) for a in keep])
    n = np.asarray(T.n, dtype=np.int32)
    ia = np.ascontiguousarray(idx, dtype=np.int32).ravel()
    uv, jd = _arr(T.U_vector), _arr(Jd)
    x0 = _arr(x0).reshape(-1, 7)
    batch = len(x0)
    X = np.zeros((batch, n_steps + 1, 7))
    U = np.zeros((batch, n_steps, 3))
    ip = C.POINTER(C.c_int32)
    rc = lib().oracle_rollout_attitude6(n.ctypes.data_as(ip), C.c_int(T.nu), grid, ia.ctypes.data_as(ip), uv.ctypes.data_as(_dp),
                                        jd.ctypes.data_as(_dp), C.c_double(h), C.c_int(n_steps), x0.ctypes.data_as(_dp),
                                        C.c_int(batch), X.ctypes.data_as(_dp), U.ctypes.data_as(_dp))
    assert rc == 0
    return X, U
