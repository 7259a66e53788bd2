"""Host-side mirror of the reference's classdef surface for the sweep path.

Same class names, property names, method names and argument meaning as the MATLAB classes; the
stage loop itself runs in libbellman.so (CUDA, sm_100a).  The MATLAB originals of these facades
live in ``matlab/`` and call the same C ABI through ``bellman_mex``.

  Dynamic_Solver   test/Dynamic_Solver.m        run(), get_optimal_path(X0, mode, ssu_num)
  Solver_position  position-control/Solver_position.m   simplified_run()
  Solver_attitude  attitude-control/Solver_attitude.m   simplified_run()
  Solver_pos_att   pos-att/Solver_pos_att.m     simplified_run(), calculate_one_channel_U_Opt()

Deviations from the shipped reference (all documented in SURVEY.md 2.3):
  * everything is fp64 (the test/test_coder.m / test/obj_1.mat convention); the ``single`` casts
    of Dynamic_Solver.m:69 and Solver_pos_att.m:265,800 are not emulated;
  * ``J_star`` is filled for every stage (the current Dynamic_Solver.m allocates but never
    writes it); ``get_optimal_path`` honours a user-supplied X0 (README.md:22);
  * ``idsum50_prev`` starts at 0 (Solver_pos_att.m:277 reads it undefined).
"""
import numpy as np

from . import tables
from ._lib import SweepGroup  # noqa: E402
from ._lib import KERNEL_AUTO, Sweep


def _unflatten(a, shape):
    """[S] column-major -> ndarray of ``shape`` (MATLAB layout)."""
    return np.asarray(a).reshape(shape, order="F")


class NearestPolicy:
    """Stand-in for ``griddedInterpolant({s1,..}, values, 'nearest')`` (Solver_position.m:144-146,
    Solver_attitude.m:249-251, Solver_pos_att.m:851-861): nearest node, clamped outside."""

    def __init__(self, gridvecs, values):
        self.GridVectors = [np.asarray(g, dtype=np.float64) for g in gridvecs]
        self.Values = np.asarray(values)

    def __call__(self, *q):
        idx = []
        for s, x in zip(self.GridVectors, q):
            x = np.asarray(x, dtype=np.float64)
            i = np.clip(np.searchsorted(s, x, side="right") - 1, 0, len(s) - 2)
            idx.append(i + ((x - s[i]) >= (s[i + 1] - x)))
        return self.Values[tuple(idx)]


# ----------------------------------------------------------------------------------------------
class Dynamic_Solver:
    """test/Dynamic_Solver.m — Kirk ch.3 two-state linear regulator by dynamic programming."""

    def __init__(self):
        # constructor defaults, Dynamic_Solver.m:47-64
        self.checkstagesXJF = 0          # debug slices hard-code indices 50:57/105
        self.Q = np.array([[0.25, 0.0], [0.0, 0.05]])
        self.A = np.array([[0.9974, 0.0539], [-0.1078, 1.1591]])
        self.B = np.array([[0.0013], [0.0539]])
        self.R = 0.05
        self.H = np.zeros((0, 0))
        self.N = 200
        self.S = 2
        self.C = 1
        self.dx = 100
        self.du = 1000
        self.x_max = 3.0
        self.x_min = -2.5
        self.u_max = 10.0
        self.u_min = -40.0
        self.s_r = None
        self.u_star = None
        self.u_star_idx = None
        self.J_star = None
        self.X1_mesh = None
        self.X2_mesh = None
        # build options (not in the reference)
        self.store_J_star = True         # keep every stage's J / u_star on the host as the reference does
        self.device = -1
        self.n_gpus = 1                  # > 1: slabs over this many GPUs driven from this one process
        self.devices = None              # GPU of every slab (default 0 .. n_gpus-1)
        self.kernel = KERNEL_AUTO
        self.verbose = False
        self._sweep = None
        self._desc = None

    def _build(self):
        d = tables.kirk_desc(self.A, self.B, self.Q, self.R, self.N, self.x_min, self.x_max, self.dx,
                             self.u_min, self.u_max, self.du,
                             store_J_all=self.store_J_star, store_idx_all=True)
        self._desc = d
        self.s_r = d.meta["s_r"]
        self.U_mesh = d.meta["U_mesh"]
        self.X1_mesh, self.X2_mesh = np.meshgrid(self.s_r, self.s_r, indexing="ij")
        return d

    def run(self):
        """Dynamic_Solver.m:66-105: backward sweep k = 1..N-1, u_star(:,:,N-k) = U_mesh(idx)."""
        d = self._build()
        if self._sweep is not None:
            self._sweep.close()
        # u_star of every stage stays on the device: one or two bytes per state are enough for du controls
        ib = 1 if d.C <= 256 else 2 if d.C <= 65536 else 4
        if self.n_gpus > 1:
            # one host thread, n slabs (bellman_group_run); every kept stage is then gathered into an
            # unsharded handle so that get_optimal_path and the per-stage reads below work as after a one-GPU run
            grp = SweepGroup(d, self.devices or list(range(self.n_gpus)), idx_bytes=ib)
            grp.run(d.N - 1, kernel=self.kernel)
            sw = self._sweep = Sweep(d, device=self.device, idx_bytes=ib)
            for k in range(d.N - 1, 0, -1):          # u_star of every stage is always kept (store_idx_all)
                sw.set_stage(k, grp.get_J(k) if (d.store_J_all or k == 1) else None, grp.get_idx(k))
            self.sweep_stats = grp.stats()
            grp.close()
        else:
            sw = self._sweep = Sweep(d, device=self.device, idx_bytes=ib)
            sw.run(d.N - 1, kernel=self.kernel, sync_each_stage=self.verbose)
        if self.verbose:
            st = sw.stats()
            print("sweep: %d stages in %.3f ms (%s kernel)" % (d.N - 1, st["ms"], sw.last_kernel))
        shape = (d.n[0], d.n[1])
        N = d.N
        if self.store_J_star:
            self.J_star = np.zeros(shape + (N,))
            self.u_star = np.zeros(shape + (N,))
            for k in range(1, N + 1):
                self.J_star[:, :, k - 1] = _unflatten(sw.get_J(k)[0], shape)
            for k in range(1, N):
                self.u_star[:, :, k - 1] = self.U_mesh[_unflatten(sw.get_idx(k)[0], shape)]
        self.u_star_idx = _unflatten(sw.get_idx(1)[0], shape) + 1      # 1-based like MATLAB
        self.F_Values = _unflatten(sw.get_J(1)[0], shape)               # obj.F.Values after the loop
        return self

    def get_optimal_path(self, X0=None, mode="Nssu", ssu_num=1):
        """Dynamic_Solver.m:108-145: forward rollout with linear interpolation of u_star values.
        X0 may be one state [2] or a batch [batch, 2].  Returns (X [.., 2, N], U [.., N])."""
        if self._sweep is None or self._sweep.current_stage != 1:
            raise RuntimeError("run(obj) must complete before get_optimal_path")
        if X0 is None:
            X0 = [2.0, 1.0]                                             # :110
        x0 = np.asarray(X0, dtype=np.float64)
        single = x0.size == 2
        x0 = x0.reshape(-1, 2)
        X, U = self._sweep.rollout(self.A, self.B, self.U_mesh, x0, mode=1 if mode == "ssu" else 0,
                                   ssu_stage=int(ssu_num))
        X = np.swapaxes(X, 1, 2)                                        # [batch, 2, N]
        return (X[0], U[0]) if single else (X, U)

    # -- per-stage archive (the reference keeps runs as saved objects, e.g. test/obj_1.mat) -------
    _saved = ("A", "B", "Q", "R", "N", "S", "C", "dx", "du", "x_min", "x_max", "u_min", "u_max", "s_r",
              "U_mesh", "X1_mesh", "X2_mesh", "J_star", "u_star", "u_star_idx")

    def save(self, file_name, name="obj"):
        """Archive the run (every stage's J_star / u_star and the settings) as a struct ``name`` in a
        .mat file: the data compare_data (:266-280) works on.  MATLAB loads it as a plain struct."""
        import scipy.io
        scipy.io.savemat(file_name, {name: {k: getattr(self, k) for k in self._saved
                                            if getattr(self, k, None) is not None}}, do_compression=True)

    @classmethod
    def load(cls, file_name, name="obj"):
        """Inverse of save(): an object holding the archived properties (not runnable until run())."""
        import scipy.io
        st = scipy.io.loadmat(file_name, squeeze_me=True, struct_as_record=False)[name]
        obj = cls()
        for k in cls._saved:
            if hasattr(st, k):
                v = getattr(st, k)
                setattr(obj, k, int(v) if k in ("N", "S", "C", "dx", "du") else
                        (float(v) if np.ndim(v) == 0 else np.asarray(v)))
        return obj

    @staticmethod
    def compare_data(obj1, obj2):
        """Dynamic_Solver.m:266-280: bit-exact comparison of two saved runs."""
        if obj1.J_star is None or obj2.J_star is None or obj1.J_star.size == 0 or obj2.J_star.size == 0:
            raise ValueError("stop throwing empty data at me")
        return bool(np.array_equal(obj1.J_star, obj2.J_star))


# ----------------------------------------------------------------------------------------------
class _AxisSolverBase:
    """Shared by Solver_position / Solver_attitude: three 2-D axes swept as one batched problem."""

    device = -1
    kernel = KERNEL_AUTO
    use_graph = True
    n_gpus = 1          # > 1: the grid is cut into slabs over this many GPUs, driven from this one process
    devices = None      # GPU of every slab (default 0 .. n_gpus-1)

    def _axis_descs(self):
        raise NotImplementedError

    def simplified_run(self, n_stages=None):
        descs = self._axis_descs()
        d = tables.stack_problems(descs)
        self._desc = d
        todo = d.N - 1 if n_stages is None else int(n_stages)
        if self.n_gpus > 1:
            # one host thread, n slabs (bellman_group_run); the kept policy is then gathered into an
            # unsharded handle so that lookups and rollouts work exactly as after a one-GPU sweep
            grp = SweepGroup(d, self.devices or list(range(self.n_gpus)))
            grp.run(todo, kernel=self.kernel)
            J, idx, stage, stats = grp.get_J(), grp.get_idx(), grp.current_stage, grp.stats()
            grp.close()
            sw = self._sweep = Sweep(d, device=self.device)
            sw.set_stage(stage, J, idx)
        else:
            sw = self._sweep = Sweep(d, device=self.device)
            sw.run(todo, kernel=self.kernel, use_graph=self.use_graph)
            J = sw.get_J()
            idx = sw.get_idx()
            stats = sw.stats()
        shape = tuple(d.n)
        self.F_Values = [_unflatten(J[p], shape) for p in range(d.P)]
        self.U_idx = [_unflatten(idx[p], shape) + 1 for p in range(d.P)]
        pol = []
        for p in range(d.P):
            grids = [d.grid[k][p] for k in range(d.D)]
            pol.append(NearestPolicy(grids, np.asarray(self.U_vector)[self.U_idx[p] - 1]))
        self.U1_Opt, self.U2_Opt, self.U3_Opt = pol
        self.sweep_stats = stats
        return self

    def get_optimal_path_simplified(self, X0, n_steps=None):
        """Closed loop on the SIMPLIFIED plant under the kept nearest policy — the per-axis loop of
        attitude-control/test/test_simplified.m:129-151 (policy held at its final sweep stage, as
        U*_Opt is after simplified_run), run on the GPU for a batch of initial states.
        X0 is [batch, 3, 2] (or [3, 2]): per axis the two states in grid order.  Returns
        X [batch, 3, n_steps+1, 2] and U [batch, 3, n_steps] (control VALUES)."""
        sw, d = self._sweep, self._desc
        X0 = np.asarray(X0, dtype=np.float64).reshape(-1, d.P, 2)
        n_steps = d.N - 1 if n_steps is None else int(n_steps)
        X = np.empty((len(X0), d.P, n_steps + 1, 2))
        U = np.empty((len(X0), d.P, n_steps))
        for p in range(d.P):
            Xp, Cp = sw.rollout_axis(d.Tc[self._rate_dim][p], X0[:, p], n_steps, self.h, self._rate_dim, prob=p)
            X[:, p], U[:, p] = Xp, np.asarray(self.U_vector)[Cp]
        return X, U


class Solver_position(_AxisSolverBase):
    """position-control/Solver_position.m — three independent (x, v) axes, three thrust levels."""
    _rate_dim = 1

    def __init__(self):
        # Solver_position.m:46-92
        self.v_min, self.v_max = -0.5, 0.5
        self.x_min, self.x_max = -0.5, 0.5
        self.n_mesh_v = 200
        self.n_mesh_x = 200
        self.Mass = 4.16
        self.Qx1 = self.Qx2 = self.Qx3 = 6.0
        self.Qv1 = self.Qv2 = self.Qv3 = 6.0
        self.R1 = self.R2 = self.R3 = 0.1
        self.T_final = 30.0
        self.h = 0.005
        self.N_stage = int(np.ceil(self.T_final / self.h))             # :75-79 (always "ceil")
        self.defaultX0 = np.zeros(6)
        self.U_vector = np.array([-0.13, 0.0, 0.13]) * 2               # :84
        self.U1_Opt = self.U2_Opt = self.U3_Opt = None

    def _axis_descs(self):
        self.N_stage = int(np.ceil(self.T_final / self.h))
        qx = (self.Qx1, self.Qx2, self.Qx3)
        qv = (self.Qv1, self.Qv2, self.Qv3)
        r = (self.R1, self.R2, self.R3)
        ds = [tables.position_axis_desc(self.x_min, self.x_max, self.n_mesh_x, self.v_min, self.v_max,
                                        self.n_mesh_v, self.U_vector, self.Mass, qx[a], qv[a], r[a],
                                        self.h, self.N_stage) for a in range(3)]
        self.n_mesh_x, self.n_mesh_v = ds[0].n                         # :100,:104
        return ds


    # --- orbital forward simulation (Solver_position.m:189-361) ------------------------------------
    mu = 398600.0                                                      # :192

    @staticmethod
    def sv_from_coe(coe, mu):
        """position-control/private/sv_from_coe.m:31-66 — state vector from [h e RA incl w TA]."""
        h, e, RA, incl, w, TA = (float(x) for x in coe)
        rp = (h ** 2 / mu) * (1 / (1 + e * np.cos(TA))) * (np.cos(TA) * np.array([1.0, 0, 0]) + np.sin(TA) * np.array([0, 1.0, 0]))
        vp = (mu / h) * (-np.sin(TA) * np.array([1.0, 0, 0]) + (e + np.cos(TA)) * np.array([0, 1.0, 0]))
        R3_W = np.array([[np.cos(RA), np.sin(RA), 0], [-np.sin(RA), np.cos(RA), 0], [0, 0, 1.0]])
        R1_i = np.array([[1.0, 0, 0], [0, np.cos(incl), np.sin(incl)], [0, -np.sin(incl), np.cos(incl)]])
        R3_w = np.array([[np.cos(w), np.sin(w), 0], [-np.sin(w), np.cos(w), 0], [0, 0, 1.0]])
        Q_pX = (R3_w @ R1_i @ R3_W).T
        return Q_pX @ rp, Q_pX @ vp

    def get_target_R0V0(self):
        """:313-331 — target A: perigee altitude 300 km, e = 0.1, equatorial, at perigee."""
        RE, e = 6378.0, 0.1
        rp = RE + 300
        ra = rp * (1 + e) / (1 - e)
        h_ = np.sqrt(2 * self.mu * rp * ra / (ra + rp))
        return self.sv_from_coe([h_, e, 0.0, 0.0, 0.0, 0.0], self.mu)

    def get_optimal_path(self, y0=None, n_steps=None, stride_out=1, tol=1e-8):
        """:189-224 on the GPU for a batch of initial relative states y0 [batch, 6] (default the
        reference's single [-1 0 0 0 0 0]): nearest policy per axis, one rkf45 call per stage.
        Returns X_ode45 [batch, n_steps/stride_out + 1, 6] and F_Opt_history [batch, n_steps/stride_out, 3]."""
        sw = self._sweep
        y0 = np.array([[-1.0, 0, 0, 0, 0, 0]]) if y0 is None else np.asarray(y0, dtype=np.float64).reshape(-1, 6)
        N = int(np.ceil(self.T_final / self.h))                        # :206
        n_steps = N - 1 if n_steps is None else int(n_steps)
        R0, V0 = self.get_target_R0V0()
        X, Cc, W = sw.rollout_orbit(self.U_vector, y0, n_steps, self.h, R0, V0, mu=self.mu, tol=tol, stride_out=stride_out)
        self.rkf45_warnings = W
        return X, np.asarray(self.U_vector)[Cc]


class Solver_attitude(_AxisSolverBase):
    """attitude-control/Solver_attitude.m — simplified_run: three (w, theta) axes, three torques."""
    _rate_dim = 0

    def __init__(self):
        # Solver_attitude.m:103-193
        self.w_min = -tables.deg2rad(50.0)
        self.w_max = -tables.deg2rad(-50.0)
        self.n_mesh_w = 1000
        self.yaw_min, self.yaw_max = -30.0, 30.0
        self.pitch_min, self.pitch_max = -20.0, 20.0
        self.roll_min, self.roll_max = -35.0, 35.0
        self.n_mesh_q = 10
        self.n_mesh_t = 300
        i1, i2, i3 = 0.02836 + 0.00016, 0.026817 + 0.00150, 0.023 + 0.00150
        i4, i5, i6 = -0.0000837, 0.000014, -0.00029
        self.InertiaM = np.array([[i1, i4, i5], [i4, i2, i6], [i5, i6, i3]])
        self.Q1 = self.Q2 = self.Q3 = 6.0
        self.Q4 = self.Q5 = self.Q6 = 6.0
        self.R1 = self.R2 = self.R3 = 4.0
        self.Qt1, self.Qt2, self.Qt3 = self.Q4, self.Q5, self.Q6
        self.T_final = 30.0
        self.h = 0.005
        self.N_stage = int(np.ceil(self.T_final / self.h))
        self.J1, self.J2, self.J3 = (self.InertiaM.ravel(order="F")[k] for k in (0, 4, 8))
        self.U_vector = np.array([-0.11, 0.0, 0.11])
        self.U1_Opt = self.U2_Opt = self.U3_Opt = None

    # Solver_attitude.m:306-309: w0 = 0, q0 = angle2quat(5, 10, -9 deg) reversed (scalar last)
    defaultX0_ode45 = np.array([0.0, 0.0, 0.0, 0.0501511024391496, 0.0833950587800888, -0.0818761044636256, 0.991880252153991])

    def get_optimal_path_simplified_testode45(self, X0=None, n_steps=None, stride_out=1):
        """:1669-1705 on the GPU for a batch X0 [batch, 7] = (w1 w2 w3 q1 q2 q3 q4): per stage
        U(k) = U{k}_Opt(X(k), 2*asin(X(3+k))), then ode45 over one stage on the full rigid-body plant
        (:1803-1849).  Returns X_ode45 [batch, n_out + 1, 7] and the applied torques [batch, n_out, 3]."""
        sw = self._sweep
        X0 = self.defaultX0_ode45[None] if X0 is None else np.asarray(X0, dtype=np.float64).reshape(-1, 7)
        n_steps = int(np.ceil(self.T_final / self.h)) - 1 if n_steps is None else int(n_steps)
        X, Cc, W = sw.rollout_attitude(self.U_vector, X0, n_steps, self.h, self.InertiaM, stride_out=stride_out)
        self.ode45_warnings = W
        return X, np.asarray(self.U_vector)[Cc]

    # --- the coupled 6-D problem (Solver_attitude.m:521-601) --------------------------------------
    def dense6_tables(self):
        """reshape_states + calculate_J_current_state_fix_shaped + spacecraft_dynamics_taylor_estimate
        (:1433-1485, :629-685, :825-925) on the mesh n_mesh_w^3 x n_mesh_q^3."""
        self.N_stage = int(np.ceil(self.T_final / self.h))
        sr = tables.linspace(self.w_min, self.w_max, self.n_mesh_w)
        lim = ((self.yaw_min, self.yaw_max), (self.pitch_min, self.pitch_max), (self.roll_min, self.roll_max))
        ang = [tables.linspace(tables.deg2rad(a), tables.deg2rad(b), self.n_mesh_q) for a, b in lim]
        return tables.attitude6_tables((sr, sr, sr), ang[0], ang[1], ang[2], self.U_vector, self.J1, self.J2, self.J3,
                                       (self.Q1, self.Q2, self.Q3, self.Q4, self.Q5, self.Q6), (self.R1, self.R2, self.R3),
                                       self.h, self.N_stage)

    def run(self, n_stages=None, max_bytes=64e9):
        """Solver_attitude.run (:521-601): the coupled sweep over (w1 w2 w3 yaw pitch roll) with 27 control
        combinations, on the GPU (bellman_dense6_run).  The reference never ran it (:282 calls a
        one-argument method with two; its default n_mesh_w = 1000 needs 2.7e13-element arrays), so choose
        n_mesh_w / n_mesh_q that fit: the tables take 9 doubles per state on the host.
        Sets F_Values (J of the last stage computed, grid shaped), U1_Opt / U2_Opt / U3_Opt (the torque
        VALUES chosen at every state, as :561-563 leave them) and U_idx6 (the three 1-based level indices)."""
        from ._lib import dense6_run
        S = float(self.n_mesh_w) ** 3 * float(self.n_mesh_q) ** 3
        if S * 9 * 8 > max_bytes:
            raise ValueError("the 6-D mesh %d^3 x %d^3 needs %.3g bytes of tables; lower n_mesh_w / n_mesh_q "
                             "(the reference's default n_mesh_w = 1000 is infeasible, SURVEY 2.3)" % (self.n_mesh_w, self.n_mesh_q, S * 72))
        T = self.dense6_tables()
        todo = T.N - 1 if n_stages is None else int(n_stages)
        J, idx, ms = dense6_run(T, todo, device=self.device)
        shape = tuple(T.n)
        nu = T.nu
        u1, u2, u3 = idx // (nu * nu), (idx // nu) % nu, idx % nu
        self.F_Values = _unflatten(J, shape)
        self.U_idx6 = [_unflatten(u, shape) + 1 for u in (u1, u2, u3)]
        U = np.asarray(self.U_vector)
        self.U1_Opt, self.U2_Opt, self.U3_Opt = (U[k - 1] for k in self.U_idx6)
        self.sweep_stats = {"ms": ms, "stages": todo, "kernel": "dense6"}
        self._dense6 = (T, idx)
        return self

    def get_optimal_path(self, X0=None, n_steps=None):
        """:1487-1530 on the GPU for a batch X0 [batch, 7] = (w1 w2 w3 q1 q2 q3 q4) (default obj.defaultX0):
        the 6-D 'nearest' policy of run(), first-order ('taylor') plant step.  Returns X [batch, n_steps+1, 7]
        and the applied torques U [batch, n_steps, 3]."""
        from ._lib import rollout_attitude6
        if getattr(self, "_dense6", None) is None:
            raise RuntimeError("run(obj) must complete before get_optimal_path")
        T, idx = self._dense6
        X0 = self.defaultX0_ode45[None] if X0 is None else np.asarray(X0, dtype=np.float64).reshape(-1, 7)
        n_steps = T.N - 1 if n_steps is None else int(n_steps)
        return rollout_attitude6(T, idx, (self.J1, self.J2, self.J3), self.h, n_steps, X0, device=self.device)

    def _axis_descs(self):
        self.N_stage = int(np.ceil(self.T_final / self.h))
        ang = ((self.yaw_min, self.yaw_max), (self.pitch_min, self.pitch_max), (self.roll_min, self.roll_max))
        qw = (self.Q1, self.Q2, self.Q3)
        qt = (self.Qt1, self.Qt2, self.Qt3)
        r = (self.R1, self.R2, self.R3)
        jj = (self.J1, self.J2, self.J3)
        return [tables.attitude_axis_desc(self.w_min, self.w_max, self.n_mesh_w, ang[a][0], ang[a][1],
                                          self.n_mesh_t, self.U_vector, jj[a], qw[a], qt[a], r[a],
                                          self.h, self.N_stage) for a in range(3)]


# ----------------------------------------------------------------------------------------------
class Solver_pos_att:
    """pos-att/Solver_pos_att.m — coupled position+attitude channels on a 4-D (x, v, theta, w)
    grid with 9 thruster on/off combinations (simplified_run / calculate_one_channel_U_Opt)."""

    def __init__(self):
        # Solver_pos_att.m:96-195
        self.v_min, self.v_max, self.n_mesh_v = -0.1, 0.1, 30
        self.x_min, self.x_max, self.n_mesh_x = -0.2, 0.2, 30
        self.w_min, self.w_max, self.n_mesh_w = tables.deg2rad(-2.0), tables.deg2rad(2.0), 15
        self.theta1_min, self.theta1_max = -5.0, 5.0
        self.theta2_min, self.theta2_max = -6.0, 6.0
        self.theta3_min, self.theta3_max = -7.0, 7.0
        self.n_mesh_t = 20
        self.Mass = 4.16
        i1, i2, i3 = 0.02836 + 0.00016, 0.026817 + 0.00150, 0.023 + 0.00150
        i4, i5, i6 = -0.0000837, 0.000014, -0.00029
        self.InertiaM = np.array([[i1, i4, i5], [i4, i2, i6], [i5, i6, i3]])
        self.J1, self.J2, self.J3 = (self.InertiaM.ravel(order="F")[k] for k in (0, 4, 8))
        self.Qx1 = self.Qx2 = self.Qx3 = 6.0
        self.Qv1 = self.Qv2 = self.Qv3 = 6.0
        self.Qt1 = self.Qt2 = self.Qt3 = 0.5
        self.Qw1 = self.Qw2 = self.Qw3 = 0.5
        self.R1 = self.R2 = self.R3 = 0.1
        self.T_final = 10.0
        self.h = 0.005
        self.N_stage = int(np.ceil(self.T_final / self.h))
        self.defaultX0 = np.zeros(9)
        T = 0.13
        self.T_dist = 9.65e-2
        on, off = np.array([0.0, T]), -np.array([0.0, T])
        self.F_Thr0 = on.copy(); self.F_Thr1 = on.copy(); self.F_Thr6 = off.copy(); self.F_Thr7 = off.copy()
        self.F_Thr2 = on.copy(); self.F_Thr3 = on.copy(); self.F_Thr8 = off.copy(); self.F_Thr9 = off.copy()
        self.F_Thr4 = on.copy(); self.F_Thr5 = on.copy(); self.F_Thr10 = off.copy(); self.F_Thr11 = off.copy()
        self.device = -1
        self.n_gpus = 1                 # > 1: every channel sweep is cut into slabs over this many GPUs, one process
        self.devices = None             # GPU of every slab (default 0 .. n_gpus-1)
        self.kernel = KERNEL_AUTO
        self.check_period = 50          # Solver_pos_att.m:273
        self.check_tol = 1e-2           # :269
        self.controllers = {}
        self._channel_ctl = {}
        self._lookup_sweeps = {}

    def _grids(self, ch):
        th = ((self.theta1_min, self.theta1_max), (self.theta2_min, self.theta2_max),
              (self.theta3_min, self.theta3_max))[ch]
        sl = tables.sym_linspace_pos_att
        return (sl(self.x_min, self.x_max, self.n_mesh_x), sl(self.v_min, self.v_max, self.n_mesh_v),
                sl(tables.deg2rad(th[0]), tables.deg2rad(th[1]), self.n_mesh_t),
                sl(self.w_min, self.w_max, self.n_mesh_w))

    def channel_desc(self, ch, failure=False):
        """Descriptor of channel 0/1/2 = x/y/z (Solver_pos_att.m:217-233); failure=True is the
        x-channel thruster-0-failed controller (:236-240)."""
        self.N_stage = int(np.ceil(self.T_final / self.h))
        thr = ((self.F_Thr0, self.F_Thr1, self.F_Thr6, self.F_Thr7),
               (self.F_Thr2, self.F_Thr3, self.F_Thr8, self.F_Thr9),
               (self.F_Thr4, self.F_Thr5, self.F_Thr10, self.F_Thr11))[ch]
        if failure:
            thr = (np.array([0.0]),) + thr[1:]
        Q = ((self.Qx1, self.Qv1, self.Qt1, self.Qw1, self.R1), (self.Qx2, self.Qv2, self.Qt2, self.Qw2, self.R2),
             (self.Qx3, self.Qv3, self.Qt3, self.Qw3, self.R3))[ch]
        J = (self.J2, self.J3, self.J1)[ch]
        s_x, s_v, s_t, s_w = self._grids(ch)
        return tables.pos_att_channel_desc(s_x, s_v, s_t, s_w, *thr, Q[0], Q[1], Q[2], Q[3], Q[4], J,
                                           self.Mass, self.T_dist, self.h, self.N_stage)

    def calculate_one_channel_U_Opt(self, ch, failure=False, file_name=None, n_stages=None):
        """Solver_pos_att.m:244-297 for one channel; returns the controller dict that the reference
        saves (F_gI values + grid vectors, U_Optimal_id (1-based), f*_allcomb)."""
        d = self.channel_desc(ch, failure)
        if self.n_gpus > 1:     # one host thread, n slabs along the dimension with the smallest halo (bellman_group_run)
            sw = SweepGroup(d, self.devices or list(range(self.n_gpus)))
        else:
            sw = Sweep(d, device=self.device)
        todo = d.N - 1 if n_stages is None else int(n_stages)
        sw.run(todo, kernel=self.kernel, check_period=self.check_period, check_tol=self.check_tol)
        shape = tuple(d.n)
        ctl = {
            "GridVectors": [d.grid[k][0] for k in range(4)],
            "F_gI_Values": _unflatten(sw.get_J()[0], shape),
            "U_Optimal_id": _unflatten(sw.get_idx()[0], shape) + 1,
            "f0_allcomb": d.meta["f0_allcomb"], "f1_allcomb": d.meta["f1_allcomb"],
            "f6_allcomb": d.meta["f6_allcomb"], "f7_allcomb": d.meta["f7_allcomb"],
            "stop_stage": sw.current_stage, "check_log": sw.check_log(), "stats": sw.stats(),
        }
        sw.close()
        if file_name:
            self.save_controller(file_name, ctl)
        return ctl

    # -- controller files (Solver_pos_att.m:291 save, :849-882 set_controller) -------------------
    @staticmethod
    def save_controller(file_name, ctl):
        """Write the variables the reference saves (:291).  F_gI is written as a struct with the two
        griddedInterpolant properties set_controller reads (GridVectors, Values), so MATLAB's
        ``C = load(file); C.F_gI.GridVectors`` works on it unchanged."""
        import scipy.io
        gv = np.empty((1, 4), dtype=object)
        for k in range(4):
            gv[0, k] = np.asarray(ctl["GridVectors"][k], dtype=np.float64).reshape(1, -1)
        scipy.io.savemat(file_name, {
            "F_gI": {"GridVectors": gv, "Values": ctl["F_gI_Values"]},
            "U_Optimal_id": np.asarray(ctl["U_Optimal_id"], dtype=np.float64),
            "f0_allcomb": ctl["f0_allcomb"], "f1_allcomb": ctl["f1_allcomb"],
            "f6_allcomb": ctl["f6_allcomb"], "f7_allcomb": ctl["f7_allcomb"],
            "stop_stage": float(ctl.get("stop_stage", 0)),
        }, do_compression=True)

    @staticmethod
    def load_controller(file_name):
        import scipy.io
        m = scipy.io.loadmat(file_name, squeeze_me=False, struct_as_record=False)
        F = m["F_gI"][0, 0]
        gv = [np.asarray(g, dtype=np.float64).ravel() for g in F.GridVectors.ravel()]
        ctl = {"GridVectors": gv, "F_gI_Values": np.asarray(F.Values, dtype=np.float64),
               "U_Optimal_id": np.asarray(m["U_Optimal_id"]).astype(np.int32)}
        for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"):
            ctl[k] = np.asarray(m[k], dtype=np.float64).ravel()
        if "stop_stage" in m:
            ctl["stop_stage"] = int(m["stop_stage"].ravel()[0])
        return ctl

    _channel_thrusters = {"x": (0, 1, 6, 7), "y": (2, 3, 8, 9), "z": (4, 5, 10, 11)}

    def set_controller(self, file, channel):
        """Solver_pos_att.m:849-882: ``file`` is a controller .mat (or the dict
        calculate_one_channel_U_Opt returns); installs Opt_F_Thr* 'nearest' policies of the channel."""
        if channel not in self._channel_thrusters:
            raise ValueError("wrong channel, must be one of x-y-z values")
        ctl = file if isinstance(file, dict) else self.load_controller(file)
        uid = np.asarray(ctl["U_Optimal_id"]).astype(np.int64) - 1
        for name, thr in zip(("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"), self._channel_thrusters[channel]):
            setattr(self, "Opt_F_Thr%d" % thr, NearestPolicy(ctl["GridVectors"], np.asarray(ctl[name]).ravel()[uid]))
        self._channel_ctl[channel] = ctl
        old = self._lookup_sweeps.pop(channel, None)
        if old is not None:
            old.close()
        return self

    def get_thruster_on_off_optimal(self, x, v, t, w, R0=None, V0=None, q=None):
        """Solver_pos_att.m:404-449 for one state: x, v relative position / velocity, t, w the body
        angles / rates (3-vectors).  With R0, V0, q given, x and v are first taken from the RSW frame
        to the body frame (:411-415, :825-847).  Returns the twelve thruster levels f0..f11."""
        x = np.asarray(x, dtype=np.float64).ravel()
        v = np.asarray(v, dtype=np.float64).ravel()
        if R0 is not None:
            M = self.ECI2body(q) @ self.RSW2ECI(R0, V0)
            x, v = M @ x, M @ v
        f = np.zeros(12)
        ang = {"x": 1, "y": 2, "z": 0}                                  # t_y/w_y, t_z/w_z, t_x/w_x
        for ci, ch in enumerate("xyz"):
            a = ang[ch]
            for thr in self._channel_thrusters[ch]:
                f[thr] = float(getattr(self, "Opt_F_Thr%d" % thr)(x[ci], v[ci], t[a], w[a]))
        return f

    def thruster_lookup_batch(self, channel, states):
        """The same 'nearest' lookup for a BATCH of channel states [batch, 4] = (x, v, theta, w) on the
        GPU (bellman_policy_lookup): returns [batch, 4] levels of the channel's four thrusters."""
        ctl = self._channel_ctl[channel]
        sw = self._channel_sweep(channel)
        c = sw.policy_lookup(states)
        return np.stack([np.asarray(ctl[k]).ravel()[c] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")], 1)

    @staticmethod
    def ECI2body(q):                                                     # :825-829
        q1, q2, q3, q4 = (float(a) for a in np.asarray(q).ravel())
        return np.array([
            [1 - 2 * (q2 ** 2 + q3 ** 2), 2 * (q1 * q2 + q3 * q4), 2 * (q1 * q3 - q2 * q4)],
            [2 * (q2 * q1 - q3 * q4), 1 - 2 * (q1 ** 2 + q3 ** 2), 2 * (q2 * q3 + q1 * q4)],
            [2 * (q3 * q1 + q2 * q4), 2 * (q3 * q2 - q1 * q4), 1 - 2 * (q1 ** 2 + q2 ** 2)]])

    @staticmethod
    def RSW2ECI(pos, vel):                                               # :831-847
        pos = np.asarray(pos, dtype=np.float64).ravel()
        vel = np.asarray(vel, dtype=np.float64).ravel()
        R = pos / np.linalg.norm(pos)
        h = np.cross(pos, vel)
        W = h / np.linalg.norm(h)
        S = np.cross(W, R)
        return np.column_stack([R, S, W])

    mu = 398600.0                                                        # :454

    def get_target_R0V0(self):
        """:759-777 — the same target orbit as Solver_position (perigee altitude 300 km, e = 0.1)."""
        RE, e = 6378.0, 0.1
        rp = RE + 300
        ra = rp * (1 + e) / (1 - e)
        h_ = np.sqrt(2 * self.mu * rp * ra / (ra + rp))
        return Solver_position.sv_from_coe([h_, e, 0.0, 0.0, 0.0, 0.0], self.mu)

    @staticmethod
    def default_X0():
        """:458-468 — dr0 = [-0.1 0 0], dv0 = 0, q0 = angle2quat(0, 3 deg, 0) reversed (scalar last), w0 = 0."""
        a = tables.deg2rad(3.0) / 2
        return np.array([-0.1, 0, 0, 0, 0, 0, 0.0, np.sin(a), 0.0, np.cos(a), 0, 0, 0])

    def _channel_sweep(self, channel):
        """Handle holding the channel's installed controller (set_controller) as its policy."""
        ctl = self._channel_ctl[channel]
        sw = self._lookup_sweeps.get(channel)
        if sw is None:
            d = self.channel_desc("xyz".index(channel))
            if len(np.ravel(ctl["f0_allcomb"])) != d.C:                  # a failure-mode controller (:236-240)
                d = self.channel_desc("xyz".index(channel), failure=True)
            if len(np.ravel(ctl["f0_allcomb"])) != d.C:
                raise ValueError("controller combinations differ from this object's thruster settings")
            for k in range(4):
                if not np.array_equal(d.grid[k][0], ctl["GridVectors"][k]):
                    raise ValueError("controller grid differs from this object's mesh settings")
            sw = self._lookup_sweeps[channel] = Sweep(d, device=self.device)
            stage = max(1, min(int(ctl.get("stop_stage", 1)) or 1, d.N - 1))
            sw.set_stage(stage, None, np.asarray(ctl["U_Optimal_id"]).astype(np.int32).ravel(order="F") - 1)
        return sw

    def get_optimal_path(self, X0=None, n_steps=None, stride_out=1, controllers=None):
        """:452-500 on the GPU for a batch of initial states X0 [batch, 13] = (dr dv q w) (default the
        reference's single state): thruster levels from the three installed channel controllers, moments
        and forces, one ode45 call per stage on the 13-state plant.  ``controllers`` = three files / dicts
        to install first (the reference loads channel_{x,y,z}_controller_1.mat, :469-471).
        Returns X_ode45 [batch, n_out + 1, 13], F_Th_Opt [batch, n_out, 12], Force_Moment_log [batch, n_out, 6]."""
        from ._lib import rollout_pos_att
        if controllers is not None:
            for ch, c in zip("xyz", controllers):
                self.set_controller(c, ch)
        X0 = self.default_X0()[None] if X0 is None else np.asarray(X0, dtype=np.float64).reshape(-1, 13)
        self.N_stage = int(np.ceil(self.T_final / self.h))
        n_steps = self.N_stage - 1 if n_steps is None else int(n_steps)
        sws = [self._channel_sweep(ch) for ch in "xyz"]
        fv = [np.stack([np.asarray(self._channel_ctl[ch][k], dtype=np.float64).ravel()
                        for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")]) for ch in "xyz"]
        R0, V0 = self.get_target_R0V0()
        X, F, FM, W = rollout_pos_att(sws, fv, X0, n_steps, self.h, R0, V0, self.InertiaM, self.Mass, self.T_dist,
                                      mu=self.mu, stride_out=stride_out)
        self.ode45_warnings = W
        return X, F, FM

    def simplified_run(self, save=False, failure_mode=True, n_stages=None):
        """Solver_pos_att.m:197-242: x, y, z channels, then the x-channel failure mode."""
        names = ("channel_x_controller_1", "channel_y_controller_1", "channel_z_controller_1")
        for ch in range(3):
            self.controllers[names[ch]] = self.calculate_one_channel_U_Opt(
                ch, file_name=names[ch] + ".mat" if save else None, n_stages=n_stages)
        if failure_mode:
            nm = "channel_x_controller_1_failure"
            self.controllers[nm] = self.calculate_one_channel_U_Opt(
                0, failure=True, file_name=nm + ".mat" if save else None, n_stages=n_stages)
        return self
