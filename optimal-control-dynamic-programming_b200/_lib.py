"""ctypes binding of libbellman.so — the same C ABI (include/bellman.h) the MEX gateway binds.

There is no fallback of any kind: if the shared library is missing, or no sm_100 device is
usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbellman.so")
MAXD = 4
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

KERNEL_AUTO, KERNEL_DIRECT, KERNEL_WINDOW, KERNEL_SPLITC, KERNEL_TILE = 0, 1, 2, 3, 4
LOCATE_UNIFORM, LOCATE_SEARCH = 0, 1

STATUS = {0: "OK", -1: "BAD_ARG", -2: "CUDA", -3: "NCCL", -4: "OOM", -5: "NOT_RUN", -6: "STATE"}


class BellmanError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bellman:{STATUS.get(code, code)}: {msg}")
        self.code = code


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


class CRunOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("kernel", C.c_int32), ("check_period", C.c_int32),
                ("check_tol", C.c_double), ("use_graph", C.c_int32), ("sync_each_stage", C.c_int32)]


class COrbitOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_steps", C.c_int32), ("stride_out", C.c_int32),
                ("max_rkf_steps", C.c_int32), ("mu", C.c_double), ("R0", C.c_double * 3), ("V0", C.c_double * 3),
                ("h", C.c_double), ("tol", C.c_double)]


class CPlantOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_steps", C.c_int32), ("stride_out", C.c_int32),
                ("max_ode_steps", C.c_int32), ("mu", C.c_double), ("R0", C.c_double * 3), ("V0", C.c_double * 3),
                ("h", C.c_double), ("rtol", C.c_double), ("atol", C.c_double), ("inertia", C.c_double * 9),
                ("mass", C.c_double), ("t_dist", C.c_double)]


class CDense6(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n", C.c_int32 * 6), ("nu", C.c_int32), ("device", C.c_int32),
                ("grid", _dp * 6), ("w_next", _dp * 3), ("a_next", _dp * 3), ("gs", _dp), ("r", _dp * 3)]


class CSlab(C.Structure):
    _fields_ = [("own_lo", C.c_int32), ("own_hi", C.c_int32), ("ext_lo", C.c_int32), ("ext_hi", C.c_int32)]


# every symbol include/bellman.h declares
EXPORTS = [
    "bellman_version", "bellman_last_error", "bellman_query_locate", "bellman_plan_slabs", "bellman_query_stencil",
    "bellman_create", "bellman_destroy", "bellman_get_unique_id", "bellman_comm_init", "bellman_halo_mode",
    "bellman_set_J", "bellman_set_stage", "bellman_stage", "bellman_stage_host", "bellman_run", "bellman_current_stage", "bellman_get_J",
    "bellman_get_idx", "bellman_get_check_log", "bellman_owned_range", "bellman_last_run_stats",
    "bellman_last_kernel", "bellman_rollout", "bellman_policy_lookup", "bellman_rollout_axis",
    "bellman_rollout_orbit", "bellman_get_points", "bellman_group_init", "bellman_group_run",
    "bellman_rollout_pos_att", "bellman_rollout_attitude", "bellman_dense6_run", "bellman_rollout_attitude6",
]

_lib = None


def load():
    """Load libbellman.so; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BellmanError(-2, f"{LIB_PATH} is missing: build it with csrc/build.sh "
                               "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.bellman_last_error.restype = C.c_char_p
    lib.bellman_last_error.argtypes = [C.c_void_p]
    lib.bellman_last_kernel.restype = C.c_char_p
    lib.bellman_last_kernel.argtypes = [C.c_void_p]
    lib.bellman_create.argtypes = [C.POINTER(CDesc), C.POINTER(C.c_void_p)]
    lib.bellman_destroy.argtypes = [C.c_void_p]
    lib.bellman_destroy.restype = None
    lib.bellman_run.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CRunOpts)]
    lib.bellman_stage.argtypes = [C.c_void_p]
    lib.bellman_stage_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bellman_set_J.argtypes = [C.c_void_p, _dp]
    lib.bellman_set_stage.argtypes = [C.c_void_p, C.c_int32, _dp, _ip]
    lib.bellman_get_J.argtypes = [C.c_void_p, C.c_int32, _dp]
    lib.bellman_get_idx.argtypes = [C.c_void_p, C.c_int32, _ip]
    lib.bellman_current_stage.argtypes = [C.c_void_p]
    lib.bellman_get_check_log.argtypes = [C.c_void_p, _dp, C.c_int32]
    lib.bellman_owned_range.argtypes = [C.c_void_p, C.POINTER(CSlab)]
    lib.bellman_last_run_stats.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int64), _dp]
    lib.bellman_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.bellman_halo_mode.argtypes = [C.c_void_p]
    lib.bellman_get_unique_id.argtypes = [C.c_void_p]
    lib.bellman_query_locate.argtypes = [C.POINTER(CDesc), _ip]
    lib.bellman_plan_slabs.argtypes = [C.POINTER(CDesc), C.c_int32, C.c_int32, C.POINTER(CSlab)]
    lib.bellman_rollout.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]
    lib.bellman_policy_lookup.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, C.c_int32, _ip]
    lib.bellman_rollout_axis.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp,
                                         _dp, C.c_int32, C.c_int32, _dp, _ip]
    lib.bellman_group_init.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.bellman_group_run.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.POINTER(CRunOpts)]
    lib.bellman_get_points.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int64, _dp, _ip]
    lib.bellman_rollout_attitude6.argtypes = [C.POINTER(CDense6), _ip, _dp, _dp, C.c_double, C.c_int32, _dp, C.c_int32, _dp, _dp]
    lib.bellman_dense6_run.argtypes = [C.POINTER(CDense6), C.c_int32, _dp, _dp, _ip, C.POINTER(C.c_float)]
    lib.bellman_rollout_pos_att.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _ip, C.POINTER(CPlantOpts), _dp, _dp, _dp,
                                            _dp, C.c_int32, _dp, _dp, _dp, _ip]
    lib.bellman_rollout_attitude.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CPlantOpts), _dp, _dp, C.c_int32, _dp, _ip, _ip]
    lib.bellman_rollout_orbit.argtypes = [C.c_void_p, C.c_int32, C.POINTER(COrbitOpts), _dp, _dp, C.c_int32, _dp, _ip, _ip]
    _lib = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def to_cdesc(d, device=-1, part_dim=-1, rank=0, nranks=1, part_cuts=None, idx_bytes=0):
    """tables.Desc -> (CDesc, keepalive)."""
    cd = CDesc()
    keep = []
    D = d.D
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
                a = _f64(a)
                keep.append(a)
                getattr(cd, name)[k] = a.ctypes.data_as(_dp)
    r = _f64(d.r)
    keep.append(r)
    cd.r = r.ctypes.data_as(_dp)
    cd.store_J_all = int(bool(d.store_J_all))
    cd.store_idx_all = int(bool(d.store_idx_all))
    cd.device, cd.part_dim, cd.rank, cd.nranks = int(device), int(part_dim), int(rank), int(nranks)
    cd.idx_bytes = int(idx_bytes or getattr(d, "idx_bytes", 0) or 0)
    if part_cuts is not None:
        pc = np.ascontiguousarray(part_cuts, dtype=np.int32)
        assert len(pc) == nranks + 1
        keep.append(pc)
        cd.part_cuts = pc.ctypes.data_as(_ip)
    return cd, keep


def query_locate(d):
    """Locate mode the library uses per (problem, dim): [P, D] of LOCATE_*.  Host-only."""
    lib = load()
    cd, keep = to_cdesc(d)
    modes = np.zeros(d.P * d.D, dtype=np.int32)
    rc = lib.bellman_query_locate(C.byref(cd), modes.ctypes.data_as(_ip))
    if rc != 0:
        raise BellmanError(rc, "bellman_query_locate")
    return modes.reshape(d.P, d.D)


def query_stencil(d):
    """Exact per-dimension bounds (lo, hi) of cell(x'_d) - i_d over every state and control.  Host-only."""
    lib = load()
    cd, keep = to_cdesc(d)
    lo = np.zeros(d.D, dtype=np.int32)
    hi = np.zeros(d.D, dtype=np.int32)
    rc = lib.bellman_query_stencil(C.byref(cd), lo.ctypes.data_as(_ip), hi.ctypes.data_as(_ip))
    if rc != 0:
        raise BellmanError(rc, "bellman_query_stencil")
    return lo, hi


def plan_slabs(d, part_dim, nranks, part_cuts=None):
    """Slab plan [(own_lo, own_hi, ext_lo, ext_hi)] per rank from the exact reach analysis.  Host-only."""
    lib = load()
    cd, keep = to_cdesc(d, nranks=nranks, part_cuts=part_cuts)
    slabs = (CSlab * nranks)()
    rc = lib.bellman_plan_slabs(C.byref(cd), part_dim, nranks, slabs)
    if rc != 0:
        raise BellmanError(rc, "bellman_plan_slabs")
    return [(s.own_lo, s.own_hi, s.ext_lo, s.ext_hi) for s in slabs]


def get_unique_id():
    lib = load()
    buf = (C.c_char * 128)()
    rc = lib.bellman_get_unique_id(buf)
    if rc != 0:
        raise BellmanError(rc, lib.bellman_last_error(None).decode())
    return bytes(buf)


class Sweep:
    """One ``bellman_handle``: device-resident tables + J/idx storage for a (batched) problem."""

    def __init__(self, desc, device=-1, part_dim=-1, rank=0, nranks=1, part_cuts=None, idx_bytes=0):
        """idx_bytes: device storage of the argmin (0 / 4 = int32, 2 = uint16, 1 = uint8); every call
        still takes and returns int32 indices."""
        self.lib = load()
        self.desc = desc
        cd, keep = to_cdesc(desc, device, part_dim, rank, nranks, part_cuts, idx_bytes)
        h = C.c_void_p()
        rc = self.lib.bellman_create(C.byref(cd), C.byref(h))
        if rc != 0:
            raise BellmanError(rc, self.lib.bellman_last_error(None).decode())
        self.h = h
        s = CSlab()
        self.lib.bellman_owned_range(self.h, C.byref(s))
        self.part_dim = part_dim
        self.slab = (s.own_lo, s.own_hi, s.ext_lo, s.ext_hi)
        self.own_shape = list(desc.n)
        if part_dim >= 0:
            self.own_shape[part_dim] = s.own_hi - s.own_lo
        self.S_own = int(np.prod(self.own_shape))

    def _check(self, rc):
        if rc != 0:
            raise BellmanError(rc, self.lib.bellman_last_error(self.h).decode())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if getattr(self, "h", None):
            self.lib.bellman_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, id128):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._check(self.lib.bellman_comm_init(self.h, buf))

    @property
    def halo_mode(self):
        """'p2p' (stage kernel stores halos into peer memory) or 'nccl' (send/recv after each stage)."""
        return "p2p" if self.lib.bellman_halo_mode(self.h) else "nccl"

    def set_J(self, J=None):
        if J is None:
            self._check(self.lib.bellman_set_J(self.h, _dp()))
        else:
            J = _f64(J).reshape(self.desc.P, self.desc.S)
            self._check(self.lib.bellman_set_J(self.h, J.ctypes.data_as(_dp)))

    def set_stage(self, stage, J=None, idx=None):
        """Resume from / load a saved controller: J [P, S] (global) and 0-based idx [P, S_own] of ``stage``."""
        Jp = _dp() if J is None else _f64(J).reshape(self.desc.P, self.desc.S)
        ip = _ip()
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(self.desc.P, -1)
            ip = idx.ctypes.data_as(_ip)
        self._check(self.lib.bellman_set_stage(self.h, int(stage), Jp if J is None else Jp.ctypes.data_as(_dp), ip))

    def run(self, n_stages=None, kernel=KERNEL_AUTO, check_period=0, check_tol=0.0, use_graph=False,
            sync_each_stage=False):
        if n_stages is None:
            n_stages = self.current_stage - 1
        o = CRunOpts(C.sizeof(CRunOpts), kernel, check_period, check_tol, int(use_graph), int(sync_each_stage))
        self._check(self.lib.bellman_run(self.h, int(n_stages), C.byref(o)))
        return self

    def stage(self):
        self._check(self.lib.bellman_stage(self.h))

    def stage_host(self, J_next=None, J_out=None, idx_out=None, kernel=KERNEL_AUTO):
        """One stage from host arrays with the copies overlapped with the kernel (bellman_stage_host):
        ``J_next`` [P, S] (None = continue from the device's J), results into ``J_out`` [P, S_own] /
        ``idx_out`` (allocated when None).  Arrays may be numpy or (pinned) torch CPU tensors' numpy views."""
        if J_next is not None:
            J_next = np.ascontiguousarray(J_next, dtype=np.float64)
            if J_next.size != self.desc.P * self.desc.S:
                raise ValueError("J_next must hold P*S values")
        if J_out is None:
            J_out = np.empty((self.desc.P, self.S_own), dtype=np.float64)
        if idx_out is None:
            idx_out = np.empty((self.desc.P, self.S_own), dtype=np.int32)
        o = CRunOpts(C.sizeof(CRunOpts), kernel, 0, 0.0, 0, 0)
        self._check(self.lib.bellman_stage_host(self.h, None if J_next is None else J_next.ctypes.data, J_out.ctypes.data,
                                                idx_out.ctypes.data, C.byref(o)))
        return J_out, idx_out

    @property
    def current_stage(self):
        return int(self.lib.bellman_current_stage(self.h))

    @property
    def last_kernel(self):
        return self.lib.bellman_last_kernel(self.h).decode()

    def stats(self):
        ms, n, mx = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.bellman_last_run_stats(self.h, C.byref(ms), C.byref(n), C.byref(mx)))
        return {"ms": ms.value, "launches": n.value, "ms_exchange": mx.value}

    def get_J(self, stage=None, out=None):
        """J of ``stage`` (default: current) for the owned slab: [P, S_own], column-major states."""
        stage = self.current_stage if stage is None else stage
        if out is None:
            out = np.empty((self.desc.P, self.S_own), dtype=np.float64)
        self._check(self.lib.bellman_get_J(self.h, int(stage), out.ctypes.data_as(_dp)))
        return out

    def get_points(self, states, prob=0, stage=None, with_idx=True):
        """J (and argmin) of ``stage`` at GLOBAL linear state indices that lie in this rank's slab."""
        stage = self.current_stage if stage is None else stage
        st = np.ascontiguousarray(states, dtype=np.int64)
        J = np.empty(len(st), dtype=np.float64)
        I = np.empty(len(st), dtype=np.int32) if with_idx else None
        self._check(self.lib.bellman_get_points(self.h, int(stage), int(prob), st.ctypes.data_as(C.POINTER(C.c_int64)),
                                                len(st), J.ctypes.data_as(_dp), I.ctypes.data_as(_ip) if with_idx else _ip()))
        return J, I

    def get_idx(self, stage=None, out=None):
        """0-based argmin control index of ``stage``: [P, S_own] int32."""
        stage = self.current_stage if stage is None else stage
        if out is None:
            out = np.empty((self.desc.P, self.S_own), dtype=np.int32)
        self._check(self.lib.bellman_get_idx(self.h, int(stage), out.ctypes.data_as(_ip)))
        return out

    def check_log(self):
        n = self.lib.bellman_get_check_log(self.h, _dp(), 0)
        out = np.zeros((max(n, 0), 3))
        if n > 0:
            self.lib.bellman_get_check_log(self.h, out.ctypes.data_as(_dp), n)
        return out

    def rollout(self, A, B, u_values, x0, mode=0, ssu_stage=1, out=None):
        """x0: [batch, 2].  Returns X [batch, N, 2], U [batch, N] (``out=(X, U)``: preallocated, e.g. pinned, arrays)."""
        x0 = _f64(x0).reshape(-1, 2)
        batch = x0.shape[0]
        N = self.desc.N
        if out is not None:
            X, U = out
            if X.dtype != np.float64 or U.dtype != np.float64 or X.size != batch * N * 2 or U.size != batch * N \
                    or not (X.flags.c_contiguous and U.flags.c_contiguous):
                raise ValueError("out must be C-contiguous float64 arrays of shapes [batch, N, 2] and [batch, N]")
        else:
            X = np.empty((batch, N, 2))
            U = np.empty((batch, N))
        A = _f64(np.asarray(A, dtype=np.float64).reshape(2, 2).ravel(order="F"))
        B = _f64(np.asarray(B, dtype=np.float64).ravel())
        uv = _f64(u_values)
        self._check(self.lib.bellman_rollout(self.h, A.ctypes.data_as(_dp), B.ctypes.data_as(_dp),
                                             uv.ctypes.data_as(_dp), x0.ctypes.data_as(_dp), batch,
                                             int(mode), int(ssu_stage), X.ctypes.data_as(_dp),
                                             U.ctypes.data_as(_dp)))
        return X, U

    def policy_lookup(self, x, prob=0, stage=None):
        """'nearest' policy lookup at states x [batch, D] of problem ``prob``: 0-based control index."""
        stage = self.current_stage if stage is None else stage
        x = _f64(x).reshape(-1, self.desc.D)
        out = np.empty(len(x), dtype=np.int32)
        self._check(self.lib.bellman_policy_lookup(self.h, int(prob), int(stage), x.ctypes.data_as(_dp), len(x),
                                                   out.ctypes.data_as(_ip)))
        return out

    def rollout_axis(self, u_inc, x0, n_steps, h_step, rate_dim, prob=0, time_varying=False, stage=None):
        """Simplified-plant rollout under the nearest policy (two-state axis problems).
        x0 [batch, 2] -> X [batch, n_steps+1, 2], control indices [batch, n_steps]."""
        stage = self.current_stage if stage is None else stage
        x0 = _f64(x0).reshape(-1, 2)
        batch = len(x0)
        X = np.empty((batch, n_steps + 1, 2))
        Cc = np.empty((batch, n_steps), dtype=np.int32)
        ui = _f64(u_inc)
        self._check(self.lib.bellman_rollout_axis(self.h, int(prob), int(bool(time_varying)), int(stage), int(rate_dim),
                                                  float(h_step), ui.ctypes.data_as(_dp), x0.ctypes.data_as(_dp), batch,
                                                  int(n_steps), X.ctypes.data_as(_dp), Cc.ctypes.data_as(_ip)))
        return X, Cc

    def rollout_orbit(self, u_values, y0, n_steps, h_step, R0, V0, mu=398600.0, tol=1e-8, stride_out=1, stage=None):
        """Orbital forward simulation of Solver_position.get_optimal_path (Solver_position.m:189-224):
        y0 [batch, 6] -> X [batch, n_steps/stride_out + 1, 6], control indices [batch, n_steps/stride_out, 3],
        rkf45 minimum-step warnings [batch]."""
        stage = self.current_stage if stage is None else stage
        y0 = _f64(y0).reshape(-1, 6)
        batch = len(y0)
        n_out = int(n_steps) // int(stride_out)
        X = np.empty((batch, n_out + 1, 6))
        Cc = np.empty((batch, n_out, 3), dtype=np.int32)
        W = np.empty(batch, dtype=np.int32)
        o = COrbitOpts(C.sizeof(COrbitOpts), int(n_steps), int(stride_out), 0, float(mu), (C.c_double * 3)(*R0),
                       (C.c_double * 3)(*V0), float(h_step), float(tol))
        uv = _f64(u_values)
        self._check(self.lib.bellman_rollout_orbit(self.h, int(stage), C.byref(o), uv.ctypes.data_as(_dp),
                                                   y0.ctypes.data_as(_dp), batch, X.ctypes.data_as(_dp),
                                                   Cc.ctypes.data_as(_ip), W.ctypes.data_as(_ip)))
        return X, Cc, W

    def rollout_attitude(self, u_values, y0, n_steps, h_step, InertiaM, rtol=1e-3, atol=1e-6, stride_out=1, stage=None):
        """Solver_attitude.get_optimal_path_simplified_testode45 (Solver_attitude.m:1669-1705): y0 [batch, 7]
        = (w1 w2 w3 q1 q2 q3 q4) -> X [batch, n_out + 1, 7], control indices [batch, n_out, 3], warnings [batch]."""
        stage = self.current_stage if stage is None else stage
        y0 = _f64(y0).reshape(-1, 7)
        batch = len(y0)
        n_out = int(n_steps) // int(stride_out)
        X = np.empty((batch, n_out + 1, 7))
        Cc = np.empty((batch, n_out, 3), dtype=np.int32)
        W = np.empty(batch, dtype=np.int32)
        o = _plant_opts(n_steps, stride_out, h_step, InertiaM, rtol=rtol, atol=atol)
        uv = _f64(u_values)
        self._check(self.lib.bellman_rollout_attitude(self.h, int(stage), C.byref(o), uv.ctypes.data_as(_dp),
                                                      y0.ctypes.data_as(_dp), batch, X.ctypes.data_as(_dp),
                                                      Cc.ctypes.data_as(_ip), W.ctypes.data_as(_ip)))
        return X, Cc, W


def dense6_run(T, n_stages, J_N=None, device=-1):
    """The coupled 6-D attitude sweep (Solver_attitude.run, Solver_attitude.m:521-601) on the tables of
    tables.attitude6_tables, entirely on the GPU (bellman_dense6_run).  Returns (J [S], idx [S] 0-based
    with idx = (u1*nu + u2)*nu + u3, device milliseconds of the stage loop)."""
    lib = load()
    keep = [_f64(g) for g in T.grid] + [_f64(w) for w in T.w_next] + [_f64(a) for a in T.a_next] + [_f64(x) for x in T.r]
    S, S3 = int(np.prod(T.n)), int(np.prod(T.n[:3]))
    # the C ABI takes bare pointers: check every extent here so a malformed table cannot be read past its end
    if not 1 <= int(T.nu) <= 8:
        raise BellmanError(-1, "nu must be in 1..8")
    for k in range(6):
        if keep[k].size != T.n[k]:
            raise BellmanError(-1, "grid[%d] has %d points, n says %d" % (k, keep[k].size, T.n[k]))
    for k in range(3):
        if keep[6 + k].size != T.nu * S3 or keep[9 + k].size != S or keep[12 + k].size != T.nu:
            raise BellmanError(-1, "w_next / a_next / r of control %d do not match n and nu" % k)
    if np.size(T.gs) != S or (J_N is not None and np.size(J_N) != S):
        raise BellmanError(-1, "gs / J_N must have one entry per state")
    cd = CDense6()
    cd.struct_size, cd.nu, cd.device = C.sizeof(CDense6), int(T.nu), int(device)
    for k in range(6):
        cd.n[k] = int(T.n[k])
        cd.grid[k] = keep[k].ctypes.data_as(_dp)
    for k in range(3):
        cd.w_next[k] = keep[6 + k].ctypes.data_as(_dp)
        cd.a_next[k] = keep[9 + k].ctypes.data_as(_dp)
        cd.r[k] = keep[12 + k].ctypes.data_as(_dp)
    gs = _f64(T.gs)
    cd.gs = gs.ctypes.data_as(_dp)
    J = np.empty(T.S)
    idx = np.empty(T.S, dtype=np.int32)
    ms = C.c_float(0)
    JN = None if J_N is None else _f64(J_N).ravel()
    rc = lib.bellman_dense6_run(C.byref(cd), int(n_stages), _dp() if JN is None else JN.ctypes.data_as(_dp),
                                J.ctypes.data_as(_dp), idx.ctypes.data_as(_ip), C.byref(ms))
    if rc != 0:
        raise BellmanError(rc, lib.bellman_last_error(None).decode())
    return J, idx, float(ms.value)


def rollout_attitude6(T, idx, J123, h, n_steps, x0, device=-1):
    """Solver_attitude.get_optimal_path (Solver_attitude.m:1487-1530) under the 6-D policy idx [S] for a batch
    x0 [batch, 7] on the GPU (bellman_rollout_attitude6).  Returns X [batch, n_steps+1, 7], U [batch, n_steps, 3]."""
    lib = load()
    keep = [_f64(g) for g in T.grid]
    cd = CDense6()
    cd.struct_size, cd.nu, cd.device = C.sizeof(CDense6), int(T.nu), int(device)
    for k in range(6):
        cd.n[k] = int(T.n[k])
        cd.grid[k] = keep[k].ctypes.data_as(_dp)
    ia = np.ascontiguousarray(idx, dtype=np.int32).ravel()
    uv, jd = _f64(T.U_vector), _f64(J123)
    if ia.size != int(np.prod(T.n)) or uv.size != T.nu or jd.size != 3 or any(keep[k].size != T.n[k] for k in range(6)):
        raise BellmanError(-1, "idx / u_values / J123 / grid do not match n and nu")
    x0 = _f64(x0).reshape(-1, 7)
    batch = len(x0)
    X = np.empty((batch, int(n_steps) + 1, 7))
    U = np.empty((batch, int(n_steps), 3))
    rc = lib.bellman_rollout_attitude6(C.byref(cd), ia.ctypes.data_as(_ip), uv.ctypes.data_as(_dp), jd.ctypes.data_as(_dp),
                                       float(h), int(n_steps), x0.ctypes.data_as(_dp), batch, X.ctypes.data_as(_dp),
                                       U.ctypes.data_as(_dp))
    if rc != 0:
        raise BellmanError(rc, lib.bellman_last_error(None).decode())
    return X, U


def _plant_opts(n_steps, stride_out, h_step, InertiaM, mu=0.0, R0=(0, 0, 0), V0=(0, 0, 0), rtol=1e-3, atol=1e-6,
                mass=0.0, t_dist=0.0):
    im = np.asarray(InertiaM, dtype=np.float64).reshape(3, 3).ravel(order="F")
    return CPlantOpts(C.sizeof(CPlantOpts), int(n_steps), int(stride_out), 0, float(mu), (C.c_double * 3)(*R0),
                      (C.c_double * 3)(*V0), float(h_step), float(rtol), float(atol), (C.c_double * 9)(*im),
                      float(mass), float(t_dist))


def rollout_pos_att(sweeps, f_values, y0, n_steps, h_step, R0, V0, InertiaM, Mass, T_dist, mu=398600.0, rtol=1e-3,
                    atol=1e-6, stride_out=1, stages=None):
    """Solver_pos_att.get_optimal_path (Solver_pos_att.m:452-500) for a batch of initial states on the GPU
    (bellman_rollout_pos_att).  sweeps: the x, y, z channel Sweeps holding their policies; f_values: per
    channel [4, C] thruster levels (f0/f1/f6/f7_allcomb); y0 [batch, 13].
    Returns X [batch, n_out + 1, 13], F_Th_Opt [batch, n_out, 12], Force_Moment_log [batch, n_out, 6],
    ode45 minimum-step warnings [batch]."""
    lib = sweeps[0].lib
    stages = [sw.current_stage for sw in sweeps] if stages is None else list(stages)
    y0 = _f64(y0).reshape(-1, 13)
    batch = len(y0)
    n_out = int(n_steps) // int(stride_out)
    X = np.empty((batch, n_out + 1, 13))
    F = np.empty((batch, n_out, 12))
    FM = np.empty((batch, n_out, 6))
    W = np.empty(batch, dtype=np.int32)
    o = _plant_opts(n_steps, stride_out, h_step, InertiaM, mu, R0, V0, rtol, atol, Mass, T_dist)
    fv = [_f64(f).reshape(4, -1) for f in f_values]
    for sw, f in zip(sweeps, fv):
        if f.shape[1] != int(sw.desc.C):
            raise BellmanError(-1, "f_values must be [4, C] with C = %d combinations of the channel, got %s" % (sw.desc.C, f.shape))
    st = (C.c_int32 * 3)(*[int(s) for s in stages])
    sweeps[0]._check(lib.bellman_rollout_pos_att(sweeps[0].h, sweeps[1].h, sweeps[2].h, st, C.byref(o),
                                                 fv[0].ctypes.data_as(_dp), fv[1].ctypes.data_as(_dp), fv[2].ctypes.data_as(_dp),
                                                 y0.ctypes.data_as(_dp), batch, X.ctypes.data_as(_dp), F.ctypes.data_as(_dp),
                                                 FM.ctypes.data_as(_dp), W.ctypes.data_as(_ip)))
    return X, F, FM, W


class SweepGroup:
    """Single-process multi-GPU: n slabs of one problem, driven together by ONE host thread
    (bellman_group_init / bellman_group_run).  ``devices`` lists the GPU of every slab; several slabs
    may share a GPU (``devices=[0, 0]`` exercises the sharded path on a single-GPU box)."""

    def __init__(self, desc, devices, part_dim=None, part_cuts=None, idx_bytes=0):
        self.desc = desc
        n = len(devices)
        if part_dim is None:           # the dimension with the smallest halo, from the host-side reach analysis
            best = None
            for pd in range(desc.D):
                try:
                    sl = plan_slabs(desc, pd, n, part_cuts)
                except BellmanError:
                    continue
                cost = max((e - c) / max(b - a, 1) for a, b, c, e in sl)
                if best is None or cost < best[0] - 1e-9:
                    best = (cost, pd)
            part_dim = best[1]
        self.part_dim = part_dim
        self.slabs = [Sweep(desc, device=dev, part_dim=part_dim, rank=r, nranks=n, part_cuts=part_cuts, idx_bytes=idx_bytes)
                      for r, dev in enumerate(devices)]
        self.lib = self.slabs[0].lib
        self._arr = (C.c_void_p * n)(*[s.h for s in self.slabs])
        rc = self.lib.bellman_group_init(self._arr, n)
        if rc != 0:
            raise BellmanError(rc, self.lib.bellman_last_error(self.slabs[0].h).decode())

    def set_J(self, J=None):
        for s in self.slabs:
            s.set_J(J)

    def run(self, n_stages=None, kernel=KERNEL_AUTO, check_period=0, check_tol=0.0):
        if n_stages is None:
            n_stages = self.current_stage - 1
        o = CRunOpts(C.sizeof(CRunOpts), kernel, check_period, check_tol, 0, 0)
        rc = self.lib.bellman_group_run(self._arr, len(self.slabs), int(n_stages), C.byref(o))
        if rc != 0:
            msgs = [self.lib.bellman_last_error(s.h).decode() for s in self.slabs]
            raise BellmanError(rc, "; ".join(m for m in msgs if m))
        return self

    @property
    def current_stage(self):
        return self.slabs[0].current_stage

    @property
    def last_kernel(self):
        return self.slabs[0].last_kernel

    def stats(self):
        return self.slabs[0].stats()

    def _gather(self, parts):
        """Stitch per-slab [P, S_own] arrays (slabs along part_dim) into the global [P, S]."""
        d, pd = self.desc, self.part_dim
        inner = int(np.prod(d.n[:pd]))
        outer = int(np.prod(d.n[pd + 1:]))
        out = np.empty((d.P, outer, d.n[pd], inner), dtype=parts[0].dtype)
        for s, a in zip(self.slabs, parts):
            lo, hi = s.slab[0], s.slab[1]
            out[:, :, lo:hi, :] = a.reshape(d.P, outer, hi - lo, inner)
        return out.reshape(d.P, d.S)

    def get_J(self, stage=None):
        return self._gather([s.get_J(stage) for s in self.slabs])

    def get_idx(self, stage=None):
        return self._gather([s.get_idx(stage) for s in self.slabs])

    def check_log(self):
        return self.slabs[0].check_log()

    def policy_lookup(self, x, prob=0, stage=None):
        """'nearest' policy lookup on the sharded policy: the slab that owns the nearest node answers,
        the others return -1; the element-wise maximum is the lookup."""
        return np.max(np.stack([s.policy_lookup(x, prob=prob, stage=stage) for s in self.slabs]), axis=0)

    def close(self):
        for s in self.slabs:
            s.close()
