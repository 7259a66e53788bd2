"""Single-process multi-GPU (bellman_group_init / bellman_group_run): ONE host thread drives every slab,
halo values go straight into the neighbouring slabs' buffers, stages are ordered by neighbour flags.
The slabs may share a GPU, so the whole sharded path — slab planning, peer stores, flag protocol, every
TMA-staged kernel with a partitioned grid — is exercised on a single-GPU box; with >= 2 GPUs the same
tests also spread the slabs over the devices.  Every stitched result is compared bit for bit with the
single-process CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KERNELS = {"direct": 1, "auto": 0, "staged": 2}


def _devices(n):
    import torch
    g = torch.cuda.device_count()
    return [r % g for r in range(n)]


def _kirk(bellman, n0=256, n1=256, C=48, N=10):
    o = bellman.Dynamic_Solver()
    t = bellman.tables
    s0, s1, u = t.linspace(-2.5, 3.0, n0), t.linspace(-2.5, 3.0, n1), t.linspace(-40.0, 10.0, C)
    A, B = o.A, o.B.ravel()
    row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
    return t.Desc(n=[n0, n1], C=C, N=N, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
                  Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
                  Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
                  q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u)).validate()


def _check(bellman, oracle_lib, d, n_slabs, part_dim, kernel, n_stages=5, part_cuts=None, JN=None, want=None):
    ora = oracle_lib.sweep(d, n_stages=n_stages, J_N=JN)
    grp = bellman.SweepGroup(d, _devices(n_slabs), part_dim=part_dim, part_cuts=part_cuts)
    if JN is not None:
        grp.set_J(JN)
    grp.run(n_stages, kernel=KERNELS[kernel])
    if want:
        assert grp.last_kernel == want, grp.last_kernel
    J, I = grp.get_J(), grp.get_idx()
    grp.close()
    assert np.array_equal(I, ora["idx_last"]), "argmin differs on %d states" % np.sum(I != ora["idx_last"])
    assert np.array_equal(J, ora["J_last"])


@pytest.mark.parametrize("n_slabs", [2, 3])
@pytest.mark.parametrize("part_dim", [0, 1])
@pytest.mark.parametrize("kernel", ["direct", "staged"])
def test_group_kirk(bellman, oracle_lib, n_slabs, part_dim, kernel):
    d = _kirk(bellman)
    _check(bellman, oracle_lib, d, n_slabs, part_dim, kernel, want="window:wide" if kernel == "staged" else "direct")


def test_group_kirk_odd_leading_dimension_and_uneven_cuts(bellman, oracle_lib):
    """255 rows (padded to 256 in every slab's layout) cut along dimension 1, then explicit slab
    boundaries (bellman_desc.part_cuts) along dimension 0, rough terminal cost."""
    d = _kirk(bellman, n0=255, n1=120, C=24, N=12)
    rng = np.random.default_rng(2)
    JN = rng.normal(size=(1, d.S)) * 3
    _check(bellman, oracle_lib, d, 2, 1, "staged", JN=JN)
    _check(bellman, oracle_lib, d, 3, 0, "auto", part_cuts=[0, 60, 170, 255], JN=JN)
    assert bellman.plan_slabs(d, 0, 3, part_cuts=[0, 60, 170, 255])[1][:2] == (60, 170)


def test_group_attitude_strip_kernel(bellman, oracle_lib):
    s = bellman.Solver_attitude()
    s.n_mesh_w, s.n_mesh_t = 400, 120
    d = bellman.tables.stack_problems(s._axis_descs())
    for part_dim in (0, 1):
        _check(bellman, oracle_lib, d, 2, part_dim, "staged", n_stages=6, want="window:strip")


@pytest.mark.parametrize("part_dim", [0, 2, 3])
def test_group_pos_att_stream_kernel(bellman, oracle_lib, part_dim):
    """the 4-D channel cut along x, theta or w (the walked dimension): k_stage_stream with peer stores;
    then the reference's Sigma-check (Solver_pos_att.m:273-285) summed over the slabs."""
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 24, 10, 12, 15
    d = sp.channel_desc(0)
    _check(bellman, oracle_lib, d, 2, part_dim, "staged", n_stages=6, want="stream")
    grp = bellman.SweepGroup(d, _devices(2), part_dim=part_dim)
    grp.run(100, check_period=50, check_tol=0.0)
    one = bellman.Sweep(d).run(100, check_period=50, check_tol=0.0)
    lg, l1 = grp.check_log(), one.check_log()
    assert lg.shape == l1.shape == (2, 3)
    assert np.array_equal(lg[:, [0, 2]], l1[:, [0, 2]])                 # stage numbers and sum(idx) exactly
    np.testing.assert_allclose(lg[:, 1], l1[:, 1], rtol=1e-13)           # sum(J): different summation order
    grp.close()
    one.close()


def test_facade_n_gpus(bellman, oracle_lib):
    """Solver_attitude with n_gpus = 2 (one process): same kept policy as the one-GPU sweep, and the
    consumers of the policy (nearest lookup, simplified rollout) work on the gathered result."""
    a = bellman.Solver_attitude()
    a.n_mesh_w, a.n_mesh_t = 160, 64
    a.simplified_run(n_stages=20)
    b = bellman.Solver_attitude()
    b.n_mesh_w, b.n_mesh_t = 160, 64
    b.n_gpus, b.devices = 2, _devices(2)
    b.simplified_run(n_stages=20)
    for p in range(3):
        assert np.array_equal(a.U_idx[p], b.U_idx[p]) and np.array_equal(a.F_Values[p], b.F_Values[p])
    # nearest lookup straight on the SHARDED policy (owner answers, the other slabs say -1)
    grp = bellman.SweepGroup(a._desc, _devices(3), part_dim=0)
    grp.run(20)
    xq = np.random.default_rng(5).uniform(-1.0, 1.0, size=(500, 2)) * [1.0, 0.6]
    for p in range(3):
        assert np.array_equal(grp.policy_lookup(xq, prob=p), a._sweep.policy_lookup(xq, prob=p))
    parts = [s.policy_lookup(xq, prob=0) for s in grp.slabs]
    assert all(np.sum(np.stack(parts) >= 0, axis=0) == 1)          # exactly one owner per query
    with pytest.raises(bellman.BellmanError):
        grp.slabs[0].rollout_axis(a._desc.Tc[0][0], xq[:4], 5, 0.005, 0)
    grp.close()
    x0 = np.tile(np.array([[0.1, -0.05]]), (3, 1))[None]
    Xa, Ua = a.get_optimal_path_simplified(x0, n_steps=15)
    Xb, Ub = b.get_optimal_path_simplified(x0, n_steps=15)
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)


def test_handles_with_different_window_sizes_coexist(bellman, oracle_lib):
    """The dynamic shared-memory limit belongs to the kernel FUNCTION, not to a handle: slabs (or two
    sweeps) that plan different window sizes for the same kernel must not lower each other's limit
    (96 x 96 Kirk grid cut in two: the two slabs' halos, hence their windows, differ)."""
    o = bellman.Dynamic_Solver()
    small = bellman.tables.kirk_desc(o.A, o.B, o.Q, o.R, 6, o.x_min, o.x_max, 96, o.u_min, o.u_max, 24,
                                     store_J_all=False, store_idx_all=False)
    _check(bellman, oracle_lib, small, 2, None, "staged", n_stages=3, want="window:wide")
    big = _kirk(bellman, n0=512, n1=384, C=64)
    a = bellman.Sweep(big)
    b = bellman.Sweep(small)          # created later, plans a smaller window for the same kernel
    a.run(2, kernel=KERNELS["staged"])
    b.run(2, kernel=KERNELS["staged"])
    oa, ob = oracle_lib.sweep(big, n_stages=2), oracle_lib.sweep(small, n_stages=2)
    assert np.array_equal(a.get_J(), oa["J_last"]) and np.array_equal(b.get_J(), ob["J_last"])
    a.close()
    b.close()


def test_facade_n_gpus_dynamic_solver_and_pos_att(bellman, oracle_lib):
    """n_gpus on the two other facades: Dynamic_Solver (every stage's J_star / u_star gathered, rollouts on
    the gathered policy) and Solver_pos_att (channel sweeps with the reference's Sigma-check and early stop)."""
    def mk(n_gpus):
        o = bellman.Dynamic_Solver()
        o.dx, o.du, o.N = 96, 40, 30
        o.n_gpus, o.devices = n_gpus, (_devices(n_gpus) if n_gpus > 1 else None)
        return o.run()
    a, b = mk(1), mk(2)
    assert np.array_equal(a.J_star, b.J_star) and np.array_equal(a.u_star, b.u_star)
    assert np.array_equal(a.u_star_idx, b.u_star_idx)
    x0 = np.random.default_rng(4).uniform(-2.0, 2.5, size=(50, 2))
    for mode, ssu in (("Nssu", 1), ("ssu", 7)):
        Xa, Ua = a.get_optimal_path(x0, mode, ssu)
        Xb, Ub = b.get_optimal_path(x0, mode, ssu)
        np.testing.assert_array_equal(Xa, Xb)
        np.testing.assert_array_equal(Ua, Ub)
    b.store_J_star = False                                   # only stage 1 of J kept; the policies of all stages still are
    b.run()
    assert np.array_equal(b.F_Values, a.F_Values)
    np.testing.assert_array_equal(b.get_optimal_path(x0)[0], a.get_optimal_path(x0)[0])

    def ch(n_gpus):
        sp = bellman.Solver_pos_att()
        sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 24, 10, 12, 15
        sp.check_tol = 0.0
        sp.n_gpus, sp.devices = n_gpus, (_devices(n_gpus) if n_gpus > 1 else None)
        return sp.calculate_one_channel_U_Opt(1, n_stages=120)
    c1, c2 = ch(1), ch(2)
    assert np.array_equal(c1["U_Optimal_id"], c2["U_Optimal_id"]) and np.array_equal(c1["F_gI_Values"], c2["F_gI_Values"])
    assert c1["stop_stage"] == c2["stop_stage"]
    assert np.array_equal(c1["check_log"][:, [0, 2]], c2["check_log"][:, [0, 2]])
