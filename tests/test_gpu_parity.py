"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs.  Bar: J bit-exact and argmin index exact against oracle/bellman_oracle.c (same normative
arithmetic), u_star exact and J <= 1e-12 relative against the reference's golden obj_1.mat."""
import numpy as np
import pytest

from conftest import to_grid

pytestmark = pytest.mark.gpu

KERNELS = {"direct": 1, "splitc": 3, "auto": 0, "window": 2, "tile": 4}


def golden_desc(bellman, g):
    return bellman.tables.kirk_desc(g["A"], g["B"], g["Q"], g["R"], g["N"], g["x_min"], g["x_max"], g["dx"],
                                    g["u_min"], g["u_max"], g["du"])


def assert_stage_equal(Jg, Ig, Jo, Io, what):
    assert np.array_equal(Ig, Io), f"{what}: argmin differs at {np.sum(Ig != Io)} of {Ig.size} states"
    assert np.array_equal(Jg, Jo), f"{what}: J differs, max rel {np.max(np.abs(Jg - Jo)) / np.max(np.abs(Jo))}"


@pytest.mark.parametrize("kernel", ["direct", "splitc", "auto"])
def test_golden_full_sweep(bellman, oracle_lib, golden, kernel):
    d = golden_desc(bellman, golden)
    N, dx = golden["N"], golden["dx"]
    ora = oracle_lib.sweep(d, keep_all=True)
    sw = bellman.Sweep(d)
    sw.run(N - 1, kernel=KERNELS[kernel])
    assert sw.current_stage == 1
    U = d.meta["U_mesh"]
    for k in range(1, N):
        Jg, Ig = sw.get_J(k), sw.get_idx(k)
        assert_stage_equal(Jg, Ig, ora["J_all"][k - 1], ora["idx_all"][k - 1], f"stage {k}")
        # and against the reference's own stored run
        assert np.array_equal(U[to_grid(Ig[0], (dx, dx))], golden["u_star"][:, :, k - 1])
        ref = golden["J_star"][:, :, k - 1]
        assert np.max(np.abs(to_grid(Jg[0], (dx, dx)) - ref)) <= 1e-12 * np.max(np.abs(ref))
    sw.close()


def test_dynamic_solver_facade_matches_golden(bellman, golden):
    obj = bellman.Dynamic_Solver()
    obj.N, obj.dx, obj.du = golden["N"], golden["dx"], golden["du"]
    obj.run()
    assert np.array_equal(obj.u_star[:, :, :-1], golden["u_star"][:, :, :-1])
    assert np.max(np.abs(obj.J_star - golden["J_star"])) <= 1e-12 * np.max(np.abs(golden["J_star"]))
    X, U = obj.get_optimal_path([2.0, 1.0])
    np.testing.assert_allclose(U[:5], [-7.30945822, -4.02204735, -2.44352418, -0.69167883, 1.37694508], atol=5e-9)
    np.testing.assert_allclose(X[:, -1], [0.02094326, -0.05365354], atol=5e-9)


def test_rollout_batch_matches_oracle(bellman, oracle_lib, golden):
    d = golden_desc(bellman, golden)
    ora = oracle_lib.sweep(d, keep_all=True)
    sw = bellman.Sweep(d).run()
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-2.6, 3.1, size=(257, 2))        # includes off-grid starts (extrapolation)
    x0[0] = [2.0, 1.0]
    x0[1] = [40.0, -35.0]                             # far off-grid: the open loop diverges to inf/nan
    for mode, ssu in ((0, 1), (1, 30)):
        Xg, Ug = sw.rollout(golden["A"], golden["B"], d.meta["U_mesh"], x0, mode=mode, ssu_stage=ssu)
        Xo, Uo = oracle_lib.rollout(d, ora["idx_all"][:, 0, :], golden["A"], golden["B"], d.meta["U_mesh"], x0,
                                    mode=mode, ssu_stage=ssu)
        np.testing.assert_array_equal(Ug, Uo)          # NaNs (diverged rollouts) compare equal
        np.testing.assert_array_equal(Xg, Xo)
    sw.close()


@pytest.mark.parametrize("kernel", ["direct", "splitc", "auto", "window"])
def test_kirk_reference_size_first_stages(bellman, oracle_lib, kernel):
    """config 1: 100x100 states x 1000 controls (Dynamic_Solver.m:49-63); 24 % of queries off-grid."""
    obj = bellman.Dynamic_Solver()
    d = obj._build()
    ora = oracle_lib.sweep(d, n_stages=6, keep_all=True)
    sw = bellman.Sweep(d).run(6, kernel=KERNELS[kernel])
    if kernel != "auto":
        assert sw.last_kernel.split(":")[0] == kernel
    for k in range(d.N - 6, d.N):
        assert_stage_equal(sw.get_J(k), sw.get_idx(k), ora["J_all"][k - 1], ora["idx_all"][k - 1], f"stage {k}")
    sw.close()


WINDOW_VARIANTS = ["wide", "ring", "wide_bar", "wide_xu0", "wide_xu1"]   # k_stage_wide (default), k_stage_window, k_stage_wide with a CTA barrier per chunk


def _window_variant(monkeypatch, variant):
    """Selects the long-control-loop window kernel; returns the name bellman_last_kernel must report."""
    if variant == "ring":
        monkeypatch.setenv("BELLMAN_NO_WIDE", "1")
        return "window:ring"
    if variant == "wide_bar":
        monkeypatch.setenv("BELLMAN_WIDE_BARRIER", "1")
    if variant.startswith("wide_xu"):      # weights of 0 / 1 dimensions through the conversion pipe (default: both)
        monkeypatch.setenv("BELLMAN_WIDE_XU", variant[-1])
    return "window:wide"


@pytest.mark.parametrize("variant", WINDOW_VARIANTS)
@pytest.mark.parametrize("shape", [(256, 192, 64), (130, 70, 33), (64, 64, 1), (34, 1000, 9)])
def test_window_kernel_random_terminal_cost(bellman, oracle_lib, monkeypatch, shape, variant):
    """TMA-staged kernel: ragged tiles, chunk tails, rough J_N, several stages (ping-pong buffers)."""
    want = _window_variant(monkeypatch, variant)
    n0, n1, C = shape
    rng = np.random.default_rng(5)
    t = bellman.tables
    obj = bellman.Dynamic_Solver()
    s0, s1, u = t.linspace(-2.5, 3.0, n0), t.linspace(-2.5, 3.0, n1), t.linspace(-40.0, 10.0, C) if C > 1 else np.array([-3.0])
    A, B = obj.A, obj.B.ravel()
    row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
    d = t.Desc(n=[n0, n1], C=C, N=6, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
               Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
               Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
               q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u)).validate()
    JN = rng.normal(size=(1, d.S)) * 5
    sw = bellman.Sweep(d)
    sw.set_J(JN)
    sw.run(4, kernel=KERNELS["window"])
    assert sw.last_kernel == want
    ora = oracle_lib.sweep(d, n_stages=4, J_N=JN)
    assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"window {shape}")
    sw.close()


@pytest.mark.parametrize("graph", [False, True, "block256"])
def test_position_reference_size(bellman, oracle_lib, monkeypatch, graph):
    """config 2: 3 axes x 201x201 x 3 controls (Solver_position.m:49-72,84); the persistent kernel with
    1024-thread CTAs (default for D = 2) and with 256-thread CTAs."""
    if graph == "block256":
        monkeypatch.setenv("BELLMAN_PERSIST_BLOCK256", "1")
    sp = bellman.Solver_position()
    d = bellman.tables.stack_problems(sp._axis_descs())
    n = 61
    ora = oracle_lib.sweep(d, n_stages=n)
    sw = bellman.Sweep(d).run(n, use_graph=bool(graph))
    if graph:
        assert sw.last_kernel == "persistent"      # one cooperative launch, grid barrier per stage
    assert sw.current_stage == ora["stage"] == d.N - n
    assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "position")
    sw.close()


def test_persistent_kernel_variants(bellman, oracle_lib, golden, monkeypatch):
    """Persistent multi-stage kernel: control lanes + store-all slots (golden Kirk), D = 4 pos-att,
    and the CUDA-graph fallback when the persistent path is disabled."""
    d = golden_desc(bellman, golden)
    ora = oracle_lib.sweep(d, keep_all=True)
    sw = bellman.Sweep(d).run(d.N - 1, use_graph=True)
    assert sw.last_kernel == "persistent"
    for k in (1, 7, d.N - 1):
        assert_stage_equal(sw.get_J(k), sw.get_idx(k), ora["J_all"][k - 1], ora["idx_all"][k - 1], f"persistent stage {k}")
    sw.close()
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 12, 10, 8, 7
    d4 = sp.channel_desc(0)
    o4 = oracle_lib.sweep(d4, n_stages=9)
    sw = bellman.Sweep(d4).run(9, use_graph=True)
    assert sw.last_kernel == "persistent"
    assert_stage_equal(sw.get_J(), sw.get_idx(), o4["J_last"], o4["idx_last"], "persistent D=4")
    sw.close()
    monkeypatch.setenv("BELLMAN_NO_PERSISTENT", "1")
    pos = bellman.tables.stack_problems(bellman.Solver_position()._axis_descs())
    op = oracle_lib.sweep(pos, n_stages=11)
    sw = bellman.Sweep(pos).run(11, use_graph=True)
    assert sw.last_kernel != "persistent"
    assert_stage_equal(sw.get_J(), sw.get_idx(), op["J_last"], op["idx_last"], "graph fallback")
    sw.close()


def test_attitude_reference_size(bellman, oracle_lib):
    """config 3 at the reference's own size: 3 axes x 1000x300 x 3 controls (Solver_attitude.m:106-144)."""
    sa = bellman.Solver_attitude()
    d = bellman.tables.stack_problems(sa._axis_descs())
    ora = oracle_lib.sweep(d, n_stages=12)
    for kernel in ("direct", "window"):
        sw = bellman.Sweep(d).run(12, kernel=KERNELS[kernel])
        assert sw.last_kernel.split(":")[0] == kernel
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "attitude " + kernel)
        sw.close()


@pytest.mark.parametrize("failure", [False, True])
def test_pos_att_reference_size(bellman, oracle_lib, failure):
    """config 5 at the reference's own size: 30x30x20x15 x 9 combos (Solver_pos_att.m:100-156);
    non-uniform grids (SEARCH locate) in three dimensions."""
    sp = bellman.Solver_pos_att()
    d = sp.channel_desc(0, failure=failure)
    assert d.C == (6 if failure else 9)
    ora = oracle_lib.sweep(d, n_stages=8)
    sw = bellman.Sweep(d).run(8)
    assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "pos-att")
    sw.close()


def test_pos_att_check_log_and_early_stop(bellman, oracle_lib):
    """Solver_pos_att.m:273-285: sums every `period` stages; a huge tolerance stops at the first check."""
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 12, 10, 8, 7
    d = sp.channel_desc(1)
    sw = bellman.Sweep(d).run(30, check_period=10, check_tol=0.0)
    log = sw.check_log()
    assert [int(x) for x in log[:, 0]] == [1990, 1980, 1970]
    J, idx = sw.get_J(), sw.get_idx()
    assert abs(log[-1, 1] - J.sum()) <= 1e-12 * abs(J.sum()) and log[-1, 2] == float((idx + 1).sum())
    sw2 = bellman.Sweep(d).run(30, check_period=10, check_tol=1e300)
    assert sw2.current_stage == 1990
    ora = oracle_lib.sweep(d, n_stages=30, check_period=10, check_tol=1e300)
    assert ora["stage"] == 1990
    assert_stage_equal(sw2.get_J(), sw2.get_idx(), ora["J_last"], ora["idx_last"], "early stop")
    sw.close(); sw2.close()


def test_random_terminal_cost_and_edge_shapes(bellman, oracle_lib):
    """set_J with a seeded random J_N (rough surface => every lerp branch, ties unlikely), smallest
    grids (2 points), C = 1, D = 3."""
    rng = np.random.default_rng(1)
    t = bellman.tables
    for (n0, n1, C) in ((2, 2, 1), (2, 37, 3), (33, 2, 7), (65, 31, 40)):
        obj = bellman.Dynamic_Solver()
        obj.dx, obj.du, obj.N = n0, C, 5
        d = obj._build()
        # make the two dims different sizes
        d.n = [n0, n1]
        s1 = t.linspace(-2.5, 3.0, n1)
        A = obj.A
        d.grid[1] = s1.reshape(1, -1); d.Tb[0] = (A[0, 1] * s1).reshape(1, -1); d.Tb[1] = (A[1, 1] * s1).reshape(1, -1)
        d.q[1] = (0.05 * (s1 * s1)).reshape(1, -1)
        d.validate()
        JN = rng.normal(size=(1, n0 * n1)) * 10
        for kernel in (1, 3):
            sw = bellman.Sweep(d)
            sw.set_J(JN)
            sw.run(2, kernel=kernel)
            ora = oracle_lib.sweep(d, n_stages=2, J_N=JN)
            assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"{n0}x{n1}x{C} k{kernel}")
            sw.close()
    # D = 3: drop the last dimension of a pos-att channel
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 9, 8, 7, 5
    d4 = sp.channel_desc(2)
    d3 = t.Desc(n=d4.n[:3], C=d4.C, N=6, grid=d4.grid[:3], src_a=[0, 1, 2], src_b=[1, -1, -1],
                Ta=d4.Ta[:3], Tb=[d4.Tb[0], None, None], Tc=[None, d4.Tc[1], d4.Tc[3]],
                q_order=[2, 0, 1], q=d4.q[:3], r=d4.r).validate()
    JN = rng.normal(size=(1, d3.S))
    sw = bellman.Sweep(d3); sw.set_J(JN); sw.run(3)
    ora = oracle_lib.sweep(d3, n_stages=3, J_N=JN)
    assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "D=3")
    sw.close()


def test_exact_ties_pick_first_index(bellman, oracle_lib):
    """Duplicate controls produce exact ties; MATLAB's min returns the first index."""
    obj = bellman.Dynamic_Solver()
    obj.dx, obj.du, obj.N = 24, 6, 4
    d = obj._build()
    for tab in (d.Tc[0], d.Tc[1], d.r):
        tab[0, 3:] = tab[0, :3]                       # controls 3..5 duplicate 0..2
    for kernel in (1, 3):
        sw = bellman.Sweep(d).run(3, kernel=kernel)
        idx = sw.get_idx()
        assert idx.max() <= 2
        ora = oracle_lib.sweep(d, n_stages=3)
        assert_stage_equal(sw.get_J(), idx, ora["J_last"], ora["idx_last"], "ties")
        sw.close()


@pytest.mark.parametrize("variant", ["wide", "ring"])
def test_kirk_scaled_full_size_spot_check(bellman, oracle_lib, monkeypatch, variant):
    """config 4 (8192 x 8192 x 512): one stage from a seeded smooth+rough J_N at full size, checked
    on 100k sampled states against the oracle's pointwise evaluator, plus size-independent
    properties: every argmin is a valid index and J_k = tot(argmin) exactly."""
    want = _window_variant(monkeypatch, variant)
    obj = bellman.Dynamic_Solver()
    obj.dx, obj.du, obj.N = 8192, 512, 200
    obj.store_J_star = False
    d = bellman.tables.kirk_desc(obj.A, obj.B, obj.Q, obj.R, obj.N, obj.x_min, obj.x_max, obj.dx, obj.u_min,
                                 obj.u_max, obj.du, store_J_all=False, store_idx_all=False)
    rng = np.random.default_rng(2)
    s = d.grid[0][0]
    JN = (0.3 * s[:, None] ** 2 + 0.1 * s[None, :] ** 2).ravel(order="F")
    JN += rng.normal(size=JN.shape) * 1e-3
    sw = bellman.Sweep(d)
    sw.set_J(JN.reshape(1, -1))
    sw.run(1)
    assert sw.last_kernel == want
    Jg, Ig = sw.get_J()[0], sw.get_idx()[0]
    assert Ig.min() >= 0 and Ig.max() < 512
    pts = rng.integers(0, d.S, size=100_000)
    pts[:4] = [0, 8191, d.S - 8192, d.S - 1]          # the four corners
    Jo, Io = oracle_lib.stage_points(d, JN, pts)
    assert np.array_equal(Ig[pts], Io) and np.array_equal(Jg[pts], Jo)
    sw.close()


def test_api_error_paths(bellman):
    """Status codes across the ABI: NOT_RUN, STATE, BAD_ARG — and that a failed call leaves the
    handle usable."""
    obj = bellman.Dynamic_Solver()
    obj.dx, obj.du, obj.N = 16, 8, 6
    d = bellman.tables.kirk_desc(obj.A, obj.B, obj.Q, obj.R, obj.N, obj.x_min, obj.x_max, obj.dx, obj.u_min,
                                 obj.u_max, obj.du, store_J_all=False, store_idx_all=False)
    sw = bellman.Sweep(d)
    assert sw.current_stage == 6
    with pytest.raises(bellman.BellmanError) as e:       # idx of the terminal stage does not exist
        sw.get_idx(6)
    assert e.value.code == -1
    sw.run(2)
    with pytest.raises(bellman.BellmanError) as e:       # stage 5 was computed but not stored
        sw.get_J(5)
    assert e.value.code == -5
    with pytest.raises(bellman.BellmanError) as e:       # stage 2 not computed yet
        sw.get_J(2)
    assert e.value.code == -5
    with pytest.raises(bellman.BellmanError) as e:       # would pass stage 1
        sw.run(10)
    assert e.value.code == -6
    with pytest.raises(bellman.BellmanError) as e:       # rollout needs store_idx_all
        sw.rollout(obj.A, obj.B, d.meta["U_mesh"], [[0.0, 0.0]])
    assert e.value.code == -6
    sw.run(3)                                            # still usable
    assert sw.current_stage == 1 and np.isfinite(sw.get_J()).all()
    sw.set_J(None)                                       # resume from a fresh terminal cost
    assert sw.current_stage == 6
    sw.close()


def test_resume_from_saved_stage(bellman, oracle_lib):
    """Checkpoint / resume: J of an intermediate stage fed back through set_J continues the sweep
    bit-identically (the state of a sweep is one J array)."""
    sp = bellman.Solver_position()
    d = bellman.tables.stack_problems(sp._axis_descs())
    a = bellman.Sweep(d).run(20)
    J_mid = a.get_J()
    a.run(15)
    b = bellman.Sweep(d)
    b.set_J(J_mid)                     # stage counter restarts at N; only the values matter (time-invariant tables)
    b.run(15)
    assert np.array_equal(a.get_J(), b.get_J()) and np.array_equal(a.get_idx(), b.get_idx())
    a.close(); b.close()


@pytest.mark.parametrize("variant", WINDOW_VARIANTS)
def test_window_kernel_multi_stage_medium(bellman, oracle_lib, monkeypatch, variant):
    """TMA-staged kernel over several stages on a grid with many tiles (1024 x 768 x 64): complete
    comparison of J and argmin with the oracle."""
    want = _window_variant(monkeypatch, variant)
    obj = bellman.Dynamic_Solver()
    t = bellman.tables
    n0, n1, C = 1024, 768, 64
    s0, s1, u = t.linspace(-2.5, 3.0, n0), t.linspace(-2.5, 3.0, n1), t.linspace(-40.0, 10.0, C)
    A, B = obj.A, obj.B.ravel()
    row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
    d = t.Desc(n=[n0, n1], C=C, N=5, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
               Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
               Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
               q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u)).validate()
    sw = bellman.Sweep(d).run(3)
    assert sw.last_kernel == want
    ora = oracle_lib.sweep(d, n_stages=3)
    assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "window 1024x768x64")
    sw.close()


def test_attitude_x4_full_size_spot_check(bellman, oracle_lib):
    """config 3 (reference grid refined 4x per dimension: 3 axes x 4000 x 1200 x 3 controls): two
    stages at full size with the lean CHAIN kernel, 60k sampled states per axis against the oracle's
    pointwise evaluator (the second stage is seeded with the GPU's own first-stage J)."""
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 4000, 1200
    d = bellman.tables.stack_problems(sa._axis_descs())
    rng = np.random.default_rng(7)
    JN = rng.normal(size=(3, d.S)) * 0.1
    sw = bellman.Sweep(d)
    sw.set_J(JN)
    sw.run(1)
    assert sw.last_kernel.split(":")[0] == "window"
    J1, I1 = sw.get_J(), sw.get_idx()
    sw.run(1)
    J2, I2 = sw.get_J(), sw.get_idx()
    for p in range(3):
        pts = rng.integers(0, d.S, size=60_000)
        pts[:4] = [0, 3999, d.S - 4000, d.S - 1]
        Jo, Io = oracle_lib.stage_points(d, JN[p], pts, p=p)
        assert np.array_equal(I1[p][pts], Io) and np.array_equal(J1[p][pts], Jo)
        Jo, Io = oracle_lib.stage_points(d, J1[p], pts, p=p)
        assert np.array_equal(I2[p][pts], Io) and np.array_equal(J2[p][pts], Jo)
    sw.close()


# ---------------------------------------------------------------------------------------------
# "next" rows: consumers of the sweep output (policy lookup, simplified-plant rollout, controller
# save / set_controller round trip, resume from a saved stage)
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_policy_lookup_matches_oracle(bellman, oracle_lib):
    rng = np.random.default_rng(11)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 9, 8, 7, 6
    d = sp.channel_desc(1)
    with bellman.Sweep(d) as sw:
        sw.run(6)
        idx = sw.get_idx()[0]
        lo = np.array([d.grid[k][0][0] for k in range(4)]); hi = np.array([d.grid[k][0][-1] for k in range(4)])
        x = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=(4096, 4))
        x[:len(d.grid[3][0])] = np.stack([np.resize(d.grid[k][0], len(d.grid[3][0])) for k in range(4)], 1)  # nodes
        x[100] = [0.5 * (d.grid[k][0][1] + d.grid[k][0][2]) for k in range(4)]                                  # midpoints
        got = sw.policy_lookup(x)
    want = oracle_lib.policy_lookup(d, idx, x)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["attitude", "position"])
def test_rollout_axis_matches_oracle(bellman, oracle_lib, which):
    t = bellman.tables
    rng = np.random.default_rng(5)
    N, h = 60, 0.01
    if which == "attitude":
        d = t.attitude_axis_desc(t.deg2rad(-0.7), t.deg2rad(0.7), 24, -5.0, 5.0, 20, [-0.01, 0.0, 0.01], 2.5,
                                 6.0, 6.0, 0.1, h, N)
        rate = 0
    else:
        d = t.position_axis_desc(-0.5, 0.5, 21, -0.5, 0.5, 17, [-0.26, 0.0, 0.26], 4.16, 6.0, 6.0, 0.1, h, N)
        rate = 1
    u_inc = d.Tc[rate][0]
    lo = np.array([d.grid[k][0][0] for k in range(2)]); hi = np.array([d.grid[k][0][-1] for k in range(2)])
    x0 = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), size=(257, 2))
    d.store_idx_all = True
    with bellman.Sweep(d) as sw:
        sw.run()
        idx_all = np.stack([sw.get_idx(k)[0] for k in range(1, N)])
        Xg, Cg = sw.rollout_axis(u_inc, x0, N - 1, h, rate, time_varying=True)
        Xf, Cf = sw.rollout_axis(u_inc, x0, 150, h, rate, stage=1)
    Xo, Co = oracle_lib.rollout_axis(d, idx_all, u_inc, x0, N - 1, h, rate, time_varying=True)
    assert np.array_equal(Cg, Co) and np.array_equal(Xg, Xo)
    Xo, Co = oracle_lib.rollout_axis(d, idx_all[0], u_inc, x0, 150, h, rate)
    assert np.array_equal(Cf, Co) and np.array_equal(Xf, Xo)


@pytest.mark.gpu
def test_axis_solver_simplified_closed_loop(bellman, oracle_lib):
    """facade: Solver_attitude.simplified_run then the test_simplified.m loop on the GPU vs the literal loop."""
    from oracle import matlab_literal as ml
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t, sa.T_final, sa.h = 40, 30, 0.5, 0.01
    sa.simplified_run()
    X0 = np.array([[[0.05, 0.1], [-0.2, -0.05], [0.0, 0.3]]])
    X, U = sa.get_optimal_path_simplified(X0, n_steps=120)
    d = sa._desc
    for p in range(3):
        J = (sa.J1, sa.J2, sa.J3)[p]
        Uopt = np.asarray(sa.U_vector)[sa.U_idx[p] - 1]
        Xl, Ul = ml.simplified_axis_rollout_literal(d.grid[0][p], d.grid[1][p], Uopt, 0, lambda u: u / J, sa.h, X0[0, p], 120)
        assert np.array_equal(U[0, p], Ul)
        assert np.array_equal(X[0, p], Xl)


@pytest.mark.gpu
def test_pos_att_controller_file_round_trip(bellman, oracle_lib, tmp_path):
    rng = np.random.default_rng(2)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 8, 7, 6, 5
    f = str(tmp_path / "channel_y_controller_1.mat")
    ctl = sp.calculate_one_channel_U_Opt(1, file_name=f, n_stages=7)
    back = sp.load_controller(f)
    assert np.array_equal(back["U_Optimal_id"], ctl["U_Optimal_id"])
    assert np.array_equal(back["F_gI_Values"], ctl["F_gI_Values"])
    for k in range(4):
        assert np.array_equal(back["GridVectors"][k], ctl["GridVectors"][k])
    sp.set_controller(f, "y")
    with pytest.raises(ValueError):
        sp.set_controller(f, "w")
    d = sp.channel_desc(1)
    lo = np.array([d.grid[k][0][0] for k in range(4)]); hi = np.array([d.grid[k][0][-1] for k in range(4)])
    x = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), size=(300, 4))
    F = sp.thruster_lookup_batch("y", x)                                    # GPU, after set_stage
    c = oracle_lib.policy_lookup(d, (ctl["U_Optimal_id"] - 1).ravel(order="F"), x)
    for j, (name, thr) in enumerate(zip(("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"), (2, 3, 8, 9))):
        assert np.array_equal(F[:, j], ctl[name][c])
        host = getattr(sp, "Opt_F_Thr%d" % thr)(x[:, 0], x[:, 1], x[:, 2], x[:, 3])      # host NearestPolicy
        assert np.array_equal(host, F[:, j])


@pytest.mark.gpu
def test_set_stage_resume_and_errors(bellman, oracle_lib):
    """bellman_set_stage: J of stage k put back on a fresh handle continues to the same bits."""
    d = bellman.tables.kirk_desc([[0.9974, 0.0539], [-0.1078, 1.1591]], [0.0013, 0.0539], [[0.25, 0.0], [0.0, 0.05]],
                                 0.05, 20, -2.5, 3.0, 48, -40.0, 10.0, 40, store_J_all=False, store_idx_all=False)
    with bellman.Sweep(d) as a:
        a.run(8)
        Jk, k = a.get_J(), a.current_stage
        a.run(6)
        Jend, iend = a.get_J(), a.get_idx()
    with bellman.Sweep(d) as b:
        b.set_stage(k, Jk)
        assert b.current_stage == k
        b.run(6)
        assert np.array_equal(b.get_J(), Jend) and np.array_equal(b.get_idx(), iend)
        with pytest.raises(bellman.BellmanError):
            b.set_stage(d.N, None, np.zeros(d.S, dtype=np.int32))
        bad = np.zeros(d.S, dtype=np.int32)
        bad[7] = d.C                                                    # one index past the control grid
        with pytest.raises(bellman.BellmanError):
            b.set_stage(3, None, bad)
        assert b.current_stage == 6                                     # a refused call leaves the handle as it was


@pytest.mark.gpu
def test_dynamic_solver_archive_and_compare_data(bellman, golden, tmp_path):
    """save -> load -> compare_data (Dynamic_Solver.m:266-280) on the golden configuration."""
    def make():
        o = bellman.Dynamic_Solver()
        for k in ("A", "B", "Q", "R", "N", "x_min", "x_max", "dx", "u_min", "u_max", "du"):
            setattr(o, k, golden[k])
        return o
    a = make().run()
    f = str(tmp_path / "run_a.mat")
    a.save(f)
    b = bellman.Dynamic_Solver.load(f)
    assert bellman.Dynamic_Solver.compare_data(a, b)
    assert np.array_equal(b.u_star, a.u_star) and np.array_equal(b.s_r, a.s_r) and b.N == a.N
    c = make().run()                                             # a second run reproduces the archive bit for bit
    assert bellman.Dynamic_Solver.compare_data(c, b)
    c.J_star[3, 4, 5] = np.nextafter(c.J_star[3, 4, 5], np.inf)
    assert not bellman.Dynamic_Solver.compare_data(c, b)
    with pytest.raises(ValueError):
        bellman.Dynamic_Solver.compare_data(bellman.Dynamic_Solver(), b)
    # the archive agrees with the reference's own saved run: u_star exact, J_star within 1e-12
    assert np.array_equal(b.u_star[:, :, :golden["N"] - 1], golden["u_star"][:, :, :golden["N"] - 1])


@pytest.mark.parametrize("shape", [(77, 45, 3), (130, 37, 2), (64, 64, 4), (33, 9, 3), (200, 150, 1)])
def test_strip_kernel_ragged_random(bellman, oracle_lib, shape):
    """k_stage_strip (attitude-type problems, <= 4 controls): ragged tiles, clamped edges, rough J,
    three stacked axes with different ranges, several stages."""
    t = bellman.tables
    rng = np.random.default_rng(shape[0])
    n_w, n_t, C = shape
    U = np.linspace(-0.11, 0.11, C) if C > 1 else np.array([0.05])
    descs = [t.attitude_axis_desc(-0.9, 0.9, n_w, -ang, ang, n_t, U, J, 6.0, 6.0, 4.0, 0.02, 6)
             for ang, J in ((30.0, 0.0285), (20.0, 0.0283), (35.0, 0.0245))]
    d = t.stack_problems(descs)
    JN = rng.normal(size=(3, d.S)) * 3
    ora = oracle_lib.sweep(d, n_stages=4, J_N=JN)
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(4, kernel=KERNELS["window"])
        assert sw.last_kernel == "window:strip"
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"strip {shape}")


FAMILIES = ["stream", "tile"]   # k_stage_stream (default for Solver_pos_att's structure) / k_stage_tile_pa


def pick_family(monkeypatch, family):
    if family == "tile":
        monkeypatch.setenv("BELLMAN_NO_STREAM", "1")


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("failure", [False, True])
def test_tile_kernel_pos_att_reference_size(bellman, oracle_lib, monkeypatch, failure, family):
    """the TMA-staged D = 4 kernels on config 5 at the reference's own size (30x30x20x15 x 9; 6 controls
    in failure mode): the streaming factorised kernel and the one-box-per-tile kernel."""
    pick_family(monkeypatch, family)
    sp = bellman.Solver_pos_att()
    d = sp.channel_desc(0, failure=failure)
    ora = oracle_lib.sweep(d, n_stages=6)
    with bellman.Sweep(d) as sw:
        sw.run(6, kernel=KERNELS["tile"])
        assert sw.last_kernel == family
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "pos-att " + family)


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("mesh", [(34, 9, 7, 5), (10, 8, 6, 4), (66, 3, 2, 9), (64, 16, 8, 8)])
def test_tile_kernel_ragged_random(bellman, oracle_lib, monkeypatch, mesh, family):
    """ragged tiles in every dimension, clamped edges, rough terminal cost, all three channels, D = 4 and D = 3."""
    pick_family(monkeypatch, family)
    t = bellman.tables
    rng = np.random.default_rng(mesh[0])
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = mesh
    for ch in range(3):
        d = sp.channel_desc(ch)
        JN = rng.normal(size=(1, d.S)) * 2
        ora = oracle_lib.sweep(d, n_stages=3, J_N=JN)
        with bellman.Sweep(d) as sw:
            sw.set_J(JN)
            sw.run(3, kernel=KERNELS["tile"])
            assert sw.last_kernel == family
            assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"{family} {mesh} ch{ch}")
    d4 = sp.channel_desc(2)
    d3 = t.Desc(n=d4.n[:3], C=d4.C, N=6, grid=d4.grid[:3], src_a=[0, 1, 2], src_b=[1, -1, -1],
                Ta=d4.Ta[:3], Tb=[d4.Tb[0], None, None], Tc=[None, d4.Tc[1], d4.Tc[3]],
                q_order=[2, 0, 1], q=d4.q[:3], r=d4.r).validate()
    JN = rng.normal(size=(1, d3.S))
    ora = oracle_lib.sweep(d3, n_stages=3, J_N=JN)
    with bellman.Sweep(d3) as sw:
        sw.set_J(JN)
        sw.run(3, kernel=KERNELS["tile"])
        assert sw.last_kernel == "tile"
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"tile D=3 {mesh}")


def test_tile_kernel_falls_back_when_not_a_stencil(bellman, oracle_lib):
    """odd leading dimension (TMA stride rule) -> the request for the staged kernel runs the direct kernel."""
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 9, 8, 7, 5
    d = sp.channel_desc(1)
    ora = oracle_lib.sweep(d, n_stages=2)
    with bellman.Sweep(d) as sw:
        sw.run(2, kernel=KERNELS["tile"])
        assert sw.last_kernel == "direct"
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "tile fallback")


@pytest.mark.parametrize("env", [{"BELLMAN_STRIP_R": "16"}, {"BELLMAN_STRIP_R": "4", "BELLMAN_STRIP_NW": "8"},
                                 {"BELLMAN_STRIP_R": "12", "BELLMAN_STRIP_NW": "2", "BELLMAN_STRIP_PF": "5"},
                                 {"BELLMAN_WIN_OCC": "3"},      # 5-CTA register budget: 2-ahead table prefetch, conversion-free locate
                                 {"BELLMAN_WIN_OCC": "3", "BELLMAN_STRIP_R": "6"},
                                 {"BELLMAN_WIN_NOSTRIP": "1"}])
def test_strip_kernel_geometries(bellman, oracle_lib, monkeypatch, env):
    """every strip geometry the planner can pick (strip length 16 is the default on large grids), the
    L2 prefetch with a distance that stays inside a small grid, and the k_stage_chain fallback."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    t = bellman.tables
    rng = np.random.default_rng(9)
    U = np.array([-0.11, 0.0, 0.11])
    descs = [t.attitude_axis_desc(-0.9, 0.9, 150, -ang, ang, 100, U, J, 6.0, 6.0, 4.0, 0.02, 6)
             for ang, J in ((30.0, 0.0285), (20.0, 0.0283), (35.0, 0.0245))]
    d = t.stack_problems(descs)
    JN = rng.normal(size=(3, d.S)) * 3
    ora = oracle_lib.sweep(d, n_stages=3, J_N=JN)
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(3, kernel=KERNELS["window"])
        assert sw.last_kernel == ("window:chain" if "BELLMAN_WIN_NOSTRIP" in env else "window:strip")
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"strip {env}")


@pytest.mark.parametrize("env", [{"BELLMAN_TILE_NT": "256"}, {"BELLMAN_TILE_GENERIC": "1"}, {"BELLMAN_TILE": "2,2,4"},
                                 {"BELLMAN_TILE_PF": "3"}, {"BELLMAN_TILE": "1,8,2", "BELLMAN_TILE_NT": "256"}])
def test_tile_kernel_variants(bellman, oracle_lib, monkeypatch, env):
    """the 256-thread form, the generic (run-time structure) tile kernel, forced tile shapes and the L2 prefetch."""
    monkeypatch.setenv("BELLMAN_NO_STREAM", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(4)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 40, 12, 10, 14
    d = bellman.tables.stack_problems([sp.channel_desc(c) for c in range(3)])
    JN = rng.normal(size=(3, d.S)) * 2
    ora = oracle_lib.sweep(d, n_stages=3, J_N=JN)
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(3, kernel=KERNELS["tile"])
        assert sw.last_kernel == "tile"
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"tile {env}")


@pytest.mark.parametrize("env", [{"BELLMAN_STREAM": "1,2,0,2"}, {"BELLMAN_STREAM": "2,3,8,4"}, {"BELLMAN_STREAM": "4,4,0,3"},
                                 {"BELLMAN_STREAM": "2,8,4,3"}, {"BELLMAN_STREAM": "2,10,0,2"},
                                 {"BELLMAN_STREAM": "2,4,0,3,2"}, {"BELLMAN_STREAM": "2,8,0,3,4"}, {"BELLMAN_STREAM": "1,3,8,2,1"}])
def test_stream_kernel_variants(bellman, oracle_lib, monkeypatch, env):
    """k_stage_stream with forced geometry "T1,T2,T3,NJ[,NP]": 32x1 / 16x2 / 8x4 column patches, 2..10 consumer
    warps, the walk cut into chunks (T3, a multiple of the ring length), slab rings of 2..4 boxes, warp
    specialisation (NP dedicated producer warps); rough
    terminal cost, three channels with different grids, several stages (ping-pong slots)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(4)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 40, 12, 10, 14
    d = bellman.tables.stack_problems([sp.channel_desc(c) for c in range(3)])
    JN = rng.normal(size=(3, d.S)) * 2
    ora = oracle_lib.sweep(d, n_stages=3, J_N=JN)
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(3, kernel=KERNELS["tile"])
        assert sw.last_kernel == "stream"
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], f"stream {env}")


@pytest.mark.parametrize("family", FAMILIES)
def test_pos_att_x4_full_size_spot_check(bellman, oracle_lib, monkeypatch, family):
    """the bench's pos-att workload at full size (one channel of 120 x 120 x 80 x 60 x 9; the thrusters
    move w by up to +-5 cells per stage here, so the tile kernel's box is much wider than in the small
    tests): two stages with the tile kernel, 60k sampled states against the oracle's pointwise evaluator."""
    pick_family(monkeypatch, family)
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 120, 120, 80, 60
    d = sp.channel_desc(0)
    lo, hi = bellman.query_stencil(d)
    assert hi[3] - lo[3] >= 6                      # the wide stencil this test is about
    rng = np.random.default_rng(11)
    JN = rng.normal(size=(1, d.S)) * 0.1
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(1)
        assert sw.last_kernel == family
        J1, I1 = sw.get_J(), sw.get_idx()
        sw.run(1)
        J2, I2 = sw.get_J(), sw.get_idx()
    pts = rng.integers(0, d.S, size=60_000)
    pts[:4] = [0, 119, d.S - 120, d.S - 1]
    Jo, Io = oracle_lib.stage_points(d, JN[0], pts)
    assert np.array_equal(I1[0][pts], Io) and np.array_equal(J1[0][pts], Jo)
    Jo, Io = oracle_lib.stage_points(d, J1[0], pts)
    assert np.array_equal(I2[0][pts], Io) and np.array_equal(J2[0][pts], Jo)


@pytest.mark.gpu
def test_stage_host_pipelined_equals_sequence_and_oracle(bellman, oracle_lib, monkeypatch):
    """bellman_stage_host: the slab-pipelined path (k_stage_wide, 8 slabs of tiles, copies on side streams)
    gives bit for bit what set_J + run(1) + get_J + get_idx give, and what the oracle gives; ragged tile
    counts, a rough J, continuing from the device's J (J_next = None), and the fallback for kernels that
    cannot run a tile range."""
    obj = bellman.Dynamic_Solver()
    t = bellman.tables
    rng = np.random.default_rng(11)
    monkeypatch.setenv("BELLMAN_HOST_SLABS", "8")       # (the default adapts to the grid: 2 slabs at these sizes)
    for n0, n1, C in ((256, 1024, 48), (130, 777, 33)):
        s0, s1, u = t.linspace(-2.5, 3.0, n0), t.linspace(-2.5, 3.0, n1), t.linspace(-40.0, 10.0, C)
        A, B = obj.A, obj.B.ravel()
        row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
        d = t.Desc(n=[n0, n1], C=C, N=6, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
                   Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
                   Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
                   q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u),
                   store_J_all=False, store_idx_all=False).validate()
        ntile1 = -(-n1 // 64)
        n_slabs = -(-ntile1 // -(-ntile1 // 8))
        JN = rng.normal(size=(1, d.S)) * 3
        ora = oracle_lib.sweep(d, n_stages=2, J_N=JN, keep_all=True)
        with bellman.Sweep(d) as sw:
            J1, I1 = sw.stage_host(JN, kernel=KERNELS["window"])
            assert sw.last_kernel == "window:wide" and sw.current_stage == d.N - 1
            assert sw.stats()["launches"] == n_slabs   # really went slab by slab
            assert_stage_equal(J1, I1, ora["J_all"][d.N - 2], ora["idx_all"][d.N - 2], "stage_host first stage")
            J2, I2 = sw.stage_host(None, kernel=KERNELS["window"])      # continue from the device's J
            assert sw.current_stage == d.N - 2
            assert_stage_equal(J2, I2, ora["J_all"][d.N - 3], ora["idx_all"][d.N - 3], "stage_host second stage")
            sw.set_J(JN); sw.run(1, kernel=KERNELS["window"])
            assert np.array_equal(sw.get_J(), J1) and np.array_equal(sw.get_idx(), I1)
    # fallback: strip kernel (attitude) and a forced direct kernel go through the plain sequence
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 400, 120
    da = t.stack_problems(sa._axis_descs())
    oa = oracle_lib.sweep(da, n_stages=1)
    with bellman.Sweep(da) as sw:
        J, I = sw.stage_host(np.zeros((da.P, da.S)))
        assert_stage_equal(J, I, oa["J_last"], oa["idx_last"], "stage_host fallback")
    monkeypatch.delenv("BELLMAN_HOST_SLABS")
    with bellman.Sweep(d) as sw:                        # default slab count on this small grid: 2
        J, I = sw.stage_host(JN, kernel=KERNELS["window"])
        assert sw.stats()["launches"] == 2
        assert np.array_equal(J, J1) and np.array_equal(I, I1)
    monkeypatch.setenv("BELLMAN_NO_HOST_PIPELINE", "1")
    with bellman.Sweep(d) as sw:
        J, I = sw.stage_host(JN, kernel=KERNELS["window"])
        assert sw.stats()["launches"] == 1
        assert np.array_equal(J, J1) and np.array_equal(I, I1)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["wide", "ring"])
def test_window_kernel_stacked_problems(bellman, oracle_lib, monkeypatch, variant):
    """Three different linear systems stacked into one descriptor (P = 3, per-problem control tables in
    k_stage_wide's constant-bank parameter, per-problem tile extrema and chunk bounds), rough terminal cost."""
    want = _window_variant(monkeypatch, variant)
    obj = bellman.Dynamic_Solver()
    t = bellman.tables
    descs = []
    for k, scale in enumerate((1.0, 0.7, 1.3)):
        A = np.array(obj.A, dtype=np.float64).reshape(2, 2) * np.array([[1.0, scale], [scale, 1.0]])
        B = np.array(obj.B, dtype=np.float64).ravel() * (1.0 + 0.5 * k)
        descs.append(t.kirk_desc(A, B, obj.Q, obj.R, 5, obj.x_min, obj.x_max, 200, obj.u_min, obj.u_max, 40,
                                 store_J_all=False, store_idx_all=False))
    d = t.stack_problems(descs)
    JN = np.random.default_rng(3).normal(size=(d.P, d.S)) * 2
    ora = oracle_lib.sweep(d, n_stages=3, J_N=JN)
    with bellman.Sweep(d) as sw:
        sw.set_J(JN)
        sw.run(3, kernel=KERNELS["window"])
        assert sw.last_kernel == want
        assert_stage_equal(sw.get_J(), sw.get_idx(), ora["J_last"], ora["idx_last"], "stacked " + variant)
