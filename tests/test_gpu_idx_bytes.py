"""uint8 / uint16 device storage of the argmin (bellman_desc.idx_bytes): every stage kernel that writes
it and every consumer that reads it, against the oracle.  The ABI keeps taking and returning int32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(bellman):
    o = bellman.Dynamic_Solver()
    kirk = bellman.tables.kirk_desc(o.A, o.B, o.Q, o.R, 8, o.x_min, o.x_max, 96, o.u_min, o.u_max, 40,
                                    store_J_all=False, store_idx_all=False)
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 128, 64
    att = bellman.tables.stack_problems(sa._axis_descs())
    sp = bellman.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 20, 6, 6, 9
    pa = sp.channel_desc(0)
    return kirk, att, pa


@pytest.mark.parametrize("idx_bytes", [1, 2])
def test_every_stage_kernel_with_narrow_idx(bellman, oracle_lib, monkeypatch, idx_bytes):
    kirk, att, pa = _cases(bellman)
    runs = [("direct", kirk, dict(kernel=1), "direct"), ("splitc", kirk, dict(kernel=3), "splitc"),
            ("persistent", kirk, dict(use_graph=True), "persistent"), ("wide", kirk, dict(kernel=2), "window:wide"),
            ("strip", att, dict(kernel=2), "window:strip"), ("stream", pa, dict(kernel=4), "stream")]
    for name, d, kw, want in runs:
        ora = oracle_lib.sweep(d, n_stages=4)
        with bellman.Sweep(d, idx_bytes=idx_bytes) as sw:
            sw.run(4, **kw)
            assert sw.last_kernel == want, (name, sw.last_kernel)
            assert np.array_equal(sw.get_idx(), ora["idx_last"]), name
            assert np.array_equal(sw.get_J(), ora["J_last"]), name
            pts = np.arange(0, d.S, 97)
            Jp, Ip = sw.get_points(pts, prob=d.P - 1)
            assert np.array_equal(Ip, ora["idx_last"][d.P - 1][pts]) and np.array_equal(Jp, ora["J_last"][d.P - 1][pts])
    monkeypatch.setenv("BELLMAN_NO_STREAM", "1")
    monkeypatch.setenv("BELLMAN_WIN_NOSTRIP", "1")
    monkeypatch.setenv("BELLMAN_NO_WIDE", "1")
    for name, d, want in (("tile", pa, "tile"), ("chain", att, "window:chain"), ("ring", kirk, "window:ring")):
        ora = oracle_lib.sweep(d, n_stages=3)
        with bellman.Sweep(d, idx_bytes=idx_bytes) as sw:
            sw.run(3, kernel=2)
            assert sw.last_kernel == want
            assert np.array_equal(sw.get_idx(), ora["idx_last"]) and np.array_equal(sw.get_J(), ora["J_last"]), name


@pytest.mark.parametrize("idx_bytes", [1, 2])
def test_consumers_with_narrow_idx(bellman, oracle_lib, golden, idx_bytes):
    """rollout on the stored per-stage policies, nearest lookup, simplified rollout, set_stage round trip,
    Sigma-check sums, single-process group — all reading 1- or 2-byte indices."""
    d = bellman.tables.kirk_desc(golden["A"], golden["B"], golden["Q"], golden["R"], golden["N"], golden["x_min"],
                                 golden["x_max"], golden["dx"], golden["u_min"], golden["u_max"], golden["du"])
    ora = oracle_lib.sweep(d, keep_all=True)
    sw = bellman.Sweep(d, idx_bytes=idx_bytes).run()
    for k in (1, 40, d.N - 1):
        assert np.array_equal(sw.get_idx(k), ora["idx_all"][k - 1])
    x0 = np.random.default_rng(0).uniform(-2.6, 3.1, size=(64, 2))
    Xg, Ug = sw.rollout(golden["A"], golden["B"], d.meta["U_mesh"], x0)
    Xo, Uo = oracle_lib.rollout(d, ora["idx_all"][:, 0, :], golden["A"], golden["B"], d.meta["U_mesh"], x0)
    np.testing.assert_array_equal(Ug, Uo)
    np.testing.assert_array_equal(Xg, Xo)
    sw.close()
    kirk, att, pa = _cases(bellman)
    oa = oracle_lib.sweep(att, n_stages=5)
    with bellman.Sweep(att, idx_bytes=idx_bytes) as s2:
        s2.run(5)
        xq = np.random.default_rng(1).uniform(-0.9, 0.9, size=(200, 2)) * [1.0, 0.5]
        for p in range(3):
            assert np.array_equal(s2.policy_lookup(xq, prob=p), oracle_lib.policy_lookup(att, oa["idx_last"][p], xq, p=p))
        Xa, Ca = s2.rollout_axis(att.Tc[0][1], xq[:32], 20, 0.005, 0, prob=1)
        Xo, Co = oracle_lib.rollout_axis(att, oa["idx_last"][1], att.Tc[0][1], xq[:32], 20, 0.005, 0, p=1)
        assert np.array_equal(Ca, Co) and np.array_equal(Xa, Xo)
        J, I = s2.get_J(), s2.get_idx()
    with bellman.Sweep(att, idx_bytes=idx_bytes) as s3:          # resume from the saved stage
        s3.set_stage(att.N - 5, J, I)
        assert np.array_equal(s3.get_idx(), I)
        s3.run(2)
        o7 = oracle_lib.sweep(att, n_stages=7)
        assert np.array_equal(s3.get_idx(), o7["idx_last"]) and np.array_equal(s3.get_J(), o7["J_last"])
    op = oracle_lib.sweep(pa, n_stages=60)
    grp = bellman.SweepGroup(pa, [0, 0], idx_bytes=idx_bytes)
    grp.run(60, check_period=50, check_tol=0.0)
    assert np.array_equal(grp.get_idx(), op["idx_last"]) and np.array_equal(grp.get_J(), op["J_last"])
    lg = grp.check_log()
    assert lg.shape == (1, 3) and lg[0, 0] == pa.N - 50
    grp.close()
    with pytest.raises(bellman.BellmanError):
        big = bellman.Dynamic_Solver()._build()                  # 1000 controls do not fit one byte
        bellman.Sweep(big, idx_bytes=1)
