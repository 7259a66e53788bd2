"""The separable 1-D tables must reproduce the reference's S x C arrays bit for bit
(oracle/matlab_literal.py materialises those arrays exactly as the .m files do)."""
import numpy as np

from oracle import matlab_literal as ml


def full_next(d, p, dim):
    """x'_dim over the full S x C array, formed with the kernel's association (Ta + Tb) + Tc."""
    D = d.D
    shape = list(d.n) + [d.C]

    def bc(vec, axis):
        sh = [1] * (D + 1)
        sh[axis] = len(vec)
        return np.asarray(vec).reshape(sh)

    x = bc(d.Ta[dim][p], d.src_a[dim])
    if d.Tb[dim] is not None:
        x = x + bc(d.Tb[dim][p], d.src_b[dim])
    if d.Tc[dim] is not None:
        x = x + bc(d.Tc[dim][p], D)
    return np.broadcast_to(x, shape)


def full_cost(d, p):
    D = d.D

    def bc(vec, axis):
        sh = [1] * (D + 1)
        sh[axis] = len(vec)
        return np.asarray(vec).reshape(sh)

    g = bc(d.q[d.q_order[0]][p], d.q_order[0])
    for m in range(1, D):
        g = g + bc(d.q[d.q_order[m]][p], d.q_order[m])
    return np.broadcast_to(g + bc(d.r[p], D), list(d.n) + [d.C])


def test_kirk_tables(bellman):
    L = ml.DynamicSolverLiteral(N=20, dx=37, du=53)
    L.setup()
    d = bellman.tables.kirk_desc(L.A, L.B, L.Q, L.R, 20, L.x_min, L.x_max, 37, L.u_min, L.u_max, 53)
    assert np.array_equal(full_next(d, 0, 0), L.X_next_M1)
    assert np.array_equal(full_next(d, 0, 1), L.X_next_M2)
    assert np.array_equal(full_cost(d, 0), L.J_current_state)


def test_position_tables(bellman):
    sp = bellman.Solver_position()
    L = ml.SolverPositionLiteral()
    assert sp.N_stage == L.N_stage == 6000
    descs = sp._axis_descs()
    assert descs[0].n == [201, 201]
    for a in range(3):
        grids, nxt, Jc = L.axis_arrays(a)
        d = descs[a]
        assert np.array_equal(d.grid[0][0], grids[0]) and np.array_equal(d.grid[1][0], grids[1])
        assert np.array_equal(full_next(d, 0, 0), nxt[0])
        assert np.array_equal(full_next(d, 0, 1), nxt[1])
        assert np.array_equal(full_cost(d, 0), Jc)


def test_attitude_tables(bellman):
    sa = bellman.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 120, 45
    L = ml.SolverAttitudeLiteral(n_mesh_w=120, n_mesh_t=45)
    assert sa.N_stage == L.N_stage == 6000
    assert (sa.J1, sa.J2, sa.J3) == tuple(L.J)
    descs = sa._axis_descs()
    for a in range(3):
        grids, nxt, Jc = L.axis_arrays(a)
        d = descs[a]
        assert np.array_equal(d.grid[0][0], grids[0]) and np.array_equal(d.grid[1][0], grids[1])
        assert np.array_equal(full_next(d, 0, 0), nxt[0])
        assert np.array_equal(full_next(d, 0, 1), nxt[1])
        assert np.array_equal(full_cost(d, 0), Jc)


def test_pos_att_tables(bellman):
    sp = bellman.Solver_pos_att()
    L = ml.SolverPosAttLiteral()
    assert sp.N_stage == L.N_stage == 2000
    for ch in range(3):
        d = sp.channel_desc(ch)
        grids, nxt, Jc, combos = L.channel_arrays(ch)
        assert d.n == [30, 30, 20, 15] and d.C == 9
        for k in range(4):
            assert np.array_equal(d.grid[k][0], grids[k])
            assert np.array_equal(full_next(d, 0, k), nxt[k])
        assert np.array_equal(full_cost(d, 0), Jc)
        for nm, cc in zip(("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb"), combos):
            assert np.array_equal(d.meta[nm], cc)


def test_grid_known_answers(bellman):
    """grid KATs."""
    t = bellman.tables
    assert len(t.sym_linspace_position(-0.5, 0.5, 200)) == 201
    for n, (neg, pos) in {30: (16, 14), 20: (11, 9), 15: (8, 7)}.items():
        v = t.sym_linspace_pos_att(-1.0, 1.0, n)
        assert len(v) == n and np.sum(v <= 0) == neg and np.sum(v > 0) == pos
    x = t.sym_linspace_pos_att(-0.2, 0.2, 30)
    dd = np.diff(x)
    assert abs(dd[0] - 0.2 / 15) < 1e-15 and abs(dd[-1] - 0.2 / 14) < 1e-15   # two spacings
    T = 0.13
    c = t.vectors_allcomb([0, T], [0, T], [0, -T], [0, -T])
    combos = list(zip(*[list(v) for v in c]))
    assert combos == [(0, 0, 0, 0), (T, 0, 0, 0), (0, T, 0, 0), (T, T, 0, 0), (0, 0, -T, 0), (0, T, -T, 0),
                      (0, 0, 0, -T), (T, 0, 0, -T), (0, 0, -T, -T)]
    # failure mode (f0 = [0]): 6 combinations survive
    assert len(t.vectors_allcomb([0], [0, T], [0, -T], [0, -T])[0]) == 6
