"""Shared pieces of the tests that replay the reference's own test scripts (test/test_u_star_M.m,
test/test_griddedInterp.m)."""
import numpy as np

# test/test_u_star_M.m:8-16 — D.get_optimal_path([2;1],'ssu',k) with the author's verdicts as comments
SSU_STAGES = (190, 180, 135, 128, 127, 30)          # "bad", "controllable but inefficient", "near optimal ...", ...
# test/test_griddedInterp.m:37,41 — the points the script drops on the interpolated surface
QUERY_POINTS = [(0.0, 0.0)] + list(zip([7, 5, 3, 3, 1, -4, -5], [10, 8, 5, 3, -1, -2, -5]))


def closed_loop_cost(obj, X, U):
    """sum_k x_k' Q x_k + R u_k^2 of one rollout; X [N, 2], U [N]."""
    return float(np.sum(obj.Q[0, 0] * X[:, 0] ** 2 + obj.Q[1, 1] * X[:, 1] ** 2) + np.sum(obj.R * U ** 2))


def check_u_star_M_verdicts(obj, rollouts):
    """rollouts: {('ssu', k) | ('Nssu', 1): (X [N, 2], U [N])}.  The comments of test_u_star_M.m as facts."""
    with np.errstate(over="ignore", invalid="ignore"):
        cost = {k: closed_loop_cost(obj, X, U) for k, (X, U) in rollouts.items()}
    assert not np.isfinite(cost[("ssu", 190)])                                  # "bad": the loop diverges
    assert np.isfinite(cost[("ssu", 180)]) and cost[("ssu", 180)] > 1.3 * cost[("ssu", 135)]   # "controllable but inefficient"
    ref = cost[("Nssu", 1)]
    for k in (135, 128, 127, 30):                                               # "near optimal control obtained at stage 1"
        assert abs(cost[("ssu", k)] - ref) < 1e-3 * ref
    X, _ = rollouts[("Nssu", 1)]
    assert np.all(np.abs(X[-1]) < 0.05)                                         # the regulator reaches the origin
    return cost


def interp_surface_desc(bellman, point):
    """One stage whose only 'control' sends every state to `point`: J_{N-1}(x) = F(point) with
    F = griddedInterpolant({s_r, s_r}, 2*X1.*X2 + X2, 'linear') (test_griddedInterp.m:27-33, :45)."""
    obj = bellman.Dynamic_Solver()
    s = bellman.tables.linspace(obj.x_min, obj.x_max, obj.dx)
    z = np.zeros((1, obj.dx))
    row = lambda v: np.full((1, 1), float(v))
    d = bellman.tables.Desc(n=[obj.dx, obj.dx], C=1, N=2, grid=[s[None].copy(), s[None].copy()], src_a=[0, 1], src_b=[-1, -1],
                            Ta=[z.copy(), z.copy()], Tb=[None, None], Tc=[row(point[0]), row(point[1])], q_order=[0, 1],
                            q=[z.copy(), z.copy()], r=row(0.0), store_J_all=True, store_idx_all=True).validate()
    X1, X2 = np.meshgrid(s, s, indexing="ij")
    return d, (2.0 * X1 * X2 + X2).ravel(order="F")
