"""Pins the oracle against the reference's only golden vector (test/obj_1.mat -> golden/obj_1.npz)
and the known-answer values of SURVEY.md 4.3 / 4.4."""
import hashlib

import numpy as np

from conftest import to_grid


def kirk_desc_from_golden(bellman, g):
    return bellman.tables.kirk_desc(g["A"], g["B"], g["Q"], g["R"], g["N"], g["x_min"], g["x_max"], g["dx"],
                                    g["u_min"], g["u_max"], g["du"])


def test_fixture_integrity(golden):
    assert hashlib.sha256(golden["u_star"].tobytes(order="F")).hexdigest() == str(golden["sha256_u_star"])
    assert hashlib.sha256(golden["J_star"].tobytes(order="F")).hexdigest() == str(golden["sha256_J_star"])
    assert str(golden["sha256_u_star"]).startswith("36d6682834fd0e80")
    assert str(golden["sha256_J_star"]).startswith("8c8ae87b0204d7cf")
    J, u = golden["J_star"], golden["u_star"]
    assert J.shape == u.shape == (35, 35, 130)
    assert J[0, 0, 0] == 165.01130170072753 and J[17, 17, 0] == 4.243255734000523
    assert J[34, 34, 0] == 234.48408257473878
    assert u[0, 0, 0] == 10.0 and u[17, 17, 0] == -2.1212121212121247 and u[34, 34, 0] == -19.292929292929294
    assert np.all(J[:, :, 129] == 0) and len(golden["u_vals"]) == 61


def test_linspace_reproduces_golden_mesh(bellman, golden):
    s = bellman.tables.linspace(golden["x_min"], golden["x_max"], golden["dx"])
    assert np.array_equal(s, golden["X1_mesh"][:, 0])
    assert np.array_equal(s, golden["X2_mesh"][0, :])


def test_c_oracle_matches_golden(bellman, oracle_lib, golden):
    d = kirk_desc_from_golden(bellman, golden)
    N, dx = golden["N"], golden["dx"]
    out = oracle_lib.sweep(d, keep_all=True)
    assert out["stage"] == 1
    U = d.meta["U_mesh"]
    for k in range(1, N):        # stages 1..N-1
        idx = to_grid(out["idx_all"][k - 1, 0], (dx, dx))
        assert np.array_equal(U[idx], golden["u_star"][:, :, k - 1]), f"u_star differs at stage {k}"
        J = to_grid(out["J_all"][k - 1, 0], (dx, dx))
        ref = golden["J_star"][:, :, k - 1]
        assert np.max(np.abs(J - ref)) <= 1e-12 * np.max(np.abs(ref)), f"J differs at stage {k}"


def test_literal_oracle_matches_golden(golden):
    from oracle import matlab_literal as ml
    L = ml.DynamicSolverLiteral(N=golden["N"], dx=golden["dx"], du=golden["du"], A=golden["A"], B=golden["B"],
                                Q=golden["Q"], R=golden["R"]).run()
    assert np.array_equal(L.u_star, golden["u_star"])
    for k in range(golden["N"] - 1):
        ref = golden["J_star"][:, :, k]
        assert np.max(np.abs(L.J_star[:, :, k] - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_rollout_known_answer(bellman, oracle_lib, golden):
    """rollout from [2;1] on the golden u_star."""
    d = kirk_desc_from_golden(bellman, golden)
    out = oracle_lib.sweep(d, keep_all=True)
    X, U = oracle_lib.rollout(d, out["idx_all"][:, 0, :], golden["A"], golden["B"], d.meta["U_mesh"], [[2.0, 1.0]])
    np.testing.assert_allclose(U[0, :5], [-7.30945822, -4.02204735, -2.44352418, -0.69167883, 1.37694508],
                               atol=5e-9)
    assert np.argmax(U[0]) + 1 == 14 and abs(U[0].max() - 7.056726089) < 5e-9
    np.testing.assert_allclose(X[0, -1], [0.02094326, -0.05365354], atol=5e-9)
    # and the literal (MATLAB-shaped) rollout agrees
    from oracle import matlab_literal as ml
    L = ml.DynamicSolverLiteral(N=golden["N"], dx=golden["dx"], du=golden["du"]).run()
    X2, U2 = L.get_optimal_path()
    np.testing.assert_allclose(U2, U[0], atol=1e-9)
    np.testing.assert_allclose(X2.T, X[0], atol=1e-9)


def test_locate_rule_agrees_between_library_and_oracle(bellman, oracle_lib, golden):
    d = kirk_desc_from_golden(bellman, golden)
    assert np.array_equal(bellman.query_locate(d).ravel(), oracle_lib.locate_modes(d))
    sp = bellman.Solver_pos_att()
    d4 = sp.channel_desc(0)
    lm = bellman.query_locate(d4)
    assert np.array_equal(lm.ravel(), oracle_lib.locate_modes(d4))
    # even point counts => two spacings => SEARCH; n_mesh_w = 15 is odd => uniform
    assert list(lm[0]) == [1, 1, 1, 0]
