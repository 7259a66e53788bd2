"""The reference's own test scripts replayed through the CUDA path (facade -> C ABI):
test/test_u_star_M.m and test/test_griddedInterp.m (see tests/test_reference_scripts_cpu.py)."""
import numpy as np
import pytest

from _ref_scripts import SSU_STAGES, QUERY_POINTS, check_u_star_M_verdicts, interp_surface_desc

pytestmark = pytest.mark.gpu


def test_script_test_u_star_M(bellman, oracle_lib):
    D = bellman.Dynamic_Solver()
    D.store_J_star = False
    D.run()                                                            # run(D), default size
    d = D._desc
    ora = oracle_lib.sweep(d, keep_all=True)
    assert np.array_equal(D._sweep.get_idx(1), ora["idx_all"][0]) and np.array_equal(D._sweep.get_J(1), ora["J_all"][0])
    rollouts = {}
    for key, mode, ssu in [(("ssu", k), "ssu", k) for k in SSU_STAGES] + [(("Nssu", 1), "Nssu", 1)]:
        X, U = D.get_optimal_path([2.0, 1.0], mode, ssu) if key[0] == "ssu" else D.get_optimal_path()
        Xo, Uo = oracle_lib.rollout(d, ora["idx_all"][:, 0, :], D.A, D.B, d.meta["U_mesh"], np.array([[2.0, 1.0]]),
                                    mode=1 if mode == "ssu" else 0, ssu_stage=ssu)
        np.testing.assert_array_equal(X.T, Xo[0])                      # NaNs of the diverging 'ssu',190 loop compare equal
        np.testing.assert_array_equal(U, Uo[0])
        rollouts[key] = (X.T, U)
    check_u_star_M_verdicts(D, rollouts)


@pytest.mark.parametrize("kernel", ["auto", "direct", "splitc"])
def test_script_test_griddedInterp(bellman, kernel):
    from test_gpu_parity import KERNELS
    for pt in QUERY_POINTS:
        d, JN = interp_surface_desc(bellman, pt)
        sw = bellman.Sweep(d)
        sw.set_J(JN.reshape(1, -1))
        sw.run(1, kernel=KERNELS[kernel])
        want = 2.0 * pt[0] * pt[1] + pt[1]
        np.testing.assert_allclose(sw.get_J()[0], want, rtol=1e-12, atol=1e-12)
        assert np.all(sw.get_idx() == 0)
        sw.close()
