"""Full-horizon parity of the CUDA path (through the C ABI) for the three classes the reference
holds no stored output for: bit-exact against the C oracle at every checkpoint, and u_star exact /
J within the stated tolerances against the MATLAB-literal fixtures (tests/golden/literal_*.npz,
made by tests/golden/make_literal_fixtures.py).  Tolerances: see tests/test_literal_full_horizon.py.
"""
import os

import numpy as np
import pytest

from test_literal_full_horizon import GOLD, check_against_fixture

pytestmark = pytest.mark.gpu


def run_to_checkpoints(sw, oracle_lib, d, checkpoints, **run_kw):
    """Runs the CUDA sweep and the oracle side by side, checkpoint to checkpoint; asserts bit
    equality at each and returns {cp: (J [P,S], idx [P,S])} from the CUDA side."""
    out, Jo, done = {}, None, 0
    for cp in checkpoints:
        sw.run(cp - done, **run_kw)
        o = oracle_lib.sweep(d, n_stages=cp - done, J_N=Jo)
        Jo, done = o["J_last"], cp
        Jg, Ig = sw.get_J(), sw.get_idx()
        assert np.array_equal(Ig, o["idx_last"]), "argmin differs from the oracle after %d stages" % cp
        assert np.array_equal(Jg, Jo), "J differs from the oracle after %d stages" % cp
        out[cp] = (Jg, Ig)
    return out


def test_position_full_5999_stages_persistent(bellman, oracle_lib):
    """config 2, every stage of Solver_position.m:132-141: 3 axes x 201x201 x 3 controls in the
    persistent kernel (one cooperative launch per checkpoint span)."""
    d = bellman.tables.stack_problems(bellman.Solver_position()._axis_descs())
    f = np.load(os.path.join(GOLD, "literal_position_axis0.npz"))
    sw = bellman.Sweep(d)
    got = run_to_checkpoints(sw, oracle_lib, d, [int(c) for c in f["checkpoints"]], use_graph=True)
    assert sw.last_kernel == "persistent" and sw.current_stage == 1
    for p in range(3):          # the three axes share every parameter (Solver_position.m:60-69)
        rows = [(cp, J[p], I[p] + 1) for cp, (J, I) in got.items()]
        print("\n".join(check_against_fixture("position axis %d (CUDA)" % (p + 1), rows, f, 5e-12)))
    sw.close()


def test_attitude_reference_grid_long(bellman, oracle_lib):
    """Solver_attitude.m:236-247, reference grid 3 x 1000 x 300 x 3: 600 stages on all axes, 1199 on
    axis 1 (the strip kernel)."""
    d = bellman.tables.stack_problems(bellman.Solver_attitude()._axis_descs())
    fx = [np.load(os.path.join(GOLD, "literal_attitude_axis%d.npz" % a)) for a in range(3)]
    sw = bellman.Sweep(d)
    got = run_to_checkpoints(sw, oracle_lib, d, [60, 300, 600, 1199])
    assert sw.last_kernel == "window:strip"
    for a in range(3):
        rows = [(cp, J[a], I[a] + 1) for cp, (J, I) in got.items() if "idx_%d" % cp in fx[a].files]
        print("\n".join(check_against_fixture("attitude axis %d (CUDA)" % (a + 1), rows, fx[a], 1e-12)))
    sw.close()


def test_pos_att_reference_grid_long(bellman, oracle_lib):
    """Solver_pos_att.m:270-286, reference grid 3 x 30x30x20x15 x 9: 250 stages with the Sigma-check of
    :273-285 running every 50 stages (the streaming factorised kernel)."""
    sp = bellman.Solver_pos_att()
    d = bellman.tables.stack_problems([sp.channel_desc(ch) for ch in range(3)])
    fx = [np.load(os.path.join(GOLD, "literal_posatt_ch%d.npz" % ch)) for ch in range(3)]
    sw = bellman.Sweep(d)
    got = run_to_checkpoints(sw, oracle_lib, d, [50, 100, 250], check_period=50, check_tol=0.0)
    assert sw.last_kernel in ("stream", "tile"), sw.last_kernel
    for ch in range(3):
        rows = [(cp, J[ch], I[ch] + 1) for cp, (J, I) in got.items() if "idx_%d" % cp in fx[ch].files]
        print("\n".join(check_against_fixture("pos-att channel %d (CUDA)" % ch, rows, fx[ch], 1e-12)))
    sw.close()
