"""Full-horizon parity of the C oracle against the MATLAB-literal restatement (CPU suite).

The reference stores no outputs for Solver_position / Solver_attitude / Solver_pos_att, so these
classes are anchored on oracle/matlab_literal.py run over the horizons the reference runs
(Solver_position.m:132-141 all 5999 stages, Solver_attitude.m:236-247, Solver_pos_att.m:270-286);
the literal sweeps take minutes and are stored as fixtures by tests/golden/make_literal_fixtures.py.

What is asserted, and the tolerance that goes with it:
  * u_star (the argmin index) EXACT at every checkpoint, every state;
  * the near-tie report of SURVEY 8c(4): no exact tie and no best-vs-second-best gap under 1e-12
    anywhere on the path (so exactness of u_star is not luck);
  * J, per stage (seeded from the literal's own J_{k+1}, the method of SURVEY 4.4): <= 1e-12 relative
    (measured ~1e-15);
  * J, accumulated over the horizon: <= 1e-12 relative to the recurrence evaluated in 80-bit
    extended precision on the same fp64 tables (position, all 5999 stages: measured 7.0e-13).  The
    literal's own formula (1-t)*lo + t*hi drifts 3.0e-12 from that yardstick over the same horizon, so
    oracle-vs-literal is bounded by 5e-12 here (measured 2.9e-12) — two fp64 formulas cannot agree
    to better than the sum of their own accumulated roundings, and MATLAB's is closed source.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sweep_to_checkpoints(oracle_lib, d, fixture, p=0):
    """Runs the C oracle to every checkpoint of the fixture.  Yields (stages_done, J[p], idx1[p])."""
    J, done = None, 0
    for cp in (int(c) for c in fixture["checkpoints"]):
        out = oracle_lib.sweep(d, n_stages=cp - done, J_N=J)
        J, done = out["J_last"], cp
        yield cp, J[p], out["idx_last"][p] + 1


def check_against_fixture(name, rows, f, tol_literal, capsys=None):
    """rows: iterable of (cp, J, idx1).  Returns the report lines."""
    assert int(f["gap_n_exact_ties"]) == 0 and int(f["gap_n_lt_1e12"]) == 0, "fixture holds a near tie"
    lines = []
    for cp, J, idx1 in rows:
        ref_idx = f["idx_%d" % cp].astype(np.int32)
        mism = np.flatnonzero(idx1 != ref_idx)
        Js = J[f["sample"]]
        jmax = float(f["Jmax_%d" % cp])
        rel = float(np.max(np.abs(Js - f["J_%d" % cp])) / jmax)
        line = "%s after %d stages: u_star mismatches %d/%d, J max rel vs literal %.3g" % (name, cp, mism.size, idx1.size, rel)
        if "exact_hi_%d" % cp in f.files:
            ex = f["exact_hi_%d" % cp].astype(np.longdouble) + f["exact_lo_%d" % cp].astype(np.longdouble)
            e_ours = float(np.max(np.abs(Js.astype(np.longdouble) - ex)) / jmax)
            e_lit = float(np.max(np.abs(f["J_%d" % cp].astype(np.longdouble) - ex)) / jmax)
            line += "; vs 80-bit recurrence: ours %.3g, literal %.3g" % (e_ours, e_lit)
            assert e_ours <= 1e-12, line
            assert np.array_equal(f["exact_idx_%d" % cp], f["idx_%d" % cp]), "literal argmin differs from the 80-bit argmin"
        lines.append(line)
        # a mismatch is reported with the smallest best-vs-second-best gap on the path
        assert mism.size == 0, line + " (smallest gap on the whole path: %.3g)" % float(f["gap_min_gap"])
        assert rel <= tol_literal, line
    lines.append("%s near-tie report over %d state-stages: min gap %.3g, gaps < 1e-10: %d, < 1e-12: %d, exact ties: %d" % (
        name, int(f["gap_state_stages"]), float(f["gap_min_gap"]), int(f["gap_n_lt_1e10"]), int(f["gap_n_lt_1e12"]),
        int(f["gap_n_exact_ties"])))
    return lines


def test_position_full_5999_stages(bellman, oracle_lib):
    """Solver_position.m:132-141, every stage; the three axes share all parameters (:60-69)."""
    d = bellman.Solver_position()._axis_descs()[0]
    f = np.load(os.path.join(GOLD, "literal_position_axis0.npz"))
    assert int(f["n_stages"]) == d.N - 1 == 5999
    print("\n".join(check_against_fixture("position axis 1", sweep_to_checkpoints(oracle_lib, d, f), f, 5e-12)))


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_attitude_reference_grid(bellman, oracle_lib, axis):
    """Solver_attitude.m:236-247 on the reference 1000 x 300 grid: 1199 stages (axis 1), 600 (2, 3)."""
    d = bellman.Solver_attitude()._axis_descs()[axis]
    f = np.load(os.path.join(GOLD, "literal_attitude_axis%d.npz" % axis))
    print("\n".join(check_against_fixture("attitude axis %d" % (axis + 1), sweep_to_checkpoints(oracle_lib, d, f), f, 1e-12)))


@pytest.mark.parametrize("ch", [0, 1, 2])
def test_pos_att_reference_grid(bellman, oracle_lib, ch):
    """Solver_pos_att.m:270-286 on the reference 30x30x20x15x9 grid: 250 stages (x), 100 (y, z);
    the reference's Sigma-check stages (:273-285) are checkpoints."""
    d = bellman.Solver_pos_att().channel_desc(ch)
    f = np.load(os.path.join(GOLD, "literal_posatt_ch%d.npz" % ch))
    rows = list(sweep_to_checkpoints(oracle_lib, d, f))
    print("\n".join(check_against_fixture("pos-att channel %d" % ch, rows, f, 1e-12)))
    # the two numbers the reference prints at its check stages: sum(F.Values) and sum(U_Optimal_id)
    by_stage = {cp: (J, idx1) for cp, J, idx1 in rows}
    for done, fsum, idsum in f["sums"]:
        if int(done) in by_stage:
            J, idx1 = by_stage[int(done)]
            assert float(np.sum(idx1)) == idsum
            assert abs(float(np.sum(J)) - fsum) <= 1e-12 * abs(fsum)


def test_per_stage_seeded_tolerance(bellman, oracle_lib):
    """SURVEY 4.4's method: seed the oracle with the literal's J_{k+1} and compare ONE stage, so the
    bound is the per-stage rounding (<= 1e-12, measured ~1e-15), not the accumulated one."""
    from oracle import matlab_literal as ml
    cases = []
    s = ml.SolverPositionLiteral()
    cases.append(("position", s.axis_arrays(0), bellman.Solver_position()._axis_descs()[0], 25))
    s = ml.SolverAttitudeLiteral()
    cases.append(("attitude", s.axis_arrays(1), bellman.Solver_attitude()._axis_descs()[1], 4))
    s = ml.SolverPosAttLiteral()
    g, nxt, Jc, _ = s.channel_arrays(2)
    cases.append(("pos-att", (g, nxt, Jc), bellman.Solver_pos_att().channel_desc(2), 2))
    for name, (grids, nxt, Jc), d, n in cases:
        F = ml.GriddedInterpolantLinear(grids, np.zeros(tuple(len(x) for x in grids)))
        worst = 0.0
        for _ in range(n):
            J_next = F.Values.ravel(order="F").copy()
            F.Values, idx = ml.ml_min_last(Jc + F(*nxt))
            Jo, Io = oracle_lib.stage(d, J_next[None, :])
            assert np.array_equal(Io[0] + 1, idx.ravel(order="F")), name
            ref = F.Values.ravel(order="F")
            worst = max(worst, float(np.max(np.abs(Jo[0] - ref)) / np.max(np.abs(ref))))
        print("%s: per-stage seeded J max rel %.3g over %d stages" % (name, worst, n))
        assert worst <= 1e-12
