"""Full-horizon sweeps of the MATLAB-literal restatement (oracle/matlab_literal.py) -> fixtures.

The reference stores no outputs for Solver_position / Solver_attitude / Solver_pos_att,
and MATLAB cannot run here, so the strongest independent anchor for those three classes is the
array-at-a-time literal restatement of their .m files, run over the horizons the reference runs:

  position   position-control/Solver_position.m:132-141   5999 stages (axis 1; the three axes share
             every parameter, :60-69)
  attitude   attitude-control/Solver_attitude.m:236-247   reference grid 1000x300; axis 1 1199 stages,
             axes 2 and 3 600 stages
  pos-att    pos-att/Solver_pos_att.m:270-286             reference grid 30x30x20x15x9; channel x 250
             stages, channels y and z 100 stages

The literal sweeps take minutes (numpy, S x C temporaries), so they run HERE, once, and the tests
compare the C oracle (CPU suite) and the CUDA path (GPU suite) with what is stored:

  * idx at every checkpoint, complete (uint8, 1-based as MATLAB returns it)
  * J at every checkpoint on a fixed sample of states (every `stride`-th state) + max|J| over all
  * near-tie report: over EVERY stage and state, the smallest relative gap between the
    best and the second-best control total, and how many gaps fall under 1e-10 / 1e-12
  * position only: the same recurrence evaluated in 80-bit extended precision on the same fp64
    tables (`exact_hi` + `exact_lo`), the neutral yardstick for the accumulated rounding of BOTH
    fp64 formulas (the literal's (1-t)*lo + t*hi and the normative fma(t, hi-lo, lo))

Run in the build container:  python tests/golden/make_literal_fixtures.py [position attitude posatt]
(about 15 minutes on 8 cores; independent cases run as separate processes).
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import matlab_literal as ml  # noqa: E402


def _gap_stats(tot, acc):
    """Relative gap between the two smallest control totals of every state."""
    part = np.partition(tot, 1, axis=-1)
    best, second = part[..., 0], part[..., 1]
    gap = (second - best) / np.maximum(np.abs(best), np.finfo(np.float64).tiny)
    acc["min_gap"] = min(acc["min_gap"], float(gap.min()))
    acc["n_lt_1e10"] += int(np.count_nonzero(gap < 1e-10))
    acc["n_lt_1e12"] += int(np.count_nonzero(gap < 1e-12))
    acc["n_exact_ties"] += int(np.count_nonzero(gap == 0.0))
    acc["state_stages"] += int(gap.size)


def _interp_any(gridvecs, V, queries, dtype):
    """GriddedInterpolantLinear.__call__ with the arithmetic carried out in `dtype`."""
    cells, ts = [], []
    for s, x in zip(gridvecs, queries):
        i = np.clip(np.searchsorted(s, x, side="right") - 1, 0, len(s) - 2)
        sl = s.astype(dtype)
        t = (x.astype(dtype) - sl[i]) / (sl[i + 1] - sl[i])
        cells.append(i)
        ts.append(t)
    D = len(gridvecs)
    vals = [V[tuple(cells[d] + ((m >> d) & 1) for d in range(D))] for m in range(1 << D)]
    for d in range(D):
        t = ts[d]
        vals = [(1 - t) * vals[2 * m] + t * vals[2 * m + 1] for m in range(len(vals) // 2)]
    return vals[0]


def _literal_sweep(grids, nxt, J_current, n_stages, checkpoints, stride, exact=False, check_period=0):
    """The reference's stage loop on the literal arrays.  checkpoints = stage counts (stages done)."""
    shape = tuple(len(g) for g in grids)
    F = ml.GriddedInterpolantLinear(grids, np.zeros(shape))
    Fx = np.zeros(shape, dtype=np.longdouble) if exact else None
    Jc_x = J_current.astype(np.longdouble) if exact else None
    S = int(np.prod(shape))
    sample = np.arange(0, S, stride, dtype=np.int64)
    acc = {"min_gap": np.inf, "n_lt_1e10": 0, "n_lt_1e12": 0, "n_exact_ties": 0, "state_stages": 0}
    out = {"sample": sample, "checkpoints": np.array(sorted(checkpoints), dtype=np.int64)}
    sums = []
    for done in range(1, n_stages + 1):
        tot = J_current + F(*nxt)
        _gap_stats(tot, acc)
        F.Values, idx = ml.ml_min_last(tot)
        if exact:
            totx = Jc_x + _interp_any(grids, Fx, nxt, np.longdouble)
            ix = np.argmin(totx, axis=-1)
            Fx = np.take_along_axis(totx, ix[..., None], axis=-1)[..., 0]
        if check_period and done % check_period == 0:
            sums.append((done, float(np.sum(F.Values.ravel(order="F"))), float(np.sum(idx))))
        if done in checkpoints:
            Jf = F.Values.ravel(order="F")
            out["idx_%d" % done] = idx.ravel(order="F").astype(np.uint8)
            out["J_%d" % done] = Jf[sample].copy()
            out["Jmax_%d" % done] = np.array(np.max(np.abs(Jf)))
            if exact:
                xf = Fx.ravel(order="F")[sample]
                hi = xf.astype(np.float64)
                out["exact_hi_%d" % done] = hi
                out["exact_lo_%d" % done] = (xf - hi.astype(np.longdouble)).astype(np.float64)
                out["exact_idx_%d" % done] = (ix + 1).ravel(order="F").astype(np.uint8)
    for k, v in acc.items():
        out["gap_" + k] = np.array(v)
    if sums:
        out["sums"] = np.array(sums)
    return out


def position_case():
    s = ml.SolverPositionLiteral()
    grids, nxt, Jc = s.axis_arrays(0)
    t0 = time.time()
    out = _literal_sweep(grids, nxt, Jc, s.N_stage - 1, {50, 400, 1000, 2000, 4000, s.N_stage - 1}, 7, exact=True)
    out["n_stages"] = np.array(s.N_stage - 1)
    np.savez_compressed(os.path.join(HERE, "literal_position_axis0.npz"), **out)
    return "position: %.0f s, min gap %.3g" % (time.time() - t0, float(out["gap_min_gap"]))


def attitude_case(axis):
    s = ml.SolverAttitudeLiteral()
    grids, nxt, Jc = s.axis_arrays(axis)
    n = 1199 if axis == 0 else 600
    cps = {60, 300, 600} | ({1199} if axis == 0 else set())
    t0 = time.time()
    out = _literal_sweep(grids, nxt, Jc, n, cps, 23)
    out["n_stages"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "literal_attitude_axis%d.npz" % axis), **out)
    return "attitude axis %d: %.0f s, min gap %.3g" % (axis, time.time() - t0, float(out["gap_min_gap"]))


def posatt_case(ch):
    s = ml.SolverPosAttLiteral()
    grids, nxt, Jc, _ = s.channel_arrays(ch)
    n = 250 if ch == 0 else 100
    cps = {50, 100} | ({250} if ch == 0 else set())
    t0 = time.time()
    out = _literal_sweep(grids, nxt, Jc, n, cps, 11, check_period=50)
    # the reference's check stages are k_s % 50 == 0 with k_s = N_stage - done (Solver_pos_att.m:273)
    out["n_stages"] = np.array(n)
    out["N_stage"] = np.array(s.N_stage)
    np.savez_compressed(os.path.join(HERE, "literal_posatt_ch%d.npz" % ch), **out)
    return "pos-att channel %d: %.0f s, min gap %.3g" % (ch, time.time() - t0, float(out["gap_min_gap"]))


def main(which):
    jobs = []
    with ProcessPoolExecutor(max_workers=7) as ex:
        if "posatt" in which:
            jobs += [ex.submit(posatt_case, ch) for ch in range(3)]
        if "position" in which:
            jobs.append(ex.submit(position_case))
        if "attitude" in which:
            jobs += [ex.submit(attitude_case, a) for a in range(3)]
        for j in jobs:
            print(j.result(), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["position", "attitude", "posatt"])
