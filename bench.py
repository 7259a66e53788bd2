#!/usr/bin/env python
"""bench.py — Bellman state·control updates/s of the backward sweep on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU; the driver launches this)

A "step" is ONE backward stage of the named workload over the whole state grid (all ranks
together): the reference's  [F.Values, idx] = min(J_current + F(x_next..), [], ctrl_dim).
Headline workload (BASELINE.json configs[3]): Kirk two-state example scaled to an
8192 x 8192 state grid x 512 controls, the grid slab-partitioned over the GPUs (strong scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bellman_state_control_updates_per_s"
UNIT = "updates/s"

WORKLOADS = {
    # name: (kind, params)
    "kirk_scaled_8192x8192x512": dict(kind="kirk", dx=8192, du=512, N=200),
    "kirk_default_100x100x1000": dict(kind="kirk", dx=100, du=1000, N=200),
    "position_3x201x201x3": dict(kind="position"),
    "attitude_x4_3x4000x1200x3": dict(kind="attitude", n_w=4000, n_t=1200),
    "attitude_x16_3x16000x4800x3": dict(kind="attitude", n_w=16000, n_t=4800),
    "pos_att_ref_30x30x20x15x9": dict(kind="pos_att", scale=1),
    "pos_att_x4_120x120x80x60x9": dict(kind="pos_att", scale=4),
    "pos_att_x8_1ch_240x240x160x120x9": dict(kind="pos_att", scale=8, channels=1),   # single-GPU slice of configs[4]
    # BASELINE.json configs[4]: the coupled position+attitude grid sized to fill the HBM of the GPUs it runs
    # on.  One channel, 480 x 480 x (160 per GPU) x 120 states = 4.4e9 states (88 GB of J + argmin) PER GPU;
    # the theta range grows with the GPU count (constant resolution), the grid is cut into theta slabs.
    "pos_att_cfg5_480x480x160Nx120x9": dict(kind="pos_att_cfg5", n=(480, 480, 160, 120)),
}
DEFAULT_WORKLOAD = "kirk_scaled_8192x8192x512"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE stage-kernel launch.  NOT measured by this run (a
# run under ncu is never a bench value): constants copied from the committed `ncu --set full` captures of
# the same command, each with the file it comes from; reported as roofline.traffic + traffic_source.
NCU_TRAFFIC = {"kirk_scaled_8192x8192x512": (544.82e6 + 766.83e6, "profiles/r02_wide_kirk_summary.txt"),
               "attitude_x16_3x16000x4800x3": (1.8434e9 + 2.7178e9, "profiles/r01_strip_att16_summary.txt"),
               "pos_att_x4_120x120x80x60x9": (2.0363e9 + 2.4638e9, "profiles/r02_stream_posatt_x4_summary.txt")}


def make_desc(bb, name, world=1):
    w = WORKLOADS[name]
    t = bb.tables
    if w["kind"] == "pos_att_cfg5":
        s = bb.Solver_pos_att()
        nx, nv, nt, nw = w["n"]
        s.n_mesh_x, s.n_mesh_v, s.n_mesh_t, s.n_mesh_w = nx, nv, nt * world, nw
        s.theta1_min, s.theta1_max = s.theta1_min * world, s.theta1_max * world   # constant resolution: the range grows with the GPUs
        return s.channel_desc(0)
    if w["kind"] == "kirk":
        o = bb.Dynamic_Solver()
        return t.kirk_desc(o.A, o.B, o.Q, o.R, w["N"], o.x_min, o.x_max, w["dx"], o.u_min, o.u_max, w["du"],
                           store_J_all=False, store_idx_all=False)
    if w["kind"] == "position":
        return t.stack_problems(bb.Solver_position()._axis_descs())
    if w["kind"] == "attitude":
        s = bb.Solver_attitude()
        s.n_mesh_w, s.n_mesh_t = w["n_w"], w["n_t"]
        return t.stack_problems(s._axis_descs())
    if w["kind"] == "pos_att":
        s = bb.Solver_pos_att()
        k = w["scale"]
        s.n_mesh_x, s.n_mesh_v, s.n_mesh_t, s.n_mesh_w = 30 * k, 30 * k, 20 * k, 15 * k
        return t.stack_problems([s.channel_desc(c) for c in range(w.get("channels", 3))])
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (kind "port": MATLAB/Octave do not exist on the box) on host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample_rate(d, seconds_target, threads=0):
    """Times oracle_stage_points on a random sample of states of problem 0 sized for roughly
    `seconds_target` of CPU work.  Returns (updates/s, cores, sample description, seconds)."""
    from oracle import cbind
    cbind.build()
    # every host core, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    cbind.set_threads(threads or len(os.sched_getaffinity(0)))
    cores = cbind.num_threads()
    rng = np.random.default_rng(0)
    S = d.S
    s0 = d.grid[0][0]
    # a smooth synthetic J_{k+1} (the work per update does not depend on the values)
    Jn = np.zeros(S) if S > 2 ** 28 else rng.normal(size=S)
    n = min(S, 20000)
    pts = rng.integers(0, S, size=n)
    t0 = time.perf_counter(); cbind.stage_points(d, Jn, pts); dt = time.perf_counter() - t0
    rate = n * d.C / max(dt, 1e-9)
    n = int(min(max(rate * seconds_target / d.C, n), 5e7))
    pts = rng.integers(0, S, size=n)
    t0 = time.perf_counter(); cbind.stage_points(d, Jn, pts); dt = time.perf_counter() - t0
    return n * d.C / dt, cores, "%d random states x %d controls of one stage (%.1f s)" % (n, d.C, dt), dt


def cpu_reference_shaped_rate(bb, name):
    """The array-at-a-time numpy restatement (oracle/matlab_literal.py: S x C temporaries, one
    interpolation, one add, one min per stage — the shape of the reference's own MATLAB code) on a
    down-scaled grid of the same problem; single process, numpy's own threading."""
    from oracle import matlab_literal as ml
    w = WORKLOADS[name]
    kind = "port (numpy, reference-shaped)"
    if w["kind"] in ("position", "attitude"):
        # one axis at reference size (Solver_position.m:132-141 / Solver_attitude.m:236-247), a few stages
        L = ml.SolverPositionLiteral() if w["kind"] == "position" else ml.SolverAttitudeLiteral()
        grids, nxt, J_current = L.axis_arrays(0)
        F = ml.GriddedInterpolantLinear(grids, np.zeros((len(grids[0]), len(grids[1]))))
        n, t0 = 0, time.perf_counter()
        while n < 3 or time.perf_counter() - t0 < 2.0:
            F.Values, _idx = ml.ml_min_last(J_current + F(*nxt))
            n += 1
        dt = time.perf_counter() - t0
        return {"value": n * J_current.size / dt, "unit": UNIT, "kind": kind,
                "sample": "one axis at reference size, %s states x %d controls, %d stages (%.1f s)"
                          % ("x".join(str(len(g)) for g in grids), J_current.shape[-1], n, dt)}
    if w["kind"] in ("pos_att", "pos_att_cfg5"):
        L = ml.SolverPosAttLiteral()       # one channel at reference size (Solver_pos_att.m:270-286)
        t0 = time.perf_counter()
        n = 3
        grids = L.calculate_one_channel(0, n_stages=n, check=False)[2]
        dt = time.perf_counter() - t0
        S = int(np.prod([len(g) for g in grids]))
        return {"value": n * S * 9 / dt, "unit": UNIT, "kind": kind,
                "sample": "one channel at reference size (%d states x 9 controls), %d stages incl. set-up (%.1f s)" % (S, n, dt)}
    if w["kind"] != "kirk":
        return None
    dx, du = min(w["dx"], 256), min(w["du"], 256)
    L = ml.DynamicSolverLiteral(N=4, dx=dx, du=du)
    L.setup()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        JF = L.F(L.X_next_M1, L.X_next_M2)
        L.F.Values, _idx = ml.ml_min_last(JF + L.J_current_state)
    dt = time.perf_counter() - t0
    return {"value": n * dx * dx * du / dt, "unit": UNIT, "kind": kind,
            "sample": "%dx%d states x %d controls, %d stages (%.1f s)" % (dx, dx, du, n, dt)}


def workload_config(name, d):
    """The `config` object, identical in both arms (the driver compares them)."""
    S_all = d.S * d.P
    return {"workload": name, "grid": d.n, "controls": d.C, "problems": d.P,
            "step": "one backward stage over the whole grid",
            "l2": "J_{k+1} (%.0f MB) exceeds the 126 MB L2; no flush needed" % (S_all * 8 / 1e6)
            if S_all * 8 > 130e6 else "inputs fit L2 (stage-to-stage reuse is the workload)"}


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path on the box's host cores.  MATLAB is not
    installable (no toolchain, no network), so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bellman_b200 as bb
    d = make_desc(bb, args.workload)
    per_step = args.ref_seconds or max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample_rate(d, per_step)
    rates, dts = [], []
    for _ in range(args.steps):
        r, cores, sample, dt = cpu_sample_rate(d, per_step)
        rates.append(r); dts.append(dt)
    value = float(np.mean(rates))
    S_all = d.S * d.P
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * S_all * d.C / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, d),
            "run": {"note": "ms_per_step extrapolated from the sample to one full stage"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_slab_rate(d, seconds_target, threads=0):
    """The oracle on a CONTIGUOUS block of states of problem 0 (consecutive linear indices: whole rows of
    the grid, cache-friendly), the counterpart of cpu_sample_rate's random states.  Needs J_{k+1} on the
    host, so only for grids up to 2^28 states.  Returns (updates/s, cores, description) or None."""
    from oracle import cbind
    if d.S > 2 ** 28:
        return None
    cbind.set_threads(threads or len(os.sched_getaffinity(0)))
    cores = cbind.num_threads()
    rng = np.random.default_rng(1)
    Jn = rng.normal(size=d.S)
    n = min(d.S, 20000)
    lo = (d.S - n) // 2
    t0 = time.perf_counter(); cbind.stage_points(d, Jn, np.arange(lo, lo + n)); dt = time.perf_counter() - t0
    n = int(min(d.S, max(n, n / max(dt, 1e-9) * seconds_target)))
    lo = (d.S - n) // 2
    t0 = time.perf_counter(); cbind.stage_points(d, Jn, np.arange(lo, lo + n)); dt = time.perf_counter() - t0
    return n * d.C / dt, cores, "%d consecutive states (%.1f rows of dimension 0) x %d controls of one stage (%.1f s)" % (
        n, n / d.n[0], d.C, dt)


def sharded_parity(bb, sw, d, rank, world, part_dim, kernel, n_points=3000):
    """Outside the timed region: two stages from a zero terminal cost on the (sharded) handle, then this
    rank's states of the SECOND stage are compared bit for bit with the oracle's pointwise evaluator, which
    rebuilds J of the first stage in closed form (so this works at grid sizes no host array could hold).
    The second stage reads the halo values the NEIGHBOURS stored during the first one, so the sample is
    concentrated on the slab faces.  Returns True when every sampled state matches."""
    from oracle import cbind
    cbind.build()
    cbind.set_threads(len(os.sched_getaffinity(0)) // max(1, min(world, 8)) or 1)
    sw.set_J(None)
    sw.run(2, kernel=kernel)
    rng = np.random.default_rng(100 + rank)
    own = [(0, n) for n in d.n]
    if world > 1:
        own[part_dim] = (sw.slab[0], sw.slab[1])
    coords = []
    for k, (lo, hi) in enumerate(own):
        c = rng.integers(lo, hi, size=n_points)
        if world > 1 and k == part_dim:        # two thirds of the sample within 4 indices of a slab face
            m = n_points // 3
            c[:m] = np.minimum(lo + rng.integers(0, 4, size=m), hi - 1)
            c[m:2 * m] = np.maximum(hi - 1 - rng.integers(0, 4, size=m), lo)
        coords.append(c.astype(np.int64))
    lin = np.zeros(n_points, dtype=np.int64)
    stride = 1
    for k in range(d.D):
        lin += coords[k] * stride
        stride *= d.n[k]
    ok = True
    for p in range(d.P):
        Jg, Ig = sw.get_points(lin, prob=p)
        Jo, Io = cbind.stage_points(d, None, lin, p=p)
        ok = ok and bool(np.array_equal(Jg, Jo) and np.array_equal(Ig, Io))
    return ok


def rollout_rate(bb, local):
    """get_optimal_path rollouts of the default Kirk problem (100 x 100 x 1000, N = 200) from a
    64 x 64 lattice of initial states, one GPU thread per x0, through bellman_rollout (host buffers)."""
    o = bb.Dynamic_Solver()
    d = o._build()
    sw = bb.Sweep(d, device=local).run()
    g = np.linspace(o.x_min, o.x_max, 64)
    x0 = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
    import torch
    X = torch.empty((len(x0), d.N, 2), dtype=torch.float64).pin_memory().numpy()      # pinned host buffers for the results
    U = torch.empty((len(x0), d.N), dtype=torch.float64).pin_memory().numpy()
    sw.rollout(o.A, o.B, d.meta["U_mesh"], x0, out=(X, U))     # warm-up
    reps, dts = 7, []
    for _ in range(reps):
        t0 = time.perf_counter()
        sw.rollout(o.A, o.B, d.meta["U_mesh"], x0, out=(X, U))
        dts.append(time.perf_counter() - t0)
    dt = float(np.median(dts))
    # one 'nearest' policy query per call, as a user's own simulation loop would issue them (U_Opt(x, v) per step)
    q = np.array([[0.3, -0.2]])
    sw.policy_lookup(q, stage=1)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        sw.policy_lookup(q, stage=1)
        lat.append(time.perf_counter() - t0)
    sw.close()
    return {"x0": len(x0), "steps": d.N - 1, "ms": dt * 1e3, "policy_lookup_single_query_us": float(np.median(lat)) * 1e6, "trajectories_per_s": len(x0) / dt,
            "state_steps_per_s": len(x0) * (d.N - 1) / dt, "call": "bellman_rollout (host x0 in, X and U out into pinned host arrays; median of %d calls)" % reps}


def plant_rollout_rate(bb, local):
    """Solver_pos_att.get_optimal_path (13-state plant, one ode45 call per stage) for a
    batch of initial states through bellman_rollout_pos_att (host buffers), against the C restatement on
    the host cores for a bounded sub-batch (the checker, timed as the CPU arm of this row)."""
    from oracle import cbind
    sp = bb.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 12, 10, 8, 7
    sp.check_period = 0
    sp.device = local
    descs, idxs, fvals = [], [], []
    for ci, ch in enumerate("xyz"):
        ctl = sp.calculate_one_channel_U_Opt(ci, n_stages=150)
        sp.set_controller(ctl, ch)
        d = sp.channel_desc(ci)
        descs.append(d)
        idxs.append((np.asarray(ctl["U_Optimal_id"]) - 1).astype(np.int32).ravel(order="F"))
        fvals.append(np.stack([d.meta[k] for k in ("f0_allcomb", "f1_allcomb", "f6_allcomb", "f7_allcomb")]))
    rng = np.random.default_rng(1)
    batch, n_steps = 4096, 200
    y0 = np.zeros((batch, 13))
    y0[:, 0:3] = rng.uniform(-0.18, 0.18, size=(batch, 3))
    y0[:, 3:6] = rng.uniform(-0.08, 0.08, size=(batch, 3))
    y0[:, 6:9] = rng.uniform(-0.04, 0.04, size=(batch, 3))
    y0[:, 9] = np.sqrt(1 - np.sum(y0[:, 6:9] ** 2, axis=1))
    y0[:, 10:13] = rng.uniform(-0.03, 0.03, size=(batch, 3))
    sp.get_optimal_path(y0[:64], n_steps=8)                       # warm-up (creates the channel handles)
    dts = []
    for _ in range(3):
        t0 = time.perf_counter()
        Xg, Fg, _ = sp.get_optimal_path(y0, n_steps=n_steps)
        dts.append(time.perf_counter() - t0)
    dt = float(np.median(dts))
    t0 = time.perf_counter()
    sp.get_optimal_path(y0, n_steps=n_steps, stride_out=n_steps)   # only the final states come back: the kernel alone
    dt_last = time.perf_counter() - t0
    # a batch that fills the GPU (8 warps of 255 registers per SM): the same states, eight times over
    big = np.tile(y0, (8, 1))
    t0 = time.perf_counter()
    sp.get_optimal_path(big, n_steps=n_steps, stride_out=n_steps)
    dt_big = time.perf_counter() - t0
    R0, V0 = sp.get_target_R0V0()
    nb = 256
    t0 = time.perf_counter()
    Xo, Fo, _, _ = cbind.rollout_pos_att(descs, idxs, fvals, y0[:nb], n_steps, sp.h, R0, V0, sp.InertiaM, sp.Mass, sp.T_dist)
    dt_cpu = time.perf_counter() - t0
    same = np.all(Fg[:nb] == Fo, axis=(1, 2))
    ok = bool(same.mean() >= 0.95 and np.allclose(Xg[:nb][same], Xo[same], rtol=0, atol=1e-9))
    return {"x0": batch, "stages": n_steps, "ms": dt * 1e3, "trajectories_per_s": batch / dt,
            "ode45_steps_per_s": batch * n_steps * 10 / dt, "ms_final_state_only": dt_last * 1e3,
            "trajectories_per_s_batch_32768": len(big) / dt_big,
            "cpu_port_trajectories_per_s": nb / dt_cpu,
            "cpu_cores": cbind.num_threads(), "parity_vs_oracle": "pass" if ok else "FAIL",
            "call": "bellman_rollout_pos_att (host x0 in; X, thruster levels, forces/moments out; median of 3 calls)"}


def dense6_rate(bb, local):
    """the coupled 6-D attitude sweep (Solver_attitude.run) on a 24^3 x 10^3 mesh with
    27 control combinations, device time of the stage loop (bellman_dense6_run; tables built on the host
    once, as the reference precomputes its next-state arrays), two stages on a smaller mesh from a rough
    terminal cost checked bit for bit against the C restatement."""
    from oracle import cbind
    sa = bb.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_q = 24, 10
    sa.device = local
    T = sa.dense6_tables()
    stages = 10
    bb.dense6_run(T, 2, device=local)                     # warm-up (module load, first-touch)
    _, _, ms = bb.dense6_run(T, stages, device=local)
    # parity and the CPU arm of this row on a 12^3 x 6^3 mesh (the C restatement needs ~1.2 us per update and core)
    sa.n_mesh_w, sa.n_mesh_q = 12, 6
    Ts = sa.dense6_tables()
    JN = np.random.default_rng(0).normal(size=Ts.S)
    Jg, Ig, _ = bb.dense6_run(Ts, 2, J_N=JN, device=local)
    t0 = time.perf_counter()
    Jo, Io = cbind.dense6_run(Ts, 2, J_N=JN)
    dt_cpu = time.perf_counter() - t0
    ok = bool(np.array_equal(Jg, Jo) and np.array_equal(Ig, Io))
    upd = T.S * T.nu ** 3
    return {"grid": list(T.n), "controls": T.nu ** 3, "states": T.S, "stages": stages, "ms_per_step": ms / stages,
            "value": upd / (ms / stages * 1e-3), "unit": UNIT, "kernel": "dense6",
            "hbm_frac": T.S * 52 / (ms / stages * 1e-3) / 1e9 / measured_peak_gbs()[0],
            "cpu_port_updates_per_s": 2 * Ts.S * Ts.nu ** 3 / dt_cpu, "cpu_cores": cbind.num_threads(),
            "parity_vs_oracle": "pass" if ok else "FAIL", "parity_sample": "2 stages on 12^3 x 6^3 from a rough terminal cost, bit-equal"}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "direct", "window", "splitc", "tile"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads (N = 1 default run)")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: CPU seconds per step (default: sized so that the run ends within minutes)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep equal slabs (no trial-run balancing)")
    ap.add_argument("--idx-bytes", type=int, default=4, choices=[1, 2, 4],
                    help="device storage of the argmin (4 = int32, the canonical 20 B/state of SURVEY 8d)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import bellman_b200 as bb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
        args.gpus = world
    torch.cuda.set_device(local)
    numa = None
    if world > 1:
        # one process per GPU: keep the host thread (and so the pinned buffers it first touches) on the CPUs
        # local to this rank's GPU, so the e2e copies do not cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            hdl = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = len(cpus)
        except Exception:      # no NVML / not permitted: run unbound
            numa = None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d = make_desc(bb, args.workload, world)
    # slab dimension: the one whose halo is smallest (host-side reach analysis, no GPU needed)
    def pick_part_dim(dd):
        if world == 1:
            return -1
        best = None
        for pd in range(dd.D):
            try:
                sl = bb.plan_slabs(dd, pd, world)
            except bb.BellmanError:
                continue
            cost = max((e - c) / max(b - a, 1) for a, b, c, e in sl)
            if best is None or cost < best[0] - 1e-9:
                best = (cost, pd)
        return best[1]

    part_dim = pick_part_dim(d)

    def open_sweep(pd, dd=None, cuts=None):
        s = bb.Sweep(dd if dd is not None else d, device=local, part_dim=pd, rank=rank, nranks=world, part_cuts=cuts,
                     idx_bytes=args.idx_bytes)
        if world > 1:
            ids = [bb.get_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            s.comm_init(ids[0])
        return s

    sw = open_sweep(part_dim)
    if world > 1 and sw.halo_mode == "nccl" and part_dim != d.D - 1:
        # without peer memory the send/recv fallback wants contiguous slabs: cut the last dimension
        sw.close()
        part_dim = d.D - 1
        sw = open_sweep(part_dim)
    kernel = {"auto": bb.KERNEL_AUTO, "direct": bb.KERNEL_DIRECT, "window": bb.KERNEL_WINDOW, "tile": bb.KERNEL_TILE,
              "splitc": bb.KERNEL_SPLITC}[args.kernel]
    use_graph = d.S * d.P < 4_000_000 and world == 1

    S_all = d.S * d.P
    upd_per_step = S_all * d.C
    K, W = args.steps, args.warmup
    if W + K + 2 > d.N - 1:
        raise SystemExit("steps+warmup exceed the horizon of this workload")

    # ---- slab balancing (untimed): the cost per index along the slab dimension is not uniform (clamped
    # queries, halo stores), so equal slabs leave the fastest rank waiting for the slowest at every stage.
    # Two trial rounds of 2 stages: per-rank kernel time -> new boundaries (bellman_desc.part_cuts)
    # proportional to the measured speed, multiples of 32 indices.
    cuts = None
    if world > 1 and not args.no_balance:
        n_p = d.n[part_dim]
        for _round in range(2):
            sw.run(2, kernel=kernel)
            stt = sw.stats()
            t_mine = torch.tensor([stt["ms"] - stt["ms_exchange"], float(sw.slab[1] - sw.slab[0])], dtype=torch.float64, device="cuda")
            allt = [torch.zeros_like(t_mine) for _ in range(world)]
            dist.all_gather(allt, t_mine)
            ms_r = np.array([float(t[0]) for t in allt])
            rows_r = np.array([float(t[1]) for t in allt])
            speed = rows_r / ms_r
            want = n_p * speed / speed.sum()
            edges = np.concatenate([[0.0], np.cumsum(want)])
            new = [int(round(e / 32.0)) * 32 for e in edges]
            new[0], new[-1] = 0, n_p
            if any(b - a < 64 for a, b in zip(new[:-1], new[1:])) or new == (cuts or []):
                break
            cuts = new
            sw.close()
            try:
                sw = open_sweep(part_dim, cuts=cuts)
            except bb.BellmanError:
                cuts = None
                sw = open_sweep(part_dim)
                break

    # ---- device-resident throughput ("value") ------------------------------------------------
    sw.run(W, kernel=kernel, use_graph=use_graph)                      # untimed warm-up stages
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    sw.run(K, kernel=kernel, use_graph=use_graph)                      # EXACTLY K stages, CUDA-event timed
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    st = sw.stats()
    if os.environ.get("BELLMAN_BENCH_VERBOSE"):
        print("rank %d slab %s: device ms %.3f, barrier/exchange ms %.3f, kernel %s" %
              (rank, sw.slab, st["ms"], st["ms_exchange"], sw.last_kernel), file=sys.stderr, flush=True)
    ms_dev = max_over_ranks(st["ms"])                                  # device time, max over ranks
    ms_x = max_over_ranks(st["ms_exchange"])
    clocks = sampler.stop() if rank == 0 else None
    launches = st["launches"]
    value = upd_per_step * K / (ms_dev * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers ("e2e") ---------------------------------
    # the seam-level drop-in: each step the caller hands J_{k+1} in host memory (as MATLAB's
    # F.Values would be), and reads J_k and the argmin back.
    e2e = None
    if not args.no_e2e:
        Ke = max(2, min(K, 5))
        own = sw.S_own * d.P
        pin_in = torch.empty(S_all, dtype=torch.float64).pin_memory().numpy()
        pin_J = torch.empty(own, dtype=torch.float64).pin_memory().numpy().reshape(d.P, -1)
        pin_I = torch.empty(own, dtype=torch.int32).pin_memory().numpy().reshape(d.P, -1)
        pin_in[:] = np.random.default_rng(7).normal(size=S_all) if S_all <= (1 << 27) else 0.0
        seq_call = "bellman_set_J(host) + bellman_run(1) + bellman_get_J(host) + bellman_get_idx(host)"

        brk = [0.0, 0.0, 0.0, 0.0]      # host time inside each blocking call of the sequential round trip

        def step_sequential():
            t = [time.perf_counter()]
            sw.set_J(pin_in.reshape(d.P, -1)); t.append(time.perf_counter())
            sw.run(1, kernel=kernel); t.append(time.perf_counter())
            sw.get_J(out=pin_J); t.append(time.perf_counter())
            sw.get_idx(out=pin_I); t.append(time.perf_counter())
            for k in range(4):
                brk[k] += t[k + 1] - t[k]

        def step_pipelined():       # one ABI call; copies overlap the kernel where the stage kernel can run tile ranges
            sw.stage_host(pin_in.reshape(d.P, -1), pin_J, pin_I, kernel=kernel)

        def timed(step):
            step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(Ke):
                step()
            barrier()
            return max_over_ranks(time.perf_counter() - t0)

        dt_seq = timed(step_sequential)
        brk_ms = [round(x / (Ke + 1) * 1e3, 3) for x in brk]
        ref_J, ref_I = pin_J.copy(), pin_I.copy()
        dt = timed(step_pipelined)
        same = bool(np.array_equal(ref_J, pin_J) and np.array_equal(ref_I, pin_I))
        pdim = part_dim if world > 1 else d.D - 1
        ext = (sw.slab[3] - sw.slab[2]) if world > 1 else d.n[pdim]
        h2d = int(S_all // d.n[pdim] * ext * 8)
        e2e = {"value": upd_per_step * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(own * 12), "steps": Ke, "ms_per_step": dt / Ke * 1e3,
               "call": "bellman_stage_host(J_next host in, J and idx host out)",
               "sequential": {"value": upd_per_step * Ke / dt_seq, "ms_per_step": dt_seq / Ke * 1e3, "call": seq_call,
                              "rank0_ms_set_J_run_get_J_get_idx": brk_ms},
               "matches_sequential": same}

    # ---- outside the timed region: bit-exact spot check of this rank's slab against the oracle -------
    def all_ok(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item()) == 1

    parity = {"main": "pass" if all_ok(sharded_parity(bb, sw, d, rank, world, part_dim, kernel)) else "FAIL"}
    if e2e is not None:          # the pipelined host call must reproduce the plain sequence bit for bit
        parity["e2e_pipelined_equals_sequential"] = "pass" if all_ok(e2e["matches_sequential"]) else "FAIL"
    main_kernel, main_halo, main_launches = sw.last_kernel, (sw.halo_mode if world > 1 else None), st["launches"]

    # ---- configs[4]: the pos-att grid sized to the HBM of the GPUs in use (weak scaling: 4.4e9 states per GPU)
    cfg5 = None
    if not args.no_others and args.workload == DEFAULT_WORKLOAD:
        sw.close()
        sw = None
        name5 = "pos_att_cfg5_480x480x160Nx120x9"
        free_b, _tot = torch.cuda.mem_get_info()
        d5 = make_desc(bb, name5, world)
        need = d5.S // world * 21 * 1.05
        fits = all_ok(free_b > need + (4 << 30))
        if fits:
            pd5 = pick_part_dim(d5)
            s5 = open_sweep(pd5, d5)
            s5.run(3)
            barrier()
            s5.run(5)
            barrier()
            t5 = s5.stats()
            ms5 = max_over_ranks(t5["ms"]) / 5
            mx5 = max_over_ranks(t5["ms_exchange"]) / 5
            ok5 = all_ok(sharded_parity(bb, s5, d5, rank, world, pd5, bb.KERNEL_AUTO, n_points=1500))
            parity["cfg5"] = "pass" if ok5 else "FAIL"
            peak5, _ = measured_peak_gbs()
            cfg5 = {"workload": name5, "grid": d5.n, "controls": d5.C, "states_per_gpu": d5.S // world,
                    "scaling": "weak", "partition": "dim %d slabs over %d ranks" % (pd5, world) if world > 1 else "none",
                    "halo": s5.halo_mode if world > 1 else None, "kernel": s5.last_kernel, "ms_per_step": ms5,
                    "exchange_ms_per_step": mx5, "value": d5.S * d5.C / (ms5 * 1e-3), "unit": UNIT,
                    "hbm_frac": d5.S // world * 20 / (ms5 * 1e-3) / 1e9 / peak5,
                    "bytes_per_gpu": int(d5.S // world * 20), "sharded_parity": parity["cfg5"]}
            s5.close()
        else:
            cfg5 = {"workload": name5, "skipped": "needs %.0f GB per GPU, %.0f GB free" % (need / 1e9, free_b / 1e9)}

    if rank != 0:
        if sw is not None:
            sw.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the stage kernel ---------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    kernel_name = main_kernel
    halo_mode = main_halo
    bytes_per_launch = (S_all / world) * (16 + args.idx_bytes)   # read J_{k+1}, write J_k, write the argmin (int32 unless --idx-bytes)
    ms_kernel = (ms_dev - ms_x) / K
    achieved = bytes_per_launch / (ms_kernel * 1e-3) / 1e9
    # secondary bound: fp64 pipe, 64 lanes/clk/SM x 148 SMs at the clock seen
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # SASS-counted, interior loop: k_stage_wide issues 15 instructions per update on the fp64 pipe (11 DADD, 3 DFMA,
    # 1 DSETP) plus 2 I2F.F64 on the conversion pipe; k_stage_window 17 on the fp64 pipe
    fp64_ops_per_update = 15.0 if kernel_name == "window:wide" else 17.0
    fp64_peak = 64 * 148 * mhz * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0],
                "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1], "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "idx_bytes": args.idx_bytes, "kernel_ms": ms_kernel, "kernel": kernel_name,
                "fp64_secondary": {"ops_per_update": fp64_ops_per_update,
                                   "achieved_ops_per_s": value / world * fp64_ops_per_update,
                                   "peak_ops_per_s": fp64_peak,
                                   "frac": value / world * fp64_ops_per_update / fp64_peak,
                                   "conversion_pipe_ops_per_update": 2.0 if kernel_name == "window:wide" else 0.0,
                                   "note": "C>=16 makes the stage fp64-issue bound, not HBM bound; the fp64 pipe and the "
                                           "shared-memory loads overlap only partly (DESIGN.md section 4)"}}

    # other configurations of BASELINE.json, device-resident, same timing rules (N = 1 only):
    # the small-control ones are where the HBM roofline is the relevant bound
    others = {}
    if world == 1 and not args.no_others and args.workload == DEFAULT_WORKLOAD:
        for name, steps in (("attitude_x16_3x16000x4800x3", 20), ("attitude_x4_3x4000x1200x3", 40),
                            ("position_3x201x201x3", 400), ("kirk_default_100x100x1000", 40),
                            ("pos_att_x4_120x120x80x60x9", 5)):
            d2 = make_desc(bb, name)
            s2 = bb.Sweep(d2, device=local)
            g2 = d2.S * d2.P < 4_000_000
            s2.run(3, use_graph=g2)
            s2.run(steps, use_graph=g2)
            t2 = s2.stats()
            ms2 = t2["ms"] / steps
            by2 = d2.S * d2.P * 20
            others[name] = {"ms_per_step": ms2, "value": d2.S * d2.P * d2.C / (ms2 * 1e-3), "kernel": s2.last_kernel,
                            "hbm_frac": by2 / (ms2 * 1e-3) / 1e9 / peak, "gpu_launches": t2["launches"],
                            "j_bytes": d2.S * d2.P * 8}
            s2.close()
        # the workload on which HBM IS the binding roofline (3 controls, J far larger than L2): same
        # definition of achieved / peak as the headline roofline object, reported inside it
        hb = others.get("attitude_x16_3x16000x4800x3")
        if hb and roofline is not None:
            roofline["hbm_bound_workload"] = {
                "workload": "attitude_x16_3x16000x4800x3", "kernel": hb["kernel"], "kernel_ms": hb["ms_per_step"],
                "achieved": hb["hbm_frac"] * peak, "peak": peak, "unit": "GB/s", "frac": hb["hbm_frac"],
                "bytes_per_launch": 3 * 16000 * 4800 * 20, "traffic": NCU_TRAFFIC["attitude_x16_3x16000x4800x3"][0],
                "traffic_source": NCU_TRAFFIC["attitude_x16_3x16000x4800x3"][1]}
        others["rollout_64x64_x0"] = rollout_rate(bb, local)
        # the two "next" rows of SURVEY 8f are reported beside the headline, never instead of it: a failure
        # here is recorded in the line and must not take the main measurement down with it
        for name, fn in (("pos_att_plant_rollout_4096_x0", plant_rollout_rate), ("attitude6_24x24x24x10x10x10x27", dense6_rate)):
            try:
                others[name] = fn(bb, local)
            except Exception as e:      # noqa: BLE001
                others[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r, cores, sample, _ = cpu_sample_rate(d, 8.0)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "reference_shaped": cpu_reference_shaped_rate(bb, args.workload)}
        slab = cpu_slab_rate(d, 6.0)
        if slab:
            cpu["contiguous_slab"] = {"value": slab[0], "unit": UNIT, "cores": slab[1], "sample": slab[2]}
        # the reference-shaped CPU rate of the small-control classes measured above (position, attitude, pos-att)
        for name in ("position_3x201x201x3", "attitude_x16_3x16000x4800x3", "pos_att_x4_120x120x80x60x9"):
            if name in others:
                others[name]["cpu_reference_shaped"] = cpu_reference_shaped_rate(bb, name)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, d),
            "run": {"partition": ("dim %d slabs over %d ranks; halo: %s" % (
                        part_dim, world,
                        "stored into peer memory by the stage kernel (NVLink P2P); stages ordered by neighbour-only release/acquire flags"
                        if halo_mode == "p2p" else "grouped ncclSend/ncclRecv after each stage"))
                    if world > 1 else "none",
                    "slab_cuts": cuts, "cuda_graph": bool(use_graph)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(main_launches), "roofline": roofline,
            "cpu_baseline": cpu, "wall_ms": wall_ms, "exchange_ms_per_step": ms_x / K,
            "sharded_parity": parity["main"], "parity_checks": parity, "cfg5": cfg5, "host_cpus_bound_per_rank": numa,
            "other_workloads": others}
    print(json.dumps(line), flush=True)
    if sw is not None:
        sw.close()
    if world > 1:
        dist.destroy_process_group()
    if "FAIL" in parity.values():
        raise SystemExit("sharded parity check FAILED: %s" % parity)


if __name__ == "__main__":
    main()
