"""Small instances of every hand-rolled mbarrier / TMA / peer-store kernel, for compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
Each case is also checked against the oracle (so a sanitizer-clean run is also a correct one)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_b200 as bb  # noqa: E402
from oracle import cbind  # noqa: E402


def check(name, d, n_stages, kernel, want, group=0, JN=None):
    ora = cbind.sweep(d, n_stages=n_stages, J_N=JN)
    sw = bb.SweepGroup(d, [0] * group) if group else bb.Sweep(d)
    if JN is not None:
        sw.set_J(JN)
    sw.run(n_stages, kernel=kernel)
    ok = sw.last_kernel == want and np.array_equal(sw.get_J(), ora["J_last"]) and np.array_equal(sw.get_idx(), ora["idx_last"])
    print("%-28s kernel=%-14s %s" % (name, sw.last_kernel, "OK" if ok else "MISMATCH"), flush=True)
    sw.close()
    return ok


def main():
    rng = np.random.default_rng(0)
    o = bb.Dynamic_Solver()
    kirk = bb.tables.kirk_desc(o.A, o.B, o.Q, o.R, 6, o.x_min, o.x_max, 96, o.u_min, o.u_max, 24,
                               store_J_all=False, store_idx_all=False)
    sa = bb.Solver_attitude()
    sa.n_mesh_w, sa.n_mesh_t = 96, 48
    att = bb.tables.stack_problems(sa._axis_descs())
    sp = bb.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 20, 6, 6, 9
    pa = sp.channel_desc(0)
    ok = True
    ok &= check("window wide (Kirk)", kirk, 3, bb.KERNEL_WINDOW, "window:wide", JN=rng.normal(size=(1, kirk.S)))
    os.environ["BELLMAN_NO_WIDE"] = "1"
    ok &= check("window ring (Kirk)", kirk, 3, bb.KERNEL_WINDOW, "window:ring", JN=rng.normal(size=(1, kirk.S)))
    del os.environ["BELLMAN_NO_WIDE"]
    ok &= check("strip (attitude)", att, 3, bb.KERNEL_WINDOW, "window:strip")
    ok &= check("stream (pos-att)", pa, 3, bb.KERNEL_TILE, "stream", JN=rng.normal(size=(1, pa.S)))
    os.environ["BELLMAN_NO_STREAM"] = "1"
    ok &= check("tile (pos-att)", pa, 2, bb.KERNEL_TILE, "tile")
    del os.environ["BELLMAN_NO_STREAM"]
    ok &= check("group x2 window wide", kirk, 3, bb.KERNEL_WINDOW, "window:wide", group=2)
    ok &= check("group x2 stream", pa, 3, bb.KERNEL_TILE, "stream", group=2)
    ok &= check("persistent (position-like)", kirk, 5, bb.KERNEL_AUTO, "splitc")
    print("SANITIZE_SMALL", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
