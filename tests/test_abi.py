"""The C-ABI library loads without a GPU, exports every symbol include/bellman.h declares, and its
host-only entry points work.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bellman.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bellman_[a-z_0-9A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(bellman):
    lib = ctypes.CDLL(bellman.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"libbellman.so lacks {s}"
    assert sorted(bellman.EXPORTS) == syms
    assert lib.bellman_version() == 1


def test_struct_sizes_match(bellman):
    """bad struct_size must be rejected (ABI versioning)."""
    from bellman_b200 import _lib
    d = bellman.Dynamic_Solver()
    d.dx, d.du, d.N = 8, 5, 4
    desc = d._build()
    cd, keep = _lib.to_cdesc(desc)
    cd.struct_size = 12
    modes = np.zeros(2, dtype=np.int32)
    assert _lib.load().bellman_query_locate(ctypes.byref(cd), modes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))) == -1


def test_bad_descriptors_are_rejected(bellman):
    d = bellman.Dynamic_Solver()
    d.dx, d.du, d.N = 8, 5, 4
    desc = d._build()
    desc.grid[0] = desc.grid[0][:, ::-1].copy()      # not increasing
    with pytest.raises(bellman.BellmanError):
        bellman.query_locate(desc)


def test_slab_plan_covers_reach(bellman, oracle_lib):
    """ext range of each rank must contain every cell its owned states touch (checked by brute force)."""
    d = bellman.Dynamic_Solver()
    d.dx, d.du, d.N = 64, 40, 4
    desc = d._build()
    for nranks in (2, 3, 4):
        slabs = bellman.plan_slabs(desc, 1, nranks)
        assert slabs[0][0] == 0 and slabs[-1][1] == 64
        for r, (lo, hi, elo, ehi) in enumerate(slabs):
            assert 0 <= elo <= lo < hi <= ehi <= 64
            if r:
                assert lo == slabs[r - 1][1]
            # brute force: x'_2 for all (i, j in [lo,hi), c)
            x2 = (desc.Ta[1][0][:, None] + desc.Tb[1][0][None, lo:hi])[:, :, None] + desc.Tc[1][0][None, None, :]
            s = desc.grid[1][0]
            inv_h = (len(s) - 1) / (s[-1] - s[0])
            cell = np.clip(np.floor(x2 * inv_h - s[0] * inv_h), 0, len(s) - 2).astype(int)
            assert cell.min() >= elo and cell.max() + 2 <= ehi


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="box has a GPU")
def test_create_fails_loudly_without_gpu(bellman):
    d = bellman.Dynamic_Solver()
    d.dx, d.du, d.N = 8, 5, 4
    with pytest.raises(bellman.BellmanError) as ei:
        d.run()
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_mex_gateway_compiles():
    """The MEX gateway is syntax-checked against a stub mex.h (no MATLAB in this image)."""
    import subprocess
    mdir = os.path.join(ROOT, "optimal-control-dynamic-programming_b200", "matlab")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(mdir, "stub"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(mdir, "bellman_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for f in ("Dynamic_Solver.m", "Solver_position.m", "Solver_attitude.m", "Solver_pos_att.m"):
        src = open(os.path.join(mdir, f)).read()
        assert src.startswith("classdef " + f[:-2] + " < handle") and "bellman_mex('create'" in src


def test_degenerate_inputs_rejected_before_any_cuda_call(bellman):
    """Empty / degenerate inputs: every grid needs >= 2 points, C >= 1, N >= 2.  Validation runs
    before the library touches CUDA, so the code is BAD_ARG (-1) with or without a GPU."""
    t = bellman.tables
    o = bellman.Dynamic_Solver()

    def base():
        return t.kirk_desc(o.A, o.B, o.Q, o.R, 4, -1.0, 1.0, 8, -1.0, 1.0, 5)

    d1 = base()
    d1.n = [1, 8]
    for tab in (d1.grid, d1.Ta, d1.q):
        tab[0] = tab[0][:, :1]
    d1.Ta[1] = d1.Ta[1][:, :1]
    d2 = base()
    d2.C = 0
    d3 = base()
    d3.N = 1
    d4 = base()
    d4.q_order = [0, 0]
    for d in (d1, d2, d3, d4):
        with pytest.raises(bellman.BellmanError) as e:
            bellman.Sweep(d)
        assert e.value.code == -1, str(e.value)


def _brute_stencil(d, p=0):
    """cell(x'_k) - i_k over every state and control, by the exact bin rule in state units (numpy)."""
    out = []
    for k in range(d.D):
        s = d.grid[k][p]
        n = [len(d.grid[j][p]) for j in range(d.D)]
        sa, sb = d.src_a[k], d.src_b[k]
        if sa != k and sb != k:
            out.append(None)
            continue
        ta = d.Ta[k][p]
        x = ta[:, None, None] if sa == k else ta[None, :, None]
        own = np.arange(n[k])[:, None, None]
        if d.Tb[k] is not None and sb >= 0:
            tb = d.Tb[k][p]
            if sb == k and sa == k:
                x = (ta + tb)[:, None, None]
            elif sb == k:
                x = ta[None, :, None] + tb[:, None, None]
            else:
                x = x + tb[None, :, None]
        if d.Tc[k] is not None:
            x = x + d.Tc[k][p][None, None, :]
        cell = np.clip(np.searchsorted(s, x, side="right") - 1, 0, len(s) - 2)
        off = cell - own
        out.append((int(off.min()), int(off.max())))
    return out


def test_query_stencil_matches_brute_force(bellman):
    """bellman_query_stencil (host-only): the bounds the tile kernel sizes its box with are safe
    (contain every query) and tight; exact for SEARCH dimensions."""
    t = bellman.tables
    sp = bellman.Solver_pos_att()
    descs = []
    for mesh in ((30, 30, 20, 15), (12, 10, 8, 15), (34, 9, 7, 5), (64, 32, 16, 40)):
        sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = mesh
        descs += [sp.channel_desc(c) for c in range(3)]
    descs.append(t.attitude_axis_desc(-0.9, 0.9, 200, -30.0, 30.0, 120, [-0.11, 0.0, 0.11], 0.0285, 6.0, 6.0, 4.0, 0.02, 6))
    descs.append(t.position_axis_desc(-0.5, 0.5, 200, -0.5, 0.5, 200, [-0.26, 0.0, 0.26], 4.16, 6.0, 6.0, 0.1, 0.005, 6))
    descs.append(t.kirk_desc([[0.9974, 0.0539], [-0.1078, 1.1591]], [0.0013, 0.0539], [[0.25, 0.0], [0.0, 0.05]],
                             0.05, 5, -2.5, 3.0, 64, -40.0, 10.0, 33))
    for d in descs:
        lo, hi = bellman.query_stencil(d)
        modes = bellman.query_locate(d)[0]
        for k, b in enumerate(_brute_stencil(d)):
            if b is None:
                assert lo[k] > hi[k]
                continue
            assert lo[k] <= b[0] and hi[k] >= b[1], (d.meta.get("class"), k, lo[k], hi[k], b)      # safe
            assert lo[k] >= b[0] - 1 and hi[k] <= b[1] + 1, (d.meta.get("class"), k, lo[k], hi[k], b)  # tight
            if modes[k] != 0:                                                                        # SEARCH: exact
                assert (lo[k], hi[k]) == b
